"""Soak test of the dense tcgen05 kernel: random shapes against the C oracle (sampled rows) and
repeated large runs that must reproduce bit for bit (a race in the pipeline would show as a diff)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget_s = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
rng = np.random.default_rng(seed)
dev = torch.device("cuda:0")
t0 = time.time()
cases = 0
while time.time() - t0 < budget_s * 0.6:
    M = int(rng.choice([rng.integers(1, 600), rng.integers(600, 6000), rng.integers(6000, 40000)]))
    N = int(rng.choice([rng.integers(1, 600), rng.integers(600, 6000), rng.integers(6000, 60000)]))
    q, t = synth.make_pair(M, N, seed=int(rng.integers(1 << 30)))
    if N > 10 and rng.random() < 0.5:
        t[rng.integers(0, N, 5)] = t[rng.integers(0, N, 5)]          # duplicates -> ties
    base = int(rng.integers(0, 1000))
    sel = np.unique(rng.integers(0, M, min(M, 200)))
    od2, oidx = oracle.c_top2(q[sel], t, base)
    d2, idx = backend.top2(torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev), t_index_base=base,
                           algo=backend.FM_ALGO_TCGEN05)
    assert np.array_equal(d2.cpu().numpy().view(np.uint32)[sel], od2), (M, N, base)
    assert np.array_equal(idx.cpu().numpy()[sel], oidx), (M, N, base)
    cases += 1
q, t = synth.make_pair(50000, 50000, seed=seed + 7)
qd, td = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
ref = None
reps = 0
while time.time() - t0 < budget_s:
    d2, idx = backend.top2(qd, td, algo=backend.FM_ALGO_TCGEN05)
    cur = (d2.clone(), idx.clone())
    if ref is None:
        ref = cur
    else:
        assert torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]), "non-deterministic result"
    reps += 1
print("fuzz ok: %d random shapes, %d identical 50k x 50k runs" % (cases, reps))
