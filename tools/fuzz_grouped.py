"""Soak test of the grouped tcgen05 kernel: random group mixes against the C oracle, and repeated
launches of one large mix that must reproduce bit for bit."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget_s = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
rng = np.random.default_rng(seed)
dev = torch.device("cuda:0")
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t0 = time.time()
cases = 0
while time.time() - t0 < budget_s * 0.7:
    G = int(rng.integers(1, 400))
    hi = int(rng.choice([3, 40, 140, 300, 700]))
    nq = rng.integers(0, hi + 1, G)
    nt = rng.integers(0, hi + 1, G)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    if q_off[-1] == 0 or t_off[-1] == 0:
        continue
    qp, tp = synth.siftlike(int(q_off[-1]), rng), synth.siftlike(int(t_off[-1]), rng)
    if t_off[-1] > 20:
        tp[rng.integers(0, t_off[-1], 8)] = tp[rng.integers(0, t_off[-1], 8)]      # ties
    if min(q_off[-1], t_off[-1]) > 50:
        k = int(min(q_off[-1], t_off[-1]) // 3)
        qp[:k] = tp[:k]                                                          # exact matches
    od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
    d2, idx, t2q, _ = backend.grouped_mutual(d(qp), d(q_off), d(tp), d(t_off), algo=backend.FM_ALGO_TCGEN05)
    assert np.array_equal(d2.cpu().numpy().view(np.uint32), od2), (G, hi)
    assert np.array_equal(idx.cpu().numpy(), oidx), (G, hi)
    assert np.array_equal(t2q.cpu().numpy(), ot2q), (G, hi)
    cases += 1
qp, qo, tp, to = synth.make_groups(3000, 32, 512, seed=seed + 3)
args = (d(qp), d(qo), d(tp), d(to))
ref = None
reps = 0
while time.time() - t0 < budget_s:
    out = backend.grouped_mutual(*args, algo=backend.FM_ALGO_TCGEN05)
    cur = tuple(o.clone() for o in out[:3])
    if ref is None:
        ref = cur
    else:
        assert all(torch.equal(a, b) for a, b in zip(ref, cur)), "non-deterministic result"
    reps += 1
print("grouped fuzz ok: %d random mixes, %d identical launches" % (cases, reps))
