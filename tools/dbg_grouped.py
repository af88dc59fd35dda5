import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth
cases = [([5], [4]), ([40], [300]), ([130], [20]), ([0, 5, 130], [4, 0, 1]), ([600], [700])]
for nq, nt in cases:
    nq, nt = np.array(nq), np.array(nt)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    rng = np.random.default_rng(1)
    qp, tp = synth.siftlike(max(int(q_off[-1]), 1), rng)[:int(q_off[-1])], synth.siftlike(max(int(t_off[-1]), 1), rng)[:int(t_off[-1])]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = backend.grouped_mutual(d(qp), d(q_off), d(tp), d(t_off), algo=2)
    torch.cuda.synchronize()
    od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
    ok = (np.array_equal(out[0].cpu().numpy().view(np.uint32), od2), np.array_equal(out[1].cpu().numpy(), oidx), np.array_equal(out[2].cpu().numpy(), ot2q))
    print(list(nq), list(nt), ok, flush=True)
