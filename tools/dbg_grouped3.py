import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
nq = np.array([int(x) for x in sys.argv[1].split(",")]); nt = np.array([int(x) for x in sys.argv[2].split(",")])
q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64); t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
rng = np.random.default_rng(1)
qp, tp = synth.siftlike(int(q_off[-1]), rng), synth.siftlike(int(t_off[-1]), rng)
od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
try:
    out = backend.grouped_mutual(d(qp), d(q_off), d(tp), d(t_off), algo=2)
    torch.cuda.synchronize()
    print(sys.argv[1], sys.argv[2], np.array_equal(out[0].cpu().numpy().view(np.uint32), od2), np.array_equal(out[1].cpu().numpy(), oidx), np.array_equal(out[2].cpu().numpy(), ot2q), flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "EXC", str(e)[:80], flush=True)
