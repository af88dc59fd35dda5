import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
qpool, q_off, tpool, t_off = synth.make_groups(int(sys.argv[1]) if len(sys.argv) > 1 else 200, 1, 300, seed=21)
od2, oidx, ot2q = oracle.c_grouped_mutual(qpool, q_off, tpool, t_off)
for algo in (1, 2, 2, 2):
    try:
        out = backend.grouped_mutual(d(qpool), d(q_off), d(tpool), d(t_off), algo=algo)
        torch.cuda.synchronize()
        ok = (np.array_equal(out[0].cpu().numpy().view(np.uint32), od2), np.array_equal(out[1].cpu().numpy(), oidx), np.array_equal(out[2].cpu().numpy(), ot2q))
        print("algo", algo, ok, flush=True)
        if not all(ok):
            bad = np.nonzero((out[0].cpu().numpy().view(np.uint32) != od2).any(1) | (out[1].cpu().numpy() != oidx).any(1))[0]
            g = np.searchsorted(q_off, bad, side="right") - 1
            print(" bad rows", len(bad), "groups", np.unique(g)[:10], "first", bad[:5], out[1].cpu().numpy()[bad[:3]], oidx[bad[:3]], out[0].cpu().numpy().view(np.uint32)[bad[:3]], od2[bad[:3]])
            badt = np.nonzero(out[2].cpu().numpy() != ot2q)[0]
            print(" bad t", len(badt), badt[:5])
    except Exception as e:
        print("algo", algo, "EXC", str(e)[:200], flush=True); break
