import sys; sys.path.insert(0, ".")
import numpy as np, torch
from fast_match_b200 import backend, synth
M = N = 50000
q, t = synth.make_pair(M, N, seed=1237)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
qp, tp = pin(q), pin(t)
out = (pin(np.empty((M, 2), np.uint32)), pin(np.empty((M, 2), np.int32)), None, pin(np.empty(M, np.uint8)))
for i in range(5):
    print("call", i, file=sys.stderr)
    backend.top2_host(qp, tp, want_dist=False, tau=0.7, out=out)
