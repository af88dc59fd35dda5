import sys, torch
sys.path.insert(0, ".")
from fast_match_b200 import backend, synth
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
q, t = synth.make_pair(M, N, seed=1237)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
for _ in range(3):
    backend.top2(qd, td, algo=backend.FM_ALGO_TCGEN05)
torch.cuda.synchronize()
