import sys; sys.path.insert(0, ".")
import torch, numpy as np
from fast_match_b200 import backend, synth
q, t = synth.make_pair(50000, 50000, seed=1237)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
out = (torch.empty((50000, 2), dtype=torch.int32, device="cuda"), torch.empty((50000, 2), dtype=torch.int32, device="cuda"), torch.empty(50000, dtype=torch.uint8, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5): backend.ratio_match(qd, td, 0.7, out=out)
torch.cuda.synchronize()
evs = []
for _ in range(40):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); backend.ratio_match(qd, td, 0.7, out=out); b.record(); evs.append((a, b))
torch.cuda.synchronize()
print(["%.3f" % a.elapsed_time(b) for a, b in evs])
