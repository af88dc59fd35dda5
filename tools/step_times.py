"""Per-step device time of the first 40 steps after different kinds of untimed preparation
(what makes the first ~10 steps of a short run slower than the steady state?)."""
import sys; sys.path.insert(0, ".")
import torch, numpy as np
from fast_match_b200 import backend, synth
q, t = synth.make_pair(50000, 50000, seed=1237)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
out = (torch.empty((50000, 2), dtype=torch.int32, device="cuda"), torch.empty((50000, 2), dtype=torch.int32, device="cuda"), torch.empty(50000, dtype=torch.uint8, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label, prep):
    torch.cuda.synchronize()
    import time; time.sleep(1.0)           # let the GPU fall back to idle, as after host-side data generation
    prep()
    for _ in range(5):
        flush.zero_(); backend.ratio_match(qd, td, 0.7, out=out)
    torch.cuda.synchronize()
    evs = []
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); backend.ratio_match(qd, td, 0.7, out=out); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    print("%-28s first 20 mean %.4f  last 20 mean %.4f  first 6: %s" % (label, np.mean(ts[:20]), np.mean(ts[20:]), ["%.3f" % x for x in ts[:6]]))
def flushes(n):
    for _ in range(n): flush.zero_()
def gemm():
    a8 = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device="cuda"); b8 = a8.t()
    for _ in range(13): torch._int_mm(a8, b8)
def steps(n):
    for _ in range(n): backend.ratio_match(qd, td, 0.7, out=out)
run("nothing", lambda: None)
run("int8 GEMM x13", gemm)
run("40 flushes", lambda: flushes(40))
run("30 steps without flush", lambda: steps(30))
run("30 flush+step", lambda: [(flush.zero_(), backend.ratio_match(qd, td, 0.7, out=out)) for _ in range(30)])
