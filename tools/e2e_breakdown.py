import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from fast_match_b200 import backend, synth
M = N = 50000
q, t = synth.make_pair(M, N, seed=1237)
pin = lambda a: torch.from_numpy(a).pin_memory()
qp, tp = pin(q), pin(t)
qd, td = torch.empty_like(qp, device="cuda"), torch.empty_like(tp, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("H2D q+t (12.8 MB) ms:", timeit(lambda: (qd.copy_(qp, non_blocking=True), td.copy_(tp, non_blocking=True))))
print("top2 device ms:", timeit(lambda: backend.top2(qd, td)))
d2, idx = backend.top2(qd, td)
hd2, hidx = torch.empty((M, 2), dtype=torch.int32).pin_memory(), torch.empty((M, 2), dtype=torch.int32).pin_memory()
print("D2H d2+idx (0.8 MB) ms:", timeit(lambda: (hd2.copy_(d2, non_blocking=True), hidx.copy_(idx, non_blocking=True))))
o_d2, o_idx, o_mask = np.empty((M, 2), np.uint32), np.empty((M, 2), np.int32), np.empty(M, np.uint8)
pd2, pidx, pmask = hd2.numpy().view(np.uint32), hidx.numpy(), torch.empty(M, dtype=torch.uint8).pin_memory().numpy()
print("host API pinned ms:", timeit(lambda: backend.top2_host(qp.numpy(), tp.numpy(), want_dist=False, tau=0.7, out=(pd2, pidx, None, pmask)), 10))
print("host API pageable ms:", timeit(lambda: backend.top2_host(q, t, want_dist=False, tau=0.7, out=(o_d2, o_idx, None, o_mask)), 10))
