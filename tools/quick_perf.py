"""Scratch timing of the dense kernels (CUDA events, L2 flushed between launches)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from fast_match_b200 import backend, synth

def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

def main():
    sizes = [(5000, 5000), (50000, 50000)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
    for M, N in sizes:
        q, t = synth.make_pair(M, N, seed=1237)
        qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        ref = None
        for name, algo in (("mma", backend.FM_ALGO_MMA_SYNC), ("tc", backend.FM_ALGO_TCGEN05)):
            if name == "mma" and (M * N > 3e9 or os.environ.get("FM_QUICK_TC_ONLY")): continue
            try:
                out = backend.top2(qd, td, algo=algo)
                torch.cuda.synchronize()
            except Exception as e:
                print(name, "FAILED", e); continue
            if ref is None: ref = out
            else: print("  agree:", bool((ref[0] == out[0]).all() and (ref[1] == out[1]).all()))
            med, mn = timeit(lambda: backend.top2(qd, td, algo=algo))
            ops = 2.0 * M * N * 128
            print(json.dumps(dict(M=M, N=N, algo=name, ms_med=round(med, 4), ms_min=round(mn, 4),
                                  tops=round(ops / med / 1e9, 1), frac_4p5=round(ops / med / 1e9 / 4500, 4))))
main()
