"""End-to-end time of fm_top2_host_u8 (pinned buffers): adaptive choice vs FM_HOST_HALVES = 1 / 2 on the same box."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, ".")
    import time, numpy as np, torch
    from fast_match_b200 import backend, synth
    M = N = 50000
    q, t = synth.make_pair(M, N, seed=1237)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    qp, tp = pin(q), pin(t)
    out = (pin(np.empty((M, 2), np.uint32)), pin(np.empty((M, 2), np.int32)), None, pin(np.empty(M, np.uint8)))
    for _ in range(3):
        backend.top2_host(qp, tp, want_dist=False, tau=0.7, out=out)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        backend.top2_host(qp, tp, want_dist=False, tau=0.7, out=out)
        ts.append(time.perf_counter() - t0)
    qd, td = torch.empty((M, 128), dtype=torch.uint8, device="cuda"), torch.empty((N, 128), dtype=torch.uint8, device="cuda")
    hs = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); qd.copy_(torch.from_numpy(qp), non_blocking=True); td.copy_(torch.from_numpy(tp), non_blocking=True); b.record(); torch.cuda.synchronize()
        hs.append(a.elapsed_time(b))
    print("halves=%s: median %.3f ms  min %.3f ms   (H2D of the 12.8 MB alone: %.3f ms)" % (sys.argv[1], 1e3 * sorted(ts)[15], 1e3 * min(ts), sorted(hs)[2]))
else:
    for h in ("auto", "1", "2"):
        env = dict(os.environ)
        if h != "auto":
            env["FM_HOST_HALVES"] = h
        subprocess.call([sys.executable, __file__, h], env=env)
