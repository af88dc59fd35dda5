"""Phase timeline of the dense tcgen05 kernel (globaltimer stamps, -DFM_TC_PROF build).
usage: python tools/timeline.py [M N]   (prebuilt fast_match_b200/libfmatch_prof.so is used if present)"""
import ctypes, os, sys
sys.path.insert(0, ".")
import numpy as np
from fast_match_b200 import build
lib = os.path.join(os.path.dirname(build.LIB), "libfmatch_prof.so")
if not os.path.exists(lib):
    build.build(defines=["FM_TC_PROF"], out=lib)
os.environ["FM_LIB"] = lib
import torch
from fast_match_b200 import backend, synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
N = int(sys.argv[2]) if len(sys.argv) > 2 else M
q, t = synth.make_pair(M, N, seed=1237)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
L = backend.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["entry", "setup done", "first MMA", "last commit", "epi first acc", "epi done", "exit"]
for it in range(3):
    flush.zero_()
    backend.profile_enable(True); backend.profile_read(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); backend.top2(qd, td, algo=2); b.record(); torch.cuda.synchronize()
    kms, kn = backend.profile_read(reset=True)
    buf = (ctypes.c_ulonglong * 1280)()
    L.fm_debug_timeline(buf)
    tl = np.array(list(buf), dtype=np.float64).reshape(160, 8)
    live = tl[:, 0] > 0
    tl = tl[live]
    t0 = tl[:, 0].min()
    print("run %d: call %.1f us, main kernel (events) %.1f us, CTAs %d" % (it, 1e3 * a.elapsed_time(b), 1e3 * kms, live.sum()))
    for k, nm in enumerate(names):
        col = tl[:, k]
        col = col[col > 0]
        if len(col):
            r = (col - t0) / 1e3
            print("   %-14s min %8.1f  med %8.1f  max %8.1f us   (n=%d)" % (nm, r.min(), np.median(r), r.max(), len(r)))
out = (ctypes.c_ulonglong * 16)()
L.fm_debug_prof(out, 1)
backend.top2(qd, td, algo=2); L.fm_debug_prof(out, 1)
v = list(out)
ctas, tiles = max(v[9], 1), max(v[10], 1)
mt = max(v[11], 1)
n = max(v[7], 1)
print("cycles/cta %.0f tiles/cta %.1f cycles/tile %.0f" % (v[8] / ctas, tiles / ctas, v[8] / tiles))
print("MMA warp per tile: wait full %.0f, wait tmem_empty %.0f, wait a_full per tile %.0f; producer wait empty %.0f" % (v[0] / mt, v[1] / mt, v[3] / mt, v[2] / tiles))
print("epilogue per tile per warp: wait tmem_full %.0f busy %.0f" % (v[4] / n, v[6] / n))
# per-worker table: work range -> segments, and when the worker's epilogue finished
if M == 50000 and os.environ.get("FM_TL_TABLE", "1") == "1":
    plan = (ctypes.c_longlong * 8)()
    L.fm_debug_plan(ctypes.c_int64(M), ctypes.c_int64(N), plan)
    pair, mrows, mblocks, ntiles, workers = plan[0], plan[1], plan[2], plan[3], plan[4]
    work = mblocks * ntiles
    print("plan: pair %d mblock_rows %d mblocks %d ntiles %d workers %d" % (pair, mrows, mblocks, ntiles, workers))
    step = 2 if pair else 1
    for w in range(workers):
        b, e = work * w // workers, work * (w + 1) // workers
        segs = []
        p = b
        while p < e:
            nb = min(e, (p // ntiles + 1) * ntiles)
            segs.append(nb - p)
            p = nb
        end = (tl[w * step, 5] - t0) / 1e3
        print("w %2d steps %d segs %-16s first_tile %3d  epi done %.1f us" % (w, e - b, segs, b % ntiles, end))
