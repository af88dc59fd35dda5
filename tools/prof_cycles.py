"""In-kernel cycle breakdown of the tcgen05 kernel (needs the -DFM_TC_PROF build)."""
import ctypes, os, sys
sys.path.insert(0, ".")
from fast_match_b200 import build
lib = os.path.join(os.path.dirname(build.LIB), "libfmatch_prof.so")
build.build(defines=["FM_TC_PROF"], out=lib)
os.environ["FM_LIB"] = lib
import torch
from fast_match_b200 import backend, synth
for arg in sys.argv[1:] or ["50000"]:
    M = N = int(arg)
    q, t = synth.make_pair(M, N, seed=1237)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    L = backend.lib()
    out = (ctypes.c_ulonglong * 16)()
    backend.top2(qd, td, algo=2); L.fm_debug_prof(out, 1)
    backend.top2(qd, td, algo=2); L.fm_debug_prof(out, 1)
    v = list(out)
    ctas, tiles = max(v[9], 1), max(v[10], 1)
    print(f"N={N} ctas={v[9]} tiles/cta={tiles/ctas:.1f} cycles/cta={v[8]/ctas:.0f} cycles/tile={v[8]/tiles:.0f}")
    mt = max(v[11], 1)
    print(f"  MMA warp per tile: wait full={v[0]/mt:.0f} wait tmem_empty={v[1]/mt:.0f}; producer wait empty={v[2]/tiles:.0f}")
    n = max(v[7], 1)
    print(f"  epilogue per tile per warp: wait tmem_full={v[4]/n:.0f} busy={v[6]/n:.0f}")
