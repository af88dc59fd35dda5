import ctypes, os, sys
sys.path.insert(0, ".")
from fast_match_b200 import build
lib = os.path.join(os.path.dirname(build.LIB), "libfmatch_prof.so")
build.build(defines=["FM_TC_PROF"], out=lib)
os.environ["FM_LIB"] = lib
import torch, bench
from fast_match_b200 import backend
class A: groups = 10000
L = backend.lib(); out = (ctypes.c_ulonglong * 16)()
bench.grouped_leg(A, torch.device("cuda:0"), {}, backend, False); L.fm_debug_gprof(out, 1)
a = bench.grouped_leg(A, torch.device("cuda:0"), {}, backend, False); L.fm_debug_gprof(out, 1)
v = list(out); calls = 8  # 3 warm + 5 timed
ctas = v[12]; units = v[2]; eu = v[10]
print("ms", a["ms"], "ctas", ctas, "units(all calls)", units, "cycles/cta/call", v[11] / ctas)
print("producer per unit: wait empty %.0f  own work %.0f" % (v[0] / units, v[1] / units))
print("mma per unit: wait full %.0f  wait tmem_empty %.0f  issue %.0f" % (v[3] / units, v[4] / units, v[5] / units))
print("epilogue(warp4 = group 0) per own unit: wait tmem_full %.0f  loads+keys+minima %.0f  state merge %.0f  write-out %.0f" % (v[6] / eu, v[7] / eu, v[8] / eu, v[9] / eu))
pieces = max(v[15], 1)
print("  per 32-column piece (warp 4): TMEM load + wait %.0f  keys %.0f  (pieces per own unit %.2f)" % (v[13] / pieces, v[14] / pieces, pieces / eu))
