"""Prebuild (here, no GPU needed) and time (on the GPU box) several -D variants of the dense kernel.
  python tools/variant_batch.py build  name=DEF1,DEF2 ...     -> fast_match_b200/libfmatch_v_<name>.so
  python tools/variant_batch.py run SIZE [SIZE ...]           -> times every prebuilt variant (+ the product lib)"""
import glob, os, subprocess, sys
sys.path.insert(0, ".")
from fast_match_b200 import build
d = os.path.dirname(build.LIB)
if sys.argv[1] == "build":
    for spec in sys.argv[2:]:
        name, defs = spec.split("=", 1)
        build.build(defines=[x for x in defs.split(",") if x], out=os.path.join(d, "libfmatch_v_%s.so" % name))
        print("built", name)
else:
    libs = [("product", build.LIB)] + [(os.path.basename(p)[len("libfmatch_v_"):-3], p) for p in sorted(glob.glob(os.path.join(d, "libfmatch_v_*.so")))]
    for name, path in libs:
        print("=== variant", name, flush=True)
        subprocess.call([sys.executable, "tools/quick_perf.py"] + sys.argv[2:], env=dict(os.environ, FM_LIB=path, FM_QUICK_TC_ONLY="1"))
