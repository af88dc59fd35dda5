"""Time the dense kernel built with extra -D defines (timing experiments; results may be wrong).
usage: python tools/variant_perf.py FM_EXPERIMENT_X[,FM_...] 50000x50000 200000x200000"""
import os, sys, subprocess
sys.path.insert(0, ".")
from fast_match_b200 import build
defs = [d for d in sys.argv[1].split(",") if d and d != "none"]
lib = os.path.join(os.path.dirname(build.LIB), "libfmatch_variant.so")
build.build(defines=defs, out=lib)
env = dict(os.environ, FM_LIB=lib)
sys.exit(subprocess.call([sys.executable, "tools/quick_perf.py"] + sys.argv[2:], env=env))
