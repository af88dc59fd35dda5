"""Time the grouped leg (config 4) and the dense 50k call for the product lib and every prebuilt libfmatch_v_*.so."""
import glob, os, subprocess, sys
sys.path.insert(0, ".")
from fast_match_b200 import build
d = os.path.dirname(build.LIB)
libs = [("product", build.LIB)] + [(os.path.basename(p)[len("libfmatch_v_"):-3], p) for p in sorted(glob.glob(os.path.join(d, "libfmatch_v_*.so")))]
for name, path in libs:
    print("=== variant", name, flush=True)
    env = dict(os.environ, FM_LIB=path, FM_QUICK_TC_ONLY="1")
    subprocess.call([sys.executable, "tools/prof_grouped.py"], env=env)
    if "--dense" in sys.argv:
        subprocess.call([sys.executable, "tools/quick_perf.py", "50000x50000"], env=env)
