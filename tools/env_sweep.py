"""Time the dense call under different environment settings (one process each).
usage: python tools/env_sweep.py SIZE "VAR=a,b,c" ["VAR2=x,y"]   -> cartesian product"""
import itertools, os, subprocess, sys
size = sys.argv[1]
axes = []
for spec in sys.argv[2:]:
    k, vs = spec.split("=")
    axes.append([(k, v) for v in vs.split(",")])
for combo in itertools.product(*axes):
    env = dict(os.environ, FM_QUICK_TC_ONLY="1")
    env.update(dict(combo))
    print("===", " ".join("%s=%s" % kv for kv in combo), flush=True)
    subprocess.call([sys.executable, "tools/quick_perf.py", size], env=env)
