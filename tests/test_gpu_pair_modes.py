"""GPU: both instantiations of the dense tcgen05 kernel (single CTA / CTA pair) on boundary shapes.
The library picks one (and its schedule: stream-K or whole M-blocks per worker) per call from the
shape; FM_TC_PAIR / FM_TC_ALIGNED force them, and are read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("pair,aligned", [("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")])
def test_forced_kernel_variant_matches_oracle(cuda, pair, aligned):
    env = dict(os.environ, FM_TC_PAIR=pair, FM_TC_ALIGNED=aligned)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_pair_mode_worker.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "pair-mode worker ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
