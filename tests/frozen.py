"""Frozen SIFT features of the README example (tests/golden/graf_features.npz, written by
tests/golden/make_features_golden.py): the parity tests replay them instead of running SIFT,
so they do not depend on the OpenCV build of the machine and can never skip."""
import collections
import hashlib
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KeyPoint = collections.namedtuple("KeyPoint", ["pt"])
_F = None


def _load():
    global _F
    if _F is None:
        g = np.load(os.path.join(GOLD, "graf_features.npz"))
        index = {str(k): i for i, k in enumerate(g["keys"])}
        _F = (index, g["off"], g["pos"], g["desc"], g)
    return _F


def pixel_key(img):
    a = np.ascontiguousarray(img)
    return hashlib.sha1(repr(a.shape).encode() + a.tobytes()).hexdigest()


def lookup(key):
    index, off, pos, desc, _ = _load()
    i = index[key]           # KeyError = the driver asked for pixels that were never frozen: a real failure
    return pos[off[i]:off[i + 1]], desc[off[i]:off[i + 1]]


def features(img, *args, **kwargs):
    """Drop-in for matchutil.get_features: (keypoints with .pt, float32 descriptors | None)."""
    pos, desc = lookup(pixel_key(img))
    kp = [KeyPoint((float(x), float(y))) for x, y in pos]
    return kp, (None if len(desc) == 0 else desc.astype(np.float32))


def query_cache():
    """The README example's query side (graf img4) as a RefMetricCache built from frozen features."""
    from oracle import fastmatch_ref
    g = _load()[4]
    tp, td = lookup(str(g["thumb4_key"]))
    p, d = lookup(str(g["img4_key"]))
    return fastmatch_ref.RefMetricCache(td, tp, tuple(int(v) for v in g["thumb4_size"]), d, p,
                                        tuple(int(v) for v in g["img4_size"]))


def target_image():
    import cv2
    return cv2.imread(os.path.join(GOLD, "graf1.png"))
