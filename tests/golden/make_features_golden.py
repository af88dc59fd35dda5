"""Freezes every SIFT feature set the README example (config 1) touches, so that the Fast-Match
parity tests replay the frozen features and never depend on the SIFT build of the machine they
run on (and therefore never skip):  python tests/golden/make_features_golden.py

Recorded (keyed by a hash of the pixels SIFT was given): the query's 600-px thumbnail and full
image (graf img4), the target's 400-px thumbnail and every grid cell of graf img1 that the flood
fill visits at tau 0.7 / 0.9 with the default options and with the two option sets of
tests/test_host_logic.py.  The matcher outputs for these features are in fastmatch_graf41.npz
(make_golden.py); this file only holds inputs."""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fast_match_b200 import imaging, matchutil  # noqa: E402
from oracle import fastmatch_ref  # noqa: E402


def pixel_key(img):
    a = np.ascontiguousarray(img)
    return hashlib.sha1(repr(a.shape).encode() + a.tobytes()).hexdigest()


def main():
    store = {}

    def recording(img, *a, **k):
        kp, ds = matchutil.get_features(img)
        pos = np.array([p.pt for p in kp], np.float32).reshape(-1, 2)
        ds = np.zeros((0, 128), np.uint8) if ds is None else ds
        assert np.array_equal(ds, np.rint(ds)) and (len(ds) == 0 or (ds.min() >= 0 and ds.max() <= 255))
        store[pixel_key(img)] = (pos, np.asarray(ds).astype(np.uint8))
        return kp, (None if len(ds) == 0 else np.asarray(ds, np.float32))

    g4 = os.path.join(HERE, "graf4.png")
    img4, img1 = cv2.imread(g4), cv2.imread(os.path.join(HERE, "graf1.png"))
    thumb4 = imaging.get_thumbnail(g4, (600, 600))
    recording(thumb4)
    recording(img4)
    cache = fastmatch_ref.RefMetricCache.from_image(g4)
    for opts in ({}, {"grid_size": (75, 75), "grid_margin": 30, "radius": 50}, {"thumb_strategy": lambda t: t * 1.2, "radius": 60}):
        for tau in (0.7, 0.9):
            fastmatch_ref.match(cache, img1, dict(opts), mutual=fastmatch_ref.cv2_mutual, features=recording)(tau)
    keys = sorted(store)
    off = np.zeros(len(keys) + 1, np.int64)
    np.cumsum([len(store[k][0]) for k in keys], out=off[1:])
    np.savez_compressed(os.path.join(HERE, "graf_features.npz"), keys=np.array(keys),
                        off=off, pos=np.concatenate([store[k][0] for k in keys]),
                        desc=np.concatenate([store[k][1] for k in keys]),
                        thumb4_key=pixel_key(thumb4), img4_key=pixel_key(img4),
                        thumb4_size=np.array([thumb4.shape[1], thumb4.shape[0]]), img4_size=np.array([img4.shape[1], img4.shape[0]]))
    print(len(keys), "feature sets,", int(off[-1]), "descriptors,", os.path.getsize(os.path.join(HERE, "graf_features.npz")), "bytes")


if __name__ == "__main__":
    main()
