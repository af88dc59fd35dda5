"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference and cv2 exist; the fixtures and this script are committed, nothing at test
time reads /root/reference).

The reference ships no golden vectors for this path, so parity is
pinned on outputs of the reference's own matcher -- cv2.BFMatcher, called exactly as
matchutil.py:39-43 / fastmatch.pyx:122-123,161-162 / Classic Matching.ipynb cell 3 call it --
on the reference's graf fixtures and on seeded synthetic sets (ties, extremes, tiny sets).
    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fast_match_b200 import synth  # noqa: E402
from oracle import fastmatch_ref  # noqa: E402

REF = "/root/reference/images/graf"


def knn2(q, t):
    ms = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q.astype(np.float32), t.astype(np.float32), k=2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), np.inf, np.float32)
    for i, m in enumerate(ms):
        for c, mm in enumerate(m):
            idx[i, c], dist[i, c] = mm.trainIdx, mm.distance
    return idx, dist


def cross(q, t):
    qi, ti, dist = fastmatch_ref.cv2_mutual(q, t)
    return np.stack([qi, ti], 1).astype(np.int32), dist


def main():
    out = {}
    # ---- graf fixtures: images (lossless PNG), frozen SIFT descriptors, cv2 answers
    sift = cv2.SIFT_create()
    feats = {}
    for n in (1, 4):
        img = cv2.imread(os.path.join(REF, "img%d.ppm" % n))
        cv2.imwrite(os.path.join(HERE, "graf%d.png" % n), img, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        kp, ds = sift.detectAndCompute(img, None)
        assert np.array_equal(ds, np.rint(ds)) and ds.min() >= 0 and ds.max() <= 255
        feats[n] = (np.array([k.pt for k in kp], np.float32), ds.astype(np.uint8))
        out["graf%d_pos" % n], out["graf%d_desc" % n] = feats[n]
    q, t = feats[4][1], feats[1][1]
    out["graf41_knn2_idx"], out["graf41_knn2_dist"] = knn2(q, t)          # Ratio-Match, config 2
    out["graf41_cross_pairs"], out["graf41_cross_dist"] = cross(q, t)      # crossCheck k=1
    out["graf44_knn2_idx"], out["graf44_knn2_dist"] = knn2(q, q)          # self-match (Metric_Cache)
    d = out["graf41_knn2_dist"].astype(np.float64)
    out["graf41_ratio_counts"] = np.array([(d[:, 0] / d[:, 1] < tau).sum() for tau in (0.6, 0.7, 0.8, 0.9)])
    # ---- synthetic: planted matches + exact duplicates (ties)
    q, t = synth.make_pair(2000, 1500, seed=77)
    t[10:40] = t[700:730]
    q[5:25] = t[700:720]
    out["syn_q"], out["syn_t"] = q, t
    out["syn_knn2_idx"], out["syn_knn2_dist"] = knn2(q, t)
    out["syn_cross_pairs"], out["syn_cross_dist"] = cross(q, t)
    # ---- tiny / degenerate sets
    rng = np.random.default_rng(5)
    tiny_q = rng.integers(0, 256, (7, 128), dtype=np.uint8)
    tiny_t = rng.integers(0, 256, (2, 128), dtype=np.uint8)
    tiny_q[0] = 0; tiny_q[1] = 255; tiny_t[0] = 255
    out["tiny_q"], out["tiny_t"] = tiny_q, tiny_t
    out["tiny_knn2_idx"], out["tiny_knn2_dist"] = knn2(tiny_q, tiny_t)
    out["tiny1_knn2_idx"], out["tiny1_knn2_dist"] = knn2(tiny_q, tiny_t[:1])   # one train row: 1 match each
    out["tiny_cross_pairs"], out["tiny_cross_dist"] = cross(tiny_q, tiny_t)
    np.savez_compressed(os.path.join(HERE, "bf_golden.npz"), **out)

    # ---- config 1: the README example, Fast-Match graf img4 (query) -> img1 (target), run with
    # the sequential restatement of the driver and cv2.BFMatcher as the matcher (exact denominators)
    img1 = cv2.imread(os.path.join(REF, "img1.ppm"))
    cache = fastmatch_ref.RefMetricCache.from_image(os.path.join(REF, "img4.ppm"))
    fm = {}
    for tau in (0.7, 0.9):
        log = []
        gm = fastmatch_ref.match(cache, img1, {"log": log}, mutual=fastmatch_ref.cv2_mutual)
        ms = gm(tau)
        key = "tau%02d" % int(tau * 100)
        fm[key + "_index"] = np.array([m[0] for m in ms], np.int64)
        fm[key + "_pos"] = np.array([m[1]["positions"] for m in ms]).reshape(-1, 2, 2)
        fm[key + "_ratio"] = np.array([m[1]["ratio"] for m in ms])
        fm[key + "_rounds"] = np.array([gm.rounds, len(log)])
        fm[key + "_log_grid"] = np.array([l["target_grid"] for l in log]).reshape(-1, 2, 2)
        fm[key + "_log_nmatch"] = np.array([len(l["matches"]) for l in log])
        print(key, "matches", len(ms), "rounds", gm.rounds)
    fm["query_desc_hash"] = np.array([int(cache.original["descriptors"].astype(np.uint64).sum()),
                                      len(cache.original["descriptors"]), int(cache.thumb["descriptors"].astype(np.uint64).sum())])
    np.savez_compressed(os.path.join(HERE, "fastmatch_graf41.npz"), **fm)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
