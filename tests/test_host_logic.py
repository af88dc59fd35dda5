"""CPU: host-side logic of the drop-in layer (no GPU, no CUDA calls).

The wave-batched flood fill (fast_match_b200/fastmatch.py) must emit exactly what the
sequential round-by-round driver emits (oracle/fastmatch_ref.py, which restates
fastmatch.pyx:56-180).  Here the two backend entry points of the product driver are replaced
by oracle-backed stand-ins so that only the host logic is under test; the same comparison
runs against the real CUDA backend in tests/test_gpu_fastmatch.py.
"""
import os

import numpy as np
import pytest

import frozen
import oracle
from oracle import fastmatch_ref
from fast_match_b200 import cache as fm_cache
from fast_match_b200 import fastmatch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_grid_geometry_matches_reference_restatement():
    rng = np.random.default_rng(0)
    for (w, h, cw, ch, m) in [(800, 640, 50, 50, 25), (801, 637, 75, 60, 30), (120, 90, 50, 50, 10), (50, 50, 50, 50, 0)]:
        img = np.zeros((h, w, 3), np.uint8)
        a = fm_cache.Grid_Cache(img, (cw, ch), None, margin=m)
        b = fastmatch_ref.RefGrid(img, (cw, ch), lambda c: c.shape, m)
        assert (a.rows, a.cols) == (b.rows, b.cols)
        for _ in range(300):
            x, y = rng.uniform(0, w), rng.uniform(0, h)
            assert a.block(x, y) == b.block(x, y)
            assert tuple(a.offset(x, y)) == tuple(b.offset(x, y))
            col, row = a.block(x, y)
            assert np.array_equal(a.center(col, row), b.center(col, row))
            px, py = rng.uniform(0, w), rng.uniform(0, h)
            assert np.array_equal(a.get_neighbor(col, row, px, py), b.get_neighbor(col, row, px, py))
            cell = a.get(x, y)
            b.get(x, y)
            assert a.last == b.last
            (x0, x1), (y0, y1) = a.rect(col, row)
            assert cell.shape[:2] == (min(y1, h) - y0, min(x1, w) - x0)   # numpy clips the crop, the logged rect is not clipped
        with pytest.raises(Exception):
            a.get(w + 1, 1)


def _stub_backend(monkeypatch):
    """Replace the two CUDA entry points used by the driver with oracle-backed stand-ins."""
    def mutual_pairs(q, t):
        return fastmatch_ref.oracle_mutual(np.asarray(q), np.asarray(t))

    def run_groups(q, q_lists, t_pool, t_starts, t_counts):
        q_off = np.concatenate([[0], np.cumsum([len(x) for x in q_lists])]).astype(np.int64)
        t_off = np.concatenate([[0], np.cumsum(t_counts)]).astype(np.int64)
        tp = np.concatenate([t_pool[s:s + c] for s, c in zip(t_starts, t_counts)] + [np.zeros((0, 128), np.uint8)])
        gather = np.concatenate(list(q_lists) + [np.zeros(0, np.int64)]).astype(np.int32)
        d2, idx, t2q = oracle.c_grouped_mutual(np.asarray(q), q_off, tp, t_off, q_gather=gather)
        out = []
        for g in range(len(q_lists)):
            qs, ts = slice(q_off[g], q_off[g + 1]), slice(t_off[g], t_off[g + 1])
            keep = oracle.mutual_pairs(idx[qs], t2q[ts])
            out.append((keep, idx[qs][keep, 0].astype(np.int64), np.sqrt(d2[qs][keep, 0].astype(np.float32))))
        return out
    monkeypatch.setattr(fastmatch, "_mutual_pairs", mutual_pairs)
    monkeypatch.setattr(fastmatch, "_run_groups", run_groups)
    monkeypatch.setattr(fastmatch, "_to_pool", lambda u8, like: np.ascontiguousarray(u8))
    monkeypatch.setattr(fastmatch, "_pool_cat", lambda a, b: np.concatenate([a, b]))


def _same_matches(a, b):
    assert len(a) == len(b)
    for (ia, da), (ib, db) in zip(a, b):
        assert int(ia) == int(ib)
        assert np.array_equal(da["positions"], db["positions"]) and da["ratio"] == db["ratio"]


def _same_logs(la, lb):
    assert len(la) == len(lb)
    for x, y in zip(la, lb):
        assert set(x) == set(y) == {"query_pos", "target_pos", "target_grid", "matches", "radius", "ratios", "margin"}
        assert np.array_equal(x["query_pos"], y["query_pos"]) and np.array_equal(x["target_pos"], y["target_pos"])
        assert x["target_grid"] == y["target_grid"] and x["radius"] == y["radius"] and x["margin"] == y["margin"]
        assert np.array_equal(x["matches"], y["matches"]) and np.array_equal(x["ratios"], y["ratios"])


@pytest.fixture(scope="module")
def graf():
    """Query cache and target image of the README example, built from the FROZEN features."""
    return frozen.query_cache(), frozen.target_image()


@pytest.mark.parametrize("opts", [{}, {"grid_size": (75, 75), "grid_margin": 30, "radius": 50},
                                  {"thumb_strategy": lambda t: t * 1.2, "radius": 60}])
def test_wave_batched_flood_fill_equals_sequential(monkeypatch, graf, opts):
    cache, img1 = graf
    _stub_backend(monkeypatch)
    for tau in (0.7, 0.9):
        log_a, log_b, stats = [], [], {}
        got = fastmatch.match(cache, img1, dict(opts, log=log_a, stats=stats, features=frozen.features))(tau)
        ref = fastmatch_ref.match(cache, img1, dict(opts, log=log_b, features=frozen.features))
        want = ref(tau)
        _same_matches(got, want)
        _same_logs(log_a, log_b)
        assert stats["launches"] < max(ref.rounds, 2)          # rounds were batched into waves
        assert stats["rounds_evaluated"] >= ref.rounds


def test_match_many_equals_one_pair_at_a_time(monkeypatch, graf):
    """Several image pairs in lock step (their waves share the grouped launches, the query caches
    share one descriptor pool): every pair's matches and log equal a single-pair run."""
    cache, img1 = graf
    other = frozen.query_cache()                     # a second cache object: its rows get a non-zero pool base
    _stub_backend(monkeypatch)
    for tau in (0.7, 0.9):
        logs = [[], [], []]
        stats = {}
        many = fastmatch.match_many([cache, other, cache], [img1, img1, img1],
                                    {"log": logs, "features": frozen.features, "stats": stats})(tau)
        log_one = []
        one = fastmatch.match(cache, img1, {"log": log_one, "features": frozen.features})(tau)
        assert len(many) == 3
        for ms, lg in zip(many, logs):
            _same_matches(ms, one)
            _same_logs(lg, log_one)
    with pytest.raises(ValueError):
        fastmatch.match_many([cache], [img1, img1])


def test_speculation_batches_many_rounds_per_launch(monkeypatch, graf):
    cache, img1 = graf
    _stub_backend(monkeypatch)
    stats = {}
    ms = fastmatch.match(cache, img1, {"stats": stats, "features": frozen.features})(0.9)
    ref = fastmatch_ref.match(cache, img1, {"features": frozen.features})
    ref(0.9)
    assert len(ms) > 0 and stats["rounds_evaluated"] >= ref.rounds
    assert stats["rounds_evaluated"] >= 10 * stats["launches"]            # >= 10 rounds per grouped launch
    assert stats["rounds_evaluated"] <= 1.5 * ref.rounds                  # bounded speculation waste


def test_threaded_cell_features_equal_sequential(monkeypatch, graf):
    """A wave's new cells are extracted on several host threads (Grid_Cache.cache_many): same
    features per cell as one-by-one visits (real cv2 SIFT), and the same matches whatever the
    thread count (frozen features)."""
    import cv2
    from fast_match_b200 import matchutil
    cache, img1 = graf
    cells = [(c, r) for c in range(3, 7) for r in range(4, 9)]
    seq = fm_cache.Grid_Cache(img1, (50, 50), matchutil.get_features, margin=25)
    par = fm_cache.Grid_Cache(img1, (50, 50), matchutil.get_features, margin=25)
    seq.cache_many(cells, None)
    par.cache_many(cells + cells[:3], fastmatch._executor(4))          # duplicates are ignored
    one = fm_cache.Grid_Cache(img1, (50, 50), matchutil.get_features, margin=25)
    for col, row in cells:
        (k1, d1), (k2, d2), (k3, d3) = seq.grid[col][row], par.grid[col][row], one.get_cell(col, row)
        assert [k.pt for k in k1] == [k.pt for k in k2] == [k.pt for k in k3]
        assert (d1 is None and d2 is None and d3 is None) or (np.array_equal(d1, d2) and np.array_equal(d1, d3))
    assert par.last is None and one.last == one.rect(*cells[-1])
    _stub_backend(monkeypatch)
    a = fastmatch.match(cache, img1, {"features": frozen.features, "sift_threads": 1})(0.9)
    b = fastmatch.match(cache, img1, {"features": frozen.features, "sift_threads": 8})(0.9)
    _same_matches(a, b)


def test_readme_example_against_frozen_cv2_run(monkeypatch, graf):
    """Config 1: Fast-Match graf img4 -> img1 at tau 0.7 / 0.9 equals the frozen run that used
    cv2.BFMatcher as the matcher.  The inputs are the frozen features (tests/frozen.py), so the test
    does not depend on this machine's SIFT and never skips."""
    cache, img1 = graf
    gold = np.load(os.path.join(GOLD, "fastmatch_graf41.npz"))
    h = gold["query_desc_hash"]
    assert (int(cache.original["descriptors"].astype(np.uint64).sum()), len(cache.original["descriptors"]),
            int(cache.thumb["descriptors"].astype(np.uint64).sum())) == (int(h[0]), int(h[1]), int(h[2]))
    _stub_backend(monkeypatch)
    for tau in (0.7, 0.9):
        key = "tau%02d" % int(tau * 100)
        log = []
        ms = fastmatch.match(cache, img1, {"log": log, "features": frozen.features})(tau)
        assert np.array_equal(np.array([m[0] for m in ms], np.int64), gold[key + "_index"])
        assert np.array_equal(np.array([m[1]["positions"] for m in ms]).reshape(-1, 2, 2), gold[key + "_pos"])
        assert np.array_equal(np.array([m[1]["ratio"] for m in ms]), gold[key + "_ratio"])
        assert len(log) == int(gold[key + "_rounds"][1])
        assert np.array_equal(np.array([l["target_grid"] for l in log]).reshape(-1, 2, 2), gold[key + "_log_grid"])
        assert np.array_equal(np.array([len(l["matches"]) for l in log]), gold[key + "_log_nmatch"])


def test_matchlist_behaves_like_knnmatch_output():
    from fast_match_b200 import matchutil
    idx = np.array([[3, 1], [2, -1], [-1, -1]], np.int32)
    d2 = np.array([[4, 9], [16, 0xFFFFFFFF], [0xFFFFFFFF, 0xFFFFFFFF]], np.int64)
    ml = matchutil.MatchList(idx, d2, idx >= 0)
    assert len(ml) == 3 and [len(m) for m in ml] == [2, 1, 0]
    assert ml[0][1].trainIdx == 1 and ml[0][1].queryIdx == 0 and ml[0][1].distance == 3.0 and ml[0][0].imgIdx == 0
    assert ml[-1] == [] and [m[0].distance for m in ml[:2]] == [2.0, 4.0]
    with pytest.raises(ValueError):
        matchutil.to_u8(np.full((2, 128), 0.5, np.float32))       # not integer valued: rejected, not rounded
    assert matchutil.to_u8(None).shape == (0, 128)


def test_cache_key_is_ripemd160_with_or_without_openssl_support(monkeypatch):
    """The reference names cache files RIPEMD-160(path) (cache.pyx:193): the pure-Python fallback
    must give the same digests as OpenSSL's."""
    import hashlib
    vectors = {b"": "9c1185a5c5e9fc54612808977ee8f548b2258d31", b"abc": "8eb208f7e05d987a9b044a8e98c6b087f15a0bfc",
               b"message digest": "5d0689ef49d2fae572b881b123a85ffa21595f36",
               b"a" * 1000000: "52783243c1697bdbe16d37f97f68f08325dc1528"}
    real_new = hashlib.new

    def no_ripemd(name, *a, **k):
        if name == "ripemd160":
            raise ValueError("unsupported hash type ripemd160")
        return real_new(name, *a, **k)
    monkeypatch.setattr(hashlib, "new", no_ripemd)
    for msg, want in list(vectors.items())[:3]:
        assert fm_cache._ripemd160(msg) == want
    assert fm_cache._ripemd160(b"a" * 10000) == "eb33e86b2400cc0a11707be717a35a9acf074a58"
    mc = fm_cache.Metric_Cache.__new__(fm_cache.Metric_Cache)
    mc.path = "data/img4.ppm"
    assert mc._key() == fm_cache._ripemd160(b"data/img4.ppm") and len(mc._key()) == 40


def test_no_cpu_fallback_without_cuda():
    import torch
    from fast_match_b200 import backend, matchutil
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(backend.FastMatchError):
        matchutil.bf_match(np.zeros((4, 128), np.uint8), np.zeros((4, 128), np.uint8), k=2)
