"""GPU: the drop-in layer (matchutil / Metric_Cache / fastmatch.match) on the CUDA backend
against the sequential CPU restatement and the frozen cv2 outputs."""
import os

import numpy as np
import pytest
import torch

import frozen
import oracle
from oracle import fastmatch_ref
from fast_match_b200 import backend, cache as fm_cache, fastmatch, matchutil, sharded, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = np.load(os.path.join(GOLD, "bf_golden.npz"))


def test_bf_match_equals_frozen_cv2_knnmatch(cuda):
    """matchutil.bf_match(k=2) and flann_match(k=2) on the graf descriptors == cv2.BFMatcher."""
    q, t = G["graf4_desc"].astype(np.float32), G["graf1_desc"].astype(np.float32)   # the reference's dtype
    for fn in (matchutil.bf_match, matchutil.flann_match):
        ml = fn(q, t, k=2)
        assert np.array_equal(ml.indices, G["graf41_knn2_idx"]) and np.array_equal(ml.distances, G["graf41_knn2_dist"])
    m = ml[17]
    assert (m[0].queryIdx, m[0].trainIdx, m[0].distance) == (17, int(G["graf41_knn2_idx"][17, 0]), float(G["graf41_knn2_dist"][17, 0]))
    ratios = np.array([mm[0].distance / mm[1].distance for mm in ml])                # Classic Matching cell 3
    assert [int((ratios < tau).sum()) for tau in (0.6, 0.7, 0.8, 0.9)] == list(G["graf41_ratio_counts"])
    # crossCheck is honoured only for k == 1 (matchutil.py:41)
    cc = matchutil.bf_match(q, t, k=1, options={"crossCheck": True})
    pairs = np.array([(m[0].queryIdx, m[0].trainIdx) for m in cc if len(m) > 0], np.int32)
    assert np.array_equal(pairs, G["graf41_cross_pairs"])
    assert np.array_equal(np.array([m[0].distance for m in cc if len(m) > 0], np.float32), G["graf41_cross_dist"])
    k1 = matchutil.bf_match(q, t, k=1)
    assert np.array_equal(k1.indices[:, 0], G["graf41_knn2_idx"][:, 0]) and all(len(m) == 1 for m in k1[:50])
    # one train row, k=2: one match per query (cv2 returns 1-element lists)
    one = matchutil.bf_match(G["tiny_q"], G["tiny_t"][:1], k=2)
    assert [len(m) for m in one] == [1] * len(G["tiny_q"])
    assert np.array_equal(one.indices, G["tiny1_knn2_idx"])


def test_metric_cache_self_distances_and_get(cuda, tmp_path):
    path = os.path.join(GOLD, "graf4.png")
    mc = fm_cache.Metric_Cache(path, {"cache_dir": str(tmp_path)})
    assert mc.original["descriptors"].is_cuda and mc.original["descriptors"].dtype == torch.uint8
    for slot in (mc.thumb, mc.original):
        u8 = slot["descriptors"].cpu().numpy()
        d2, _ = oracle.c_top2(u8, u8)
        assert np.array_equal(slot["distances"], np.sqrt(d2[:, 1].astype(np.float32)).astype(np.float64))
        assert set(("descriptors", "positions", "distances", "size")) <= set(slot)
    ds, pos, dis, idx = mc.get(400, 300, 100)
    assert len(idx) > 0 and np.array_equal(ds.cpu().numpy(), mc.original["descriptors"].cpu().numpy()[idx])
    d = np.linalg.norm(mc.original["positions"][idx] - np.array([400, 300]), axis=1)
    assert (d <= 100).all() and (np.diff(d) >= 0).all()          # sorted by distance (cache.pyx:179)
    # persistence round trip (cache.pyx:191-239)
    mc2 = fm_cache.Metric_Cache(path, {"cache_dir": str(tmp_path)})
    assert torch.equal(mc2.original["descriptors"], mc.original["descriptors"])
    assert np.array_equal(mc2.thumb["distances"], mc.thumb["distances"]) and mc2.original["size"] == mc.original["size"]
    assert np.array_equal(mc2.get_indices(400, 300, 100), idx)


def test_metric_cache_reads_the_reference_layout(cuda, tmp_path):
    """A cache directory written by the reference (cache.pyx:200-210): float32 integer-valued
    descriptors, FLANN (approximate) self-match distances, the BallTree pickled into a 0-d object
    array, file names RIPEMD-160(path).  It must load without unpickling anything, with the
    distances replaced by the exact self-match."""
    import pickle
    from sklearn.neighbors import BallTree
    path = "some/dir/img4.ppm"
    key = fm_cache._ripemd160(path.encode("utf-8"))
    desc, pos = G["graf4_desc"], G["graf4_pos"].astype(np.float64)
    thumb_desc, thumb_pos = desc[:700], pos[:700] * 0.5
    od2, _ = oracle.c_top2(desc, desc)
    exact = np.sqrt(od2[:, 1].astype(np.float32)).astype(np.float64)
    approx = exact * 1.07                                            # what an approximate index may have stored
    tree = np.array(pickle.dumps(BallTree(pos, metric="minkowski")), dtype=object)   # 0-d object array, as numpy.savez stores it
    np.savez(str(tmp_path / key), descriptors=desc.astype(np.float32), positions=pos, distances=approx,
             position_tree=tree, size=(800, 640))
    np.savez(str(tmp_path / (key + "_thumb")), positions=thumb_pos, descriptors=thumb_desc.astype(np.float32),
             distances=np.ones(700), size=(600, 480))
    mc = fm_cache.Metric_Cache(path, {"cache_dir": str(tmp_path)})      # would try to open the image if load() failed
    assert mc.original["descriptors"].dtype == torch.uint8 and np.array_equal(mc.original["descriptors"].cpu().numpy(), desc)
    assert np.array_equal(mc.original["distances"], exact)              # recomputed, not the stored approximation
    td2, _ = oracle.c_top2(thumb_desc, thumb_desc)
    assert np.array_equal(mc.thumb["distances"], np.sqrt(td2[:, 1].astype(np.float32)).astype(np.float64))
    assert mc.original["size"] == (800, 640) and mc.thumb["size"] == (600, 480)
    idx = mc.get_indices(400, 300, 100)
    d = np.linalg.norm(pos[idx] - np.array([400, 300]), axis=1)
    assert len(idx) > 0 and (d <= 100).all() and (np.diff(d) >= 0).all()
    # files written by this class carry the flag and are trusted as they are
    mc.save(str(tmp_path / "new"))
    mc2 = fm_cache.Metric_Cache(path, {"cache_dir": str(tmp_path / "new")})
    assert np.array_equal(mc2.original["distances"], exact) and torch.equal(mc2.thumb["descriptors"], mc.thumb["descriptors"])


def test_two_devices_in_one_process(cuda):
    """Per-device state of the library (shared-memory opt-in, SM count, cluster support, host-path
    context): the second GPU of a process must behave like the first."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    q, t = synth.make_pair(3000, 2500, seed=42)
    od2, oidx = oracle.c_top2(q, t)
    qpool, q_off, tpool, t_off = synth.make_groups(40, 16, 200, seed=7)
    gd2, gidx, gt2q = oracle.c_grouped_mutual(qpool, q_off, tpool, t_off)
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        d = torch.device(dev)
        d2, idx = backend.top2(torch.from_numpy(q).to(d), torch.from_numpy(t).to(d), algo=backend.FM_ALGO_TCGEN05)
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), od2) and np.array_equal(idx.cpu().numpy(), oidx)
        g = backend.grouped_mutual(torch.from_numpy(qpool).to(d), torch.from_numpy(q_off).to(d),
                                   torch.from_numpy(tpool).to(d), torch.from_numpy(t_off).to(d))
        assert np.array_equal(g[0].cpu().numpy().view(np.uint32), gd2) and np.array_equal(g[2].cpu().numpy(), gt2q)
        before = torch.cuda.current_device()
        hd2, hidx, _ = backend.top2_host(q, t, device=d.index, want_dist=False)
        assert torch.cuda.current_device() == before                     # the caller's device is restored
        assert np.array_equal(hd2, od2) and np.array_equal(hidx, oidx)


@pytest.mark.parametrize("opts", [{}, {"grid_size": (75, 75), "grid_margin": 30, "radius": 50}])
def test_fastmatch_cuda_equals_sequential_oracle(cuda, opts):
    """Config 1 (README example): fastmatch.match on graf img4 -> img1 with the CUDA backend
    (wave-batched grouped launches) == the sequential driver with the integer oracle."""
    img1 = frozen.target_image()
    ref_cache = frozen.query_cache()             # frozen SIFT features: independent of this box's OpenCV
    o, th = ref_cache.original, ref_cache.thumb
    mc = fm_cache.Metric_Cache.from_features(th["descriptors"], th["positions"], th["size"],
                                             o["descriptors"], o["positions"], o["size"])
    assert np.array_equal(mc.original["distances"], o["distances"])
    for tau in (0.7, 0.9):
        log_a, log_b, stats = [], [], {}
        got = fastmatch.match(mc, img1, dict(opts, log=log_a, stats=stats, features=frozen.features))(tau)
        ref = fastmatch_ref.match(ref_cache, img1, dict(opts, log=log_b, features=frozen.features))
        want = ref(tau)
        assert len(got) == len(want) and len(got) > 0
        for (ia, da), (ib, db) in zip(got, want):
            assert int(ia) == int(ib) and np.array_equal(da["positions"], db["positions"]) and da["ratio"] == db["ratio"]
        assert len(log_a) == len(log_b)
        for x, y in zip(log_a, log_b):
            assert x["target_grid"] == y["target_grid"] and np.array_equal(x["matches"], y["matches"])
            assert np.array_equal(x["ratios"], y["ratios"]) and x["radius"] == y["radius"] and x["margin"] == y["margin"]
        assert stats["launches"] < ref.rounds
    gold = np.load(os.path.join(GOLD, "fastmatch_graf41.npz"))
    h = gold["query_desc_hash"]
    assert (int(o["descriptors"].astype(np.uint64).sum()), len(o["descriptors"])) == (int(h[0]), int(h[1]))
    if not opts:     # the frozen run with cv2.BFMatcher as the matcher (tests/golden/make_golden.py)
        for tau, key in ((0.7, "tau70"), (0.9, "tau90")):
            ms = fastmatch.match(mc, img1, {"features": frozen.features})(tau)
            assert np.array_equal(np.array([m[0] for m in ms], np.int64), gold[key + "_index"])
            assert np.array_equal(np.array([m[1]["positions"] for m in ms]).reshape(-1, 2, 2), gold[key + "_pos"])
            assert np.array_equal(np.array([m[1]["ratio"] for m in ms]), gold[key + "_ratio"])


def test_match_many_on_cuda_equals_single_pairs(cuda):
    """fastmatch.match_many: three pairs (two cache objects) through shared grouped launches ==
    the single-pair driver == the sequential oracle driver."""
    img1 = frozen.target_image()
    ref_cache = frozen.query_cache()
    o, th = ref_cache.original, ref_cache.thumb
    mk = lambda: fm_cache.Metric_Cache.from_features(th["descriptors"], th["positions"], th["size"],
                                                     o["descriptors"], o["positions"], o["size"])
    a, b = mk(), mk()
    assert hasattr(a, "get_indices_many")
    pts = [(400, 300), (10, 10), (640, 500)]
    for got, (x, y) in zip(a.get_indices_many(pts, 100), pts):
        assert np.array_equal(got, a.get_indices(x, y, 100))
    for tau in (0.7, 0.9):
        stats = {}
        many = fastmatch.match_many([a, b, a], [img1, img1, img1], {"features": frozen.features, "stats": stats})(tau)
        want = fastmatch_ref.match(ref_cache, img1, {"features": frozen.features})(tau)
        for ms in many:
            assert len(ms) == len(want) > 0
            for (ia, da), (ib, db) in zip(ms, want):
                assert int(ia) == int(ib) and np.array_equal(da["positions"], db["positions"]) and da["ratio"] == db["ratio"]
        assert stats["rounds_evaluated"] >= 10 * stats["launches"]
    # a second tau on the same closure reuses the memoised rounds and the resident pools
    stats = {}
    gm = fastmatch.match(a, img1, {"features": frozen.features, "stats": stats})
    first = gm(0.9)
    evaluated = stats["rounds_evaluated"]
    again = gm(0.9)
    assert stats["rounds_evaluated"] == evaluated and len(again) == len(first)


def test_target_sharding_on_one_device_equals_unsharded(cuda):
    """Config 5 at reduced size: S shards processed with t_index_base + merge == one launch."""
    q, t = synth.make_pair(20000, 30001, seed=1239)
    qd, td = torch.from_numpy(q).to(cuda), torch.from_numpy(t).to(cuda)
    d2, idx = backend.top2(qd, td)
    for S in (2, 4, 8):
        keys = []
        for s in range(S):
            lo, hi = sharded.shard_range(len(t), s, S)
            keys.append(backend.top2(qd, td[lo:hi].contiguous(), t_index_base=lo, want_keys=True)[2])
        _, md2, midx = backend.merge_top2(torch.stack(keys))
        assert torch.equal(md2, d2) and torch.equal(midx, idx)
    od2, oidx = oracle.c_top2(q[:2000], t)
    assert np.array_equal(d2[:2000].cpu().numpy().view(np.uint32), od2) and np.array_equal(idx[:2000].cpu().numpy(), oidx)


def test_full_size_properties(cuda):
    """BASELINE-size configs through size-independent properties (the oracle would take too long):
    c3 50k x 50k and c4 10k groups."""
    q, t = synth.make_pair(50000, 50000, seed=1237)
    qd, td = torch.from_numpy(q).to(cuda), torch.from_numpy(t).to(cuda)
    d2, idx = backend.top2(qd, td)
    # (1) two independent kernels agree bit for bit
    d2m, idxm = backend.top2(qd, td, algo=backend.FM_ALGO_MMA_SYNC)
    assert torch.equal(d2, d2m) and torch.equal(idx, idxm)
    # (2) reported distances are the true distances of the reported indices; ascending; ties -> lower index
    sel = torch.randperm(50000, device=cuda)[:4096]
    qq = qd[sel].long()
    for c in range(2):
        tt = td[idx[sel, c].long()].long()
        assert torch.equal(((qq - tt) ** 2).sum(1).int(), d2[sel, c])
    assert (d2[:, 0] <= d2[:, 1]).all()
    tie = d2[:, 0] == d2[:, 1]
    assert (idx[tie, 0] < idx[tie, 1]).all()
    # (3) oracle on a sample of queries
    s = sel[:512].cpu().numpy()
    od2, oidx = oracle.c_top2(q[s], t)
    assert np.array_equal(d2[s].cpu().numpy().view(np.uint32), od2) and np.array_equal(idx[s].cpu().numpy(), oidx)
    # (4) self-match: every row finds itself (or a lower-indexed exact duplicate) at distance 0
    sd2, sidx = backend.top2(td, td)
    assert (sd2[:, 0] == 0).all() and (sidx[:, 0] <= torch.arange(50000, device=cuda)).all()
    # c4: grouped launch at full size, a sample of groups against the oracle
    qpool, q_off, tpool, t_off = synth.make_groups(2000, 32, 512, seed=1238)
    g = backend.grouped_mutual(torch.from_numpy(qpool).to(cuda), torch.from_numpy(q_off).to(cuda),
                               torch.from_numpy(tpool).to(cuda), torch.from_numpy(t_off).to(cuda))
    od2, oidx, ot2q = oracle.c_grouped_mutual(qpool, q_off, tpool, t_off)
    assert np.array_equal(g[0].cpu().numpy().view(np.uint32), od2) and np.array_equal(g[1].cpu().numpy(), oidx)
    assert np.array_equal(g[2].cpu().numpy(), ot2q)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_nccl(cuda, tmp_path):
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(__file__), "_sharded_nccl_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout


def test_geometric_evaluation_with_shipped_homography(cuda):
    """inlier fraction under graf's H1to4p for Ratio-Match (CUDA matcher) equals
    the value computed from the frozen cv2 matches; Fast-Match's fraction is reported alongside."""
    import cv2
    from fast_match_b200 import evaluate
    H = evaluate.load_homography(os.path.join(GOLD, "graf_H1to4p.txt"))
    d, idx = G["graf41_knn2_dist"].astype(np.float64), G["graf41_knn2_idx"]
    for tau, want_n in ((0.6, 24), (0.7, 66), (0.8, 281)):
        pos, ratios = evaluate.ratio_match_positions(G["graf4_desc"], G["graf4_pos"], G["graf1_desc"], G["graf1_pos"], tau)
        keep = np.nonzero(d[:, 0] / d[:, 1] < tau)[0]
        ref = np.stack([G["graf4_pos"][keep], G["graf1_pos"][idx[keep, 0]]], 1).astype(np.float64)
        assert len(pos) == want_n and (np.diff(ratios) >= 0).all()
        assert evaluate.inlier_fraction(pos, H)[:2] == evaluate.inlier_fraction(ref, H)[:2]
    img1 = cv2.imread(os.path.join(GOLD, "graf1.png"))
    mc = fm_cache.Metric_Cache(os.path.join(GOLD, "graf4.png"), {"save": False})
    ms = fastmatch.match(mc, img1, {})(0.7)
    ok, n, frac = evaluate.inlier_fraction(ms, H)
    assert n == len(ms) > 0 and 0.0 <= frac <= 1.0
