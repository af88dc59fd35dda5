"""Host logic of the dense kernel's persistent schedule (no GPU needed): the plan the library
reports through its test hook is replayed in Python -- every worker's runs of tiles cover the
(M-block, tile) steps exactly once, slot indices stay inside the partial-key buffer the plan sizes,
and the number of slots the merge reads per M-block equals the number of workers that wrote one."""
import bisect
import ctypes

import pytest

from fast_match_b200 import backend


def owner_of_step(begins, step):
    """The worker whose range [begins[w], begins[w+1]) holds `step` (owner_of_step in fm_tc.cu)."""
    return bisect.bisect_right(begins, step) - 1


def plan_begins(M, N, workers):
    out = (ctypes.c_longlong * (workers + 1))()
    L = backend.lib()
    L.fm_debug_plan_begins.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    assert L.fm_debug_plan_begins(M, N, out, workers + 1) == 0
    return list(out)


def plan(M, N):
    out = (ctypes.c_longlong * 8)()
    L = backend.lib()
    L.fm_debug_plan.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_longlong)]
    assert L.fm_debug_plan(M, N, out) == 0
    keys = ("pair", "mblock_rows", "mblocks", "ntiles", "workers", "slots", "aligned", "ws_kib")
    return dict(zip(keys, list(out)))


SHAPES = [(1, 1), (1, 300000), (3, 70001), (255, 255), (257, 513), (513, 255), (777, 40000), (40000, 777),
          (3000, 9000), (50000, 50000), (70000, 257), (300000, 300), (200000, 200000), (1000000, 125000),
          (1000000, 1000000)]


@pytest.mark.parametrize("M,N", SHAPES)
def test_stream_k_partition(M, N):
    check_plan(M, N)


def test_stream_k_partition_random_shapes():
    """The same invariants on seeded random shapes (log-uniform sizes up to 2M x 2M)."""
    import random
    rng = random.Random(2024)
    for _ in range(150):
        M = int(2 ** rng.uniform(0, 21))
        N = int(2 ** rng.uniform(0, 21))
        check_plan(M, N)


@pytest.mark.gpu
def test_stream_k_partition_on_the_device(cuda):
    """On the GPU box the plan uses the CTA-pair kernel (512-row M-blocks, 74 workers): same invariants."""
    import random
    rng = random.Random(7)
    assert plan(50000, 50000)["pair"] == 1 and plan(50000, 50000)["mblock_rows"] == 512
    for M, N in SHAPES + [(int(2 ** rng.uniform(8, 21)), int(2 ** rng.uniform(8, 21))) for _ in range(60)]:
        check_plan(M, N)


def check_plan(M, N):
    p = plan(M, N)
    rows, mblocks, ntiles, W = p["mblock_rows"], p["mblocks"], p["ntiles"], p["workers"]
    assert rows in (256, 512) and mblocks == -(-M // rows) and ntiles == -(-N // 256)
    work = mblocks * ntiles
    assert 1 <= W <= work
    covered = 0
    writers = {}
    begins = plan_begins(M, N, W)
    assert begins[0] == 0 and begins[W] == work and all(b > a for a, b in zip(begins, begins[1:]))
    sizes = [b - a for a, b in zip(begins, begins[1:])]
    assert max(sizes) <= 1.25 * work / W + 32            # the cost model only nudges an equal split
    for w in range(W):
        lo, hi = begins[w], begins[w + 1]
        if p["aligned"]:
            assert lo % ntiles == 0 and hi % ntiles == 0
        assert hi - lo < 2 ** 31
        step = lo
        first = True
        while step < hi:
            m, tb = divmod(step, ntiles)
            nt = min(hi - step, ntiles - tb)
            slot = 0 if tb == 0 else w - owner_of_step(begins, step - tb)
            assert first or tb == 0                      # only the first run starts inside an M-block
            assert 0 <= slot < p["slots"], (w, m, slot)
            assert slot not in writers.setdefault(m, set())
            writers[m].add(slot)
            covered += nt
            step += nt
            first = False
    assert covered == work
    for m in ([0, mblocks - 1] if mblocks > 4000 else range(mblocks)):
        f = m * ntiles
        n = owner_of_step(begins, f + ntiles - 1) - owner_of_step(begins, f) + 1
        assert writers[m] == set(range(n)), (m, writers[m], n)   # what k_merge_partial reads
    assert p["ws_kib"] * 1024 >= p["slots"] * M * 16


def test_workspace_query_matches_plan():
    L = backend.lib()
    for M, N in SHAPES[:12]:
        assert L.fm_top2_workspace_bytes(M, N) >> 10 == plan(M, N)["ws_kib"]
