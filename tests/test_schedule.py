"""Host logic of the dense kernel's persistent schedule (no GPU needed): the plan the library
reports through its test hook is replayed in Python -- every worker's runs of tiles cover the
(M-block, tile) steps exactly once, slot indices stay inside the partial-key buffer the plan sizes,
and the number of slots the merge reads per M-block equals the number of workers that wrote one."""
import ctypes

import pytest

from fast_match_b200 import backend


def cta_of_step(step, work, grid):
    """First worker whose range [work*c//grid, work*(c+1)//grid) holds `step` (fm_tc.cu)."""
    c = step * grid // work
    while work * (c + 1) // grid <= step:
        c += 1
    while c > 0 and work * c // grid > step:
        c -= 1
    return c


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
    p = plan(M, N)
    rows, mblocks, ntiles, W = p["mblock_rows"], p["mblocks"], p["ntiles"], p["workers"]
    assert rows in (256, 512) and mblocks == -(-M // rows) and ntiles == -(-N // 256)
    work = mblocks * ntiles
    assert 1 <= W <= work
    covered = 0
    writers = {}
    for w in range(W):
        if p["aligned"]:
            lo, hi = mblocks * w // W * ntiles, mblocks * (w + 1) // W * ntiles
        else:
            lo, hi = work * w // W, work * (w + 1) // W
        assert hi - lo < 2 ** 31
        step = lo
        first = True
        while step < hi:
            m, tb = divmod(step, ntiles)
            nt = min(hi - step, ntiles - tb)
            slot = 0 if tb == 0 else w - cta_of_step(step - tb, work, W)
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
        n = 1 if p["aligned"] else cta_of_step(f + ntiles - 1, work, W) - cta_of_step(f, work, W) + 1
        assert writers[m] == set(range(n)), (m, writers[m], n)   # what k_merge_partial reads
    assert p["ws_kib"] * 1024 >= p["slots"] * M * 16


def test_workspace_query_matches_plan():
    L = backend.lib()
    for M, N in SHAPES[:12]:
        assert L.fm_top2_workspace_bytes(M, N) >> 10 == plan(M, N)["ws_kib"]
