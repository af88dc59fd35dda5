"""CPU, world_size 2 over gloo: the host logic of the target-sharded path (shard ranges, one
all-gather of packed keys, merge) with oracle stand-ins for the two CUDA calls."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fast_match_b200 import sharded, synth


def test_shard_ranges_partition_the_rows():
    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            r = [sharded.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def _worker(rank, world, port, q, t, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharded.shard_range(len(t), rank, world)

        def local_top2(qq, ts, base):
            d2, idx = oracle.c_top2(qq.numpy(), ts.numpy(), base)
            return torch.from_numpy(oracle.pack_keys(d2, idx).view(np.int64))

        def merge(g):
            return torch.from_numpy(oracle.c_merge_top2(g.numpy().view(np.uint64)).view(np.int64))

        keys = sharded.sharded_top2(torch.from_numpy(q), torch.from_numpy(t[lo:hi]), lo,
                                    local_top2=local_top2, merge=merge)
        out[rank] = keys.numpy().view(np.uint64).copy()
        # the all-to-all exchange: this rank merges only its slice of the queries
        (q_lo, q_hi), sl = sharded.sharded_top2_sliced(torch.from_numpy(q), torch.from_numpy(t[lo:hi]), lo,
                                                       local_top2=local_top2, merge=merge)
        assert (q_lo, q_hi) == sharded.shard_range(len(q), rank, world)
        out[("slice", rank)] = (q_lo, q_hi, sl.numpy().view(np.uint64).copy())
        full = sharded.gather_full(sl, len(q))
        out[("full", rank)] = full.numpy().view(np.uint64).copy()
        flags = sharded.gather_full(torch.arange(q_lo, q_hi) % 3 == 0, len(q))
        assert flags.dtype == torch.bool and torch.equal(flags, torch.arange(len(q)) % 3 == 0)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_top2_equals_unsharded():
    q, t = synth.make_pair(701, 901, seed=5)          # odd: the two query slices differ in length
    t[10:20] = t[600:610]                      # ties that straddle the shard boundary
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, q, t, out), nprocs=2, join=True)
    d2, idx = oracle.c_top2(q, t)
    want = oracle.pack_keys(d2, idx)
    assert np.array_equal(out[0], want) and np.array_equal(out[1], want)
    for r in (0, 1):
        q_lo, q_hi, sl = out[("slice", r)]
        assert np.array_equal(sl, want[q_lo:q_hi])
        assert np.array_equal(out[("full", r)], want)


def _worker_2d(rank, world, port, q, t, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def local_top2(qq, ts, base):
            d2, idx = oracle.c_top2(qq.numpy(), ts.numpy(), base)
            return torch.from_numpy(oracle.pack_keys(d2, idx).view(np.int64))

        def merge(g):
            return torch.from_numpy(oracle.c_merge_top2(g.numpy().view(np.uint64)).view(np.int64))

        for sq in (1, 2, 4):
            grid = sharded.Grid2D(query_groups=sq, backend="gloo")
            lo, hi = grid.target_range(len(t))
            rows, merged = grid.top2_sliced(torch.from_numpy(q), torch.from_numpy(t[lo:hi]), len(t),
                                            local_top2=local_top2, merge=merge)
            assert rows == grid.my_rows(len(q))
            out[(sq, rank)] = (rows, merged.numpy().view(np.uint64).copy())
    finally:
        dist.destroy_process_group()


def test_four_rank_2d_grid_equals_unsharded():
    """world 4 over gloo as 1 x 4 (plain target sharding), 2 x 2 and 4 x 1 (query groups x target
    shards): every arrangement tiles the query rows exactly once with the unsharded answer."""
    q, t = synth.make_pair(403, 610, seed=8)
    t[3:9] = t[300:306]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_2d, args=(4, port, q, t, out), nprocs=4, join=True)
    d2, idx = oracle.c_top2(q, t)
    want = oracle.pack_keys(d2, idx)
    for sq in (1, 2, 4):
        covered = np.zeros(len(q), bool)
        for r in range(4):
            (lo, hi), got = out[(sq, r)]
            assert np.array_equal(got, want[lo:hi]), (sq, r)
            assert not covered[lo:hi].any()
            covered[lo:hi] = True
        assert covered.all()
