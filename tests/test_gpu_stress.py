"""GPU: unusual shapes for every kernel (checked against the C oracle on samples where the full
oracle would be slow)."""
import numpy as np
import pytest
import torch

import oracle
from fast_match_b200 import backend, synth

pytestmark = pytest.mark.gpu


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("M,N", [(1, 300000), (3, 70001), (300000, 300), (70000, 257), (2, 256), (513, 255),
                                 (40000, 777), (777, 40000)])
def test_top2_extreme_aspect_ratios(cuda, M, N):
    q, t = synth.make_pair(M, N, seed=M + N)
    qd, td = _dev(q, cuda), _dev(t, cuda)
    sel = np.unique(np.linspace(0, M - 1, min(M, 300)).astype(np.int64))
    od2, oidx = oracle.c_top2(q[sel], t)
    for algo in (backend.FM_ALGO_AUTO, backend.FM_ALGO_TCGEN05):
        d2, idx = backend.top2(qd, td, algo=algo)
        assert np.array_equal(_u32(d2)[sel], od2) and np.array_equal(idx.cpu().numpy()[sel], oidx)


def test_top2_index_base_near_int32_limit(cuda):
    q, t = synth.make_pair(1000, 5000, seed=8)
    base = (1 << 31) - 1 - 5000
    d2, idx, keys = backend.top2(_dev(q, cuda), _dev(t, cuda), t_index_base=base, want_keys=True)
    od2, oidx = oracle.c_top2(q, t, base)
    assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.pack_keys(od2, oidx))
    with pytest.raises(backend.FastMatchError):
        backend.top2(_dev(q, cuda), _dev(t, cuda), t_index_base=base + 10)


def test_grouped_large_and_tiny_groups(cuda):
    rng = np.random.default_rng(4)
    # a few very large groups (several slabs and chunks on both sides) ...
    nq = np.array([5000, 1, 3000, 129, 0, 2])
    nt = np.array([3000, 4000, 1, 257, 5, 0])
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    qp, tp = synth.siftlike(int(q_off[-1]), rng), synth.siftlike(int(t_off[-1]), rng)
    tp[100:400] = tp[1000:1300]
    qp[10:60] = qp[2000:2050]
    od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
    for algo in (backend.FM_ALGO_MMA_SYNC, backend.FM_ALGO_TCGEN05):
        d2, idx, t2q, _ = backend.grouped_mutual(_dev(qp, cuda), _dev(q_off, cuda), _dev(tp, cuda), _dev(t_off, cuda), algo=algo)
        assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(t2q.cpu().numpy(), ot2q)
    # ... and very many 1x1 / 2x3 groups (more groups than SMs x pipeline depth)
    G = 20000
    nq = rng.integers(1, 3, G)
    nt = rng.integers(1, 4, G)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    qp, tp = synth.siftlike(int(q_off[-1]), rng), synth.siftlike(int(t_off[-1]), rng)
    od2, oidx, ot2q = oracle.c_grouped_mutual(qp, q_off, tp, t_off)
    d2, idx, t2q, mutual = backend.grouped_mutual(_dev(qp, cuda), _dev(q_off, cuda), _dev(tp, cuda), _dev(t_off, cuda))
    assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(t2q.cpu().numpy(), ot2q)
    assert int(mutual.sum()) > 0


def test_repeated_calls_and_streams(cuda):
    """Workspace reuse across calls of different sizes and on a side stream."""
    sizes = [(3000, 9000), (257, 70000), (9000, 3000), (3000, 9000)]
    ref = {}
    for M, N in sizes:
        q, t = synth.make_pair(M, N, seed=M)
        ref[(M, N)] = (q, t) + oracle.c_top2(q[:200], t)
    side = torch.cuda.Stream(device=cuda)
    for rep in range(2):
        for M, N in sizes:
            q, t, od2, oidx = ref[(M, N)]
            qd, td = _dev(q, cuda), _dev(t, cuda)
            torch.cuda.synchronize()
            with torch.cuda.stream(side if rep else torch.cuda.current_stream(cuda)):
                d2, idx, r, m = backend.ratio_match(qd, td, 0.8)
            torch.cuda.synchronize()
            assert np.array_equal(_u32(d2)[:200], od2) and np.array_equal(idx.cpu().numpy()[:200], oidx)
