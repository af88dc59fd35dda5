"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle, bit for bit."""
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


def _check_top2(q, t, cuda, algo, base=0):
    d2, idx, keys = backend.top2(_dev(q, cuda), _dev(t, cuda), t_index_base=base, algo=algo, want_keys=True)
    od2, oidx = oracle.c_top2(q, t, base)
    assert np.array_equal(_u32(d2), od2)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.pack_keys(od2, oidx))


ALGOS = [backend.FM_ALGO_MMA_SYNC, backend.FM_ALGO_TCGEN05]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("M,N", [(1, 1), (1, 2), (3, 1), (5, 0), (16, 8), (127, 63), (128, 64), (129, 65),
                                 (255, 257), (256, 256), (257, 255), (300, 1000), (1000, 300),
                                 (513, 2049)])
def test_top2_shapes(cuda, algo, M, N):
    if algo == backend.FM_ALGO_TCGEN05 and not backend.device_caps()["has_tcgen05"]:
        pytest.skip("no tcgen05")
    q, t = synth.make_pair(M, max(N, 1), seed=M * 7919 + N)
    _check_top2(q, t[:N], cuda, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_top2_ties_and_extremes(cuda, algo):
    rng = np.random.default_rng(5)
    t = synth.siftlike(700, rng)
    t[100:200] = t[300:400]          # exact duplicates -> ties, lowest index must win
    t[500] = 0
    t[501] = 255
    t[502] = 255
    q = np.concatenate([t[300:420], np.zeros((3, 128), np.uint8), np.full((3, 128), 255, np.uint8),
                        synth.siftlike(200, rng)])
    _check_top2(q, t, cuda, algo)
    _check_top2(q, t, cuda, algo, base=1 << 20)
    # all rows identical: every distance ties at 0
    same = np.tile(t[:1], (300, 1))
    _check_top2(same[:130], same, cuda, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_top2_uniform_random_full_range(cuda, algo):
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (700, 128), dtype=np.uint8)
    t = rng.integers(0, 256, (900, 128), dtype=np.uint8)
    q[0] = 0; t[0] = 255      # d2 = 128*255^2, the largest possible
    _check_top2(q, t, cuda, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_top2_medium(cuda, algo):
    q, t = synth.make_pair(5000, 5000, seed=1236)
    _check_top2(q, t, cuda, algo)
    _check_top2(q, q, cuda, algo)   # self-match (Metric_Cache): slot 0 is the row itself or a lower duplicate


def test_top2_auto_and_host(cuda):
    q, t = synth.make_pair(3000, 4100, seed=3)
    d2, idx = backend.top2(_dev(q, cuda), _dev(t, cuda))
    od2, oidx = oracle.c_top2(q, t)
    assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx)
    hd2, hidx, hdist = backend.top2_host(q, t)
    assert np.array_equal(hd2, od2) and np.array_equal(hidx, oidx)
    assert np.array_equal(hdist, np.sqrt(od2.astype(np.float32)))


def test_mutual_single_big_pair_equals_grouped_and_oracle(cuda):
    """crossCheck for one big pair (two dense launches) == the grouped kernel with G = 1 == the oracle,
    including exact duplicates on both sides (ties to the lowest index in both directions)."""
    q, t = synth.make_pair(2500, 1100, seed=21)
    t[5:15] = t[700:710]
    q[100:110] = q[3:13]
    qd, td = _dev(q, cuda), _dev(t, cuda)
    d2, idx, mutual = backend.mutual_single(qd, td)
    off = torch.tensor([[0, len(q)], [0, len(t)]], dtype=torch.int64, device=cuda)
    gd2, gidx, _, gmut = backend.grouped_mutual(qd, off[0], td, off[1])
    assert torch.equal(d2, gd2[:, 0]) and torch.equal(idx, gidx[:, 0]) and torch.equal(mutual, gmut)
    od2, oidx, ot2q = oracle.np_mutual(q, t)
    keep = oracle.mutual_pairs(oidx, ot2q)
    assert np.array_equal(np.nonzero(mutual.cpu().numpy())[0], keep)
    assert np.array_equal(_u32(d2), od2[:, 0]) and np.array_equal(idx.cpu().numpy(), oidx[:, 0])
    # tiny pair: stays on the grouped kernel
    s2, si, sm = backend.mutual_single(qd[:40], td[:30])
    sd2, sidx, st2q = oracle.np_mutual(q[:40], t[:30])
    assert np.array_equal(np.nonzero(sm.cpu().numpy())[0], oracle.mutual_pairs(sidx, st2q)) and np.array_equal(si.cpu().numpy(), sidx[:, 0])
    e = backend.mutual_single(qd[:7], td[:0])
    assert not e[2].any() and (e[1] == -1).all()


def test_host_entry_point_two_halves(cuda):
    """fm_top2_host_u8 on a problem large enough for its overlapped form (queries go up in two halves,
    the second while the first is matched): pageable and pinned buffers, distances + ratio mask, an odd
    row count, and a repeat call on the same context -- all equal to the device path and the oracle."""
    M, N = 20011, 6007
    q, t = synth.make_pair(M, N, seed=12)
    d2, idx, ratio, mask = backend.ratio_match(_dev(q, cuda), _dev(t, cuda), 0.8, want_ratio=True)
    want_d2, want_idx, want_mask = _u32(d2), idx.cpu().numpy(), mask.cpu().numpy()
    rows = np.r_[0:300, 9900:10400, M - 300:M]
    od2, oidx = oracle.c_top2(q[rows], t)
    assert np.array_equal(want_d2[rows], od2) and np.array_equal(want_idx[rows], oidx)
    for _ in range(2):
        hd2, hidx, hdist, hmask = backend.top2_host(q, t, want_dist=True, tau=0.8)
        assert np.array_equal(hd2, want_d2) and np.array_equal(hidx, want_idx)
        assert np.array_equal(hdist, np.sqrt(want_d2.astype(np.float32))) and np.array_equal(hmask.astype(bool), want_mask)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    out = (pin(np.empty((M, 2), np.uint32)), pin(np.empty((M, 2), np.int32)), None, pin(np.empty(M, np.uint8)))
    pd2, pidx, _, pmask = backend.top2_host(pin(q), pin(t), want_dist=False, tau=0.8, out=out)
    assert np.array_equal(pd2, want_d2) and np.array_equal(pidx, want_idx) and np.array_equal(pmask.astype(bool), want_mask)
    # no targets at all: every slot missing, mask 0
    ed2, eidx, _, emask = backend.top2_host(q[:100], t[:0], want_dist=False, tau=0.8)
    assert (ed2 == 0xFFFFFFFF).all() and (eidx == -1).all() and not emask.any()


def test_ratio(cuda):
    q, t = synth.make_pair(4000, 4000, seed=4)
    d2, idx = backend.top2(_dev(q, cuda), _dev(t, cuda))
    for tau in (0.6, 0.7, 0.8, 0.9, 1.0):
        r, m = backend.ratio(d2[:, 0], den_d2=d2[:, 1], tau=tau)
        orr, om = oracle.np_ratio(_u32(d2)[:, 0], den_d2=_u32(d2)[:, 1], tau=tau)
        assert np.array_equal(r.cpu().numpy(), orr, equal_nan=True)
        assert np.array_equal(m.cpu().numpy(), om)
    den = torch.rand(4000, device=cuda) * 300 + 1
    r, m = backend.ratio(d2[:, 0], den_f32=den, tau=0.7)
    orr, om = oracle.c_ratio(_u32(d2)[:, 0].copy(), den_f32=den.cpu().numpy(), tau=0.7)
    assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(m.cpu().numpy(), om)
    # missing second neighbour -> +inf, mask 0
    d2b, _ = backend.top2(_dev(q[:10], cuda), _dev(t[:1], cuda))
    r, m = backend.ratio(d2b[:, 0], den_d2=d2b[:, 1], tau=0.7)
    assert torch.isinf(r).all() and not m.any()


@pytest.mark.parametrize("algo", ALGOS)
def test_ratio_match_fused(cuda, algo):
    """fm_ratio_match_u8: top-2 with the ratio test fused into the kernel write-out."""
    for (M, N, seed) in ((4000, 4100, 4), (20000, 30000, 5), (700, 1, 6), (300, 0, 7)):
        q, t = synth.make_pair(M, max(N, 1), seed=seed)
        t = t[:N]
        od2, oidx = oracle.c_top2(q, t)
        for tau in (0.7, 0.9):
            d2, idx, r, m = backend.ratio_match(_dev(q, cuda), _dev(t, cuda), tau, algo=algo, want_ratio=True)
            orr, om = oracle.np_ratio(od2[:, 0], den_d2=od2[:, 1], tau=tau)
            assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx)
            assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(m.cpu().numpy(), om)


def _check_grouped(qpool, q_off, tpool, t_off, cuda, q_gather=None):
    for algo in ALGOS:
        _check_grouped_algo(qpool, q_off, tpool, t_off, cuda, q_gather, algo)


def _check_grouped_algo(qpool, q_off, tpool, t_off, cuda, q_gather, algo):
    args = dict(q_gather=None if q_gather is None else _dev(q_gather.astype(np.int32), cuda), algo=algo)
    d2, idx, t2q, mutual = backend.grouped_mutual(_dev(qpool, cuda), _dev(q_off, cuda), _dev(tpool, cuda),
                                                  _dev(t_off, cuda), **args)
    od2, oidx, ot2q = oracle.c_grouped_mutual(qpool, q_off, tpool, t_off, q_gather=q_gather)
    assert np.array_equal(_u32(d2), od2)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(t2q.cpu().numpy(), ot2q)
    om = np.zeros(len(od2), bool)
    for g in range(len(q_off) - 1):
        sl = slice(q_off[g], q_off[g + 1])
        keep = oracle.mutual_pairs(oidx[sl], ot2q[t_off[g]:t_off[g + 1]])
        om[q_off[g] + keep] = True
    assert np.array_equal(mutual.cpu().numpy(), om)


def test_grouped_small_and_ragged(cuda):
    qpool, q_off, tpool, t_off = synth.make_groups(200, 1, 300, seed=21)
    _check_grouped(qpool, q_off, tpool, t_off, cuda)
    # empty groups and empty sides
    rng = np.random.default_rng(3)
    nq = np.array([0, 5, 0, 130, 1, 0, 64, 700])
    nt = np.array([4, 0, 0, 1, 1, 9, 64, 70])
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    qpool = synth.siftlike(int(q_off[-1]), rng)
    tpool = synth.siftlike(int(t_off[-1]), rng)
    tpool[-60:-30] = tpool[-30:]      # ties on the target side
    qpool[-300:-200] = qpool[-100:]   # ties on the query side (column argmin -> lowest local row)
    _check_grouped(qpool, q_off, tpool, t_off, cuda)


def test_grouped_gather(cuda):
    rng = np.random.default_rng(8)
    pool = synth.siftlike(5000, rng)
    nq = rng.integers(20, 400, 150)
    nt = rng.integers(5, 130, 150)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    gather = rng.integers(0, 5000, int(q_off[-1]))
    tpool = pool[rng.integers(0, 5000, int(t_off[-1]))].copy()
    noise = rng.integers(-6, 7, tpool.shape)
    tpool = np.clip(tpool.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    _check_grouped(pool, q_off, tpool, t_off, cuda, q_gather=gather)


def test_grouped_shared_cells_t_base(cuda):
    """Many groups referencing the same resident cells through t_base (the flood-fill layout),
    including groups larger than one unit (nq > 128, nt > 256) and more than 512 rows a side."""
    rng = np.random.default_rng(12)
    pool = synth.siftlike(3000, rng)
    cells = synth.siftlike(2500, rng)
    cell_start = np.array([0, 40, 41, 300, 900, 1700])
    cell_cnt = np.array([40, 1, 259, 600, 800, 800])
    which = rng.integers(0, 6, 60)
    nq = rng.integers(1, 700, 60)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(cell_cnt[which])]).astype(np.int64)
    gather = rng.integers(0, 3000, int(q_off[-1])).astype(np.int32)
    t_base = cell_start[which].astype(np.int64)
    tcat = np.concatenate([cells[s:s + c] for s, c in zip(t_base, cell_cnt[which])])
    od2, oidx, ot2q = oracle.c_grouped_mutual(pool, q_off, tcat, t_off, q_gather=gather)
    for algo in ALGOS:
        d2, idx, t2q, _ = backend.grouped_mutual(_dev(pool, cuda), _dev(q_off, cuda), _dev(cells, cuda), _dev(t_off, cuda),
                                                 q_gather=_dev(gather, cuda), t_base=_dev(t_base, cuda), algo=algo)
        assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx)
        assert np.array_equal(t2q.cpu().numpy(), ot2q)


def test_grouped_config4_sample(cuda):
    qpool, q_off, tpool, t_off = synth.make_groups(300, 32, 512, seed=1238)
    _check_grouped(qpool, q_off, tpool, t_off, cuda)


def test_merge_top2(cuda):
    q, t = synth.make_pair(3000, 4000, seed=9)
    S = 4
    keys = []
    for s in range(S):
        lo, hi = s * 1000, (s + 1) * 1000
        _, _, k = backend.top2(_dev(q, cuda), _dev(t[lo:hi], cuda), t_index_base=lo, want_keys=True)
        keys.append(k)
    keys = torch.stack(keys)
    out, d2, idx = backend.merge_top2(keys)
    od2, oidx = oracle.c_top2(q, t)
    assert np.array_equal(_u32(d2), od2) and np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(out.cpu().numpy().view(np.uint64), oracle.c_merge_top2(keys.cpu().numpy().view(np.uint64)))


def test_errors_are_loud(cuda):
    q = torch.zeros((4, 128), dtype=torch.uint8)          # CPU tensor
    with pytest.raises(backend.FastMatchError):
        backend.top2(q, q)
    with pytest.raises(backend.FastMatchError):
        backend.top2(torch.zeros((4, 64), dtype=torch.uint8, device=cuda), torch.zeros((4, 128), dtype=torch.uint8, device=cuda))
