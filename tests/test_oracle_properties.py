"""Property tests of the checker itself (CPU): the C restatement and the numpy restatement of the
matcher must agree on arbitrary small u8 inputs -- including the degenerate ones the reference's
matcher meets (duplicates, constant rows, one or two targets) -- and sharding + merge must be
transparent.  The GPU parity tests rely on these two being the same function."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


def _desc(draw, rows, palette):
    # few distinct rows -> many exact ties; values from a small palette incl. the extremes
    base = draw(st.lists(st.lists(st.sampled_from(palette), min_size=128, max_size=128), min_size=1, max_size=4))
    pick = draw(st.lists(st.integers(0, len(base) - 1), min_size=rows, max_size=rows))
    a = np.array([base[i] for i in pick], dtype=np.uint8)
    flips = draw(st.lists(st.tuples(st.integers(0, rows - 1), st.integers(0, 127), st.sampled_from(palette)),
                          max_size=rows))
    for r, c, v in flips:
        a[r, c] = v
    return a


@st.composite
def problems(draw):
    palette = draw(st.sampled_from([[0, 255], [0, 1, 2, 255], [7, 8, 9, 120, 121], list(range(0, 256, 17))]))
    m = draw(st.integers(1, 9))
    n = draw(st.integers(1, 12))
    return _desc(draw, m, palette), _desc(draw, n, palette), draw(st.integers(0, 1000))


@settings(max_examples=80, deadline=None)
@given(problems())
def test_c_and_numpy_top2_agree(p):
    q, t, base = p
    d2a, ia = oracle.c_top2(q, t, base)
    d2b, ib = oracle.np_top2(q, t, base)
    assert np.array_equal(d2a, d2b) and np.array_equal(ia, ib)
    # lexicographic (d2, index) order, ties to the lowest index, missing second slot for one target
    d2 = oracle.np_d2(q, t)
    for i in range(q.shape[0]):
        order = sorted(range(t.shape[0]), key=lambda j: (int(d2[i, j]), j))
        assert ia[i, 0] == order[0] + base and d2a[i, 0] == d2[i, order[0]]
        if t.shape[0] > 1:
            assert ia[i, 1] == order[1] + base and d2a[i, 1] == d2[i, order[1]]
        else:
            assert ia[i, 1] == -1 and d2a[i, 1] == 0xFFFFFFFF


@settings(max_examples=60, deadline=None)
@given(problems(), st.integers(1, 4))
def test_sharding_and_merge_are_transparent(p, shards):
    q, t, _ = p
    n = t.shape[0]
    cuts = [n * s // shards for s in range(shards + 1)]
    keys = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        if hi > lo:
            d2, idx = oracle.c_top2(q, t[lo:hi], lo)
        else:       # an empty shard contributes "no candidate"
            d2 = np.full((q.shape[0], 2), 0xFFFFFFFF, np.uint32)
            idx = np.full((q.shape[0], 2), -1, np.int32)
        keys.append(oracle.pack_keys(d2, idx))
    keys = np.stack(keys)
    want = oracle.pack_keys(*oracle.c_top2(q, t))
    for merged in (oracle.c_merge_top2(keys), oracle.np_merge_top2(keys)):
        assert np.array_equal(merged, want)


@settings(max_examples=60, deadline=None)
@given(problems())
def test_mutual_is_symmetric(p):
    """crossCheck pairs are the same set whichever side is called the query."""
    q, t, _ = p
    _, q2t, t2q = oracle.np_mutual(q, t)
    kept = oracle.mutual_pairs(q2t, t2q)
    pairs = sorted((int(i), int(q2t[i, 0])) for i in kept)
    _, t2q_b, q2t_b = oracle.np_mutual(t, q)
    kept_b = oracle.mutual_pairs(t2q_b, q2t_b)
    pairs_b = sorted((int(t2q_b[j, 0]), int(j)) for j in kept_b)
    assert pairs == pairs_b
    # and the C restatement of one grouped round agrees with numpy
    off = lambda n: np.array([0, n], np.int64)
    d2c, idxc, t2qc = oracle.c_grouped_mutual(q, off(len(q)), t, off(len(t)))
    d2n, idxn, _ = oracle.np_mutual(q, t)
    assert np.array_equal(d2c, d2n) and np.array_equal(idxc, idxn) and np.array_equal(t2qc, t2q)
