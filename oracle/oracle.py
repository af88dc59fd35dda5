"""CPU oracle -- TEST INFRASTRUCTURE ONLY (never imported by fast_match_b200).

Two restatements of the same algorithm:
  * numpy (``np_*``): blocked integer brute force, used for small/medium cases;
  * C (``c_*``): oracle.c via ctypes (OpenMP), used for larger cases and as the
    "port" CPU baseline in bench.py.
Both follow the observable behaviour of cv2.BFMatcher(NORM_L2) as called by the
reference (matchutil.py:39-43; fastmatch.pyx:122-124, 161-165; cache.pyx:250-252,
271-273; Classic Matching.ipynb cell 3).  Parity is pinned against cv2.BFMatcher
outputs frozen in tests/golden/ (the reference itself ships no golden vectors).
"""
import ctypes
import os
import subprocess

import numpy as np

NONE_D2 = np.uint32(0xFFFFFFFF)
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle.c -> oracle/liboracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        L.oracle_top2_u8.argtypes = [vp, i64, vp, i64, i32, vp, vp]
        L.oracle_mutual_u8.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp]
        L.oracle_grouped_mutual_u8.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
        L.oracle_merge_top2.argtypes = [vp, i32, i64, vp]
        L.oracle_ratio.argtypes = [vp, i64, vp, i64, vp, i64, ctypes.c_double, vp, vp]
        for f in ("oracle_top2_u8", "oracle_mutual_u8", "oracle_grouped_mutual_u8",
                  "oracle_merge_top2", "oracle_ratio"):
            getattr(L, f).restype = None
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8 and (a.ndim == 2 and a.shape[1] == 128), (a.dtype, a.shape)
    return a


# ----------------------------------------------------------------------------
# numpy restatement
# ----------------------------------------------------------------------------
def np_d2(q, t):
    """Exact squared L2 distances, int64 [M, N]."""
    q = q.astype(np.int64)
    t = t.astype(np.int64)
    qq = (q * q).sum(1)[:, None]
    tt = (t * t).sum(1)[None, :]
    return qq + tt - 2 * (q @ t.T)


def np_top2(q, t, t_index_base=0, block=2048):
    """Lexicographic (d2, idx) two smallest per query (bf_match k=2, matchutil.py:39-43)."""
    q, t = _u8(q), _u8(t)
    M, N = len(q), len(t)
    d2 = np.full((M, 2), NONE_D2, np.uint32)
    idx = np.full((M, 2), -1, np.int32)
    if N == 0 or M == 0:
        return d2, idx
    for m0 in range(0, M, block):
        D = np_d2(q[m0:m0 + block], t)
        key = (D << 32) | np.arange(N, dtype=np.int64)[None, :]
        k = min(2, N)
        part = np.sort(np.partition(key, k - 1, axis=1)[:, :k], axis=1)
        d2[m0:m0 + block, :k] = (part >> 32).astype(np.uint32)
        idx[m0:m0 + block, :k] = (part & 0xFFFFFFFF).astype(np.int32) + t_index_base
    return d2, idx


def np_mutual(q, t, q_gather=None):
    """crossCheck=True, k=1 round (fastmatch.pyx:122-123, 161-162) + query top-2."""
    q, t = _u8(q), _u8(t)
    if q_gather is not None:
        q = q[np.asarray(q_gather, np.int64)]
    nq, nt = len(q), len(t)
    d2, idx = np_top2(q, t)
    t2q = np.full(nt, -1, np.int32)
    if nq and nt:
        D = np_d2(q, t)
        key = (D << 32) | np.arange(nq, dtype=np.int64)[:, None]
        t2q = (key.min(axis=0) & 0xFFFFFFFF).astype(np.int32)
    return d2, idx, t2q


def mutual_pairs(q2t_idx, t2q_idx):
    """Indices i of local queries kept by crossCheck: nn_t(nn_q(i)) == i."""
    i0 = q2t_idx[:, 0]
    ok = i0 >= 0
    keep = np.zeros(len(i0), bool)
    keep[ok] = t2q_idx[i0[ok]] == np.nonzero(ok)[0]
    return np.nonzero(keep)[0]


def np_ratio(num_d2, den_d2=None, den_f32=None, tau=0.7):
    """ratio = float64(sqrt_f32(num)) / float64(den_f32) ; mask = ratio < tau."""
    num_d2 = np.asarray(num_d2, np.uint32)
    num = np.sqrt(num_d2.astype(np.float32)).astype(np.float64)
    missing = num_d2 == NONE_D2
    if den_f32 is not None:
        den = np.asarray(den_f32, np.float32).astype(np.float64)
    else:
        den_d2 = np.asarray(den_d2, np.uint32)
        den = np.sqrt(den_d2.astype(np.float32)).astype(np.float64)
        missing = missing | (den_d2 == NONE_D2)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = num / den
    r[missing] = np.inf
    return r, (r < tau)


def np_merge_top2(keys):
    """keys uint64 [S, M, 2] -> two smallest per query [M, 2]."""
    S, M, _ = keys.shape
    flat = np.transpose(keys, (1, 0, 2)).reshape(M, 2 * S)
    return np.sort(flat, axis=1)[:, :2].copy()


def pack_keys(d2, idx):
    """(d2 << 32 | idx) with missing slots -> all ones."""
    k = (d2.astype(np.uint64) << np.uint64(32)) | (idx.astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)
    k[d2 == NONE_D2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return k


# ----------------------------------------------------------------------------
# C restatement (oracle.c) through ctypes
# ----------------------------------------------------------------------------
def c_top2(q, t, t_index_base=0):
    q, t = _u8(q), _u8(t)
    M, N = len(q), len(t)
    d2 = np.empty((M, 2), np.uint32)
    idx = np.empty((M, 2), np.int32)
    lib().oracle_top2_u8(_p(q), M, _p(t), N, int(t_index_base), _p(d2), _p(idx))
    return d2, idx


def c_grouped_mutual(qpool, q_off, tpool, t_off, q_gather=None):
    qpool, tpool = _u8(qpool), _u8(tpool)
    q_off = np.ascontiguousarray(q_off, np.int64)
    t_off = np.ascontiguousarray(t_off, np.int64)
    G = len(q_off) - 1
    nq, nt = int(q_off[-1]), int(t_off[-1])
    if q_gather is not None:
        q_gather = np.ascontiguousarray(q_gather, np.int32)
    d2 = np.empty((nq, 2), np.uint32)
    idx = np.empty((nq, 2), np.int32)
    t2q = np.empty(nt, np.int32)
    lib().oracle_grouped_mutual_u8(_p(qpool), _p(q_gather), _p(q_off), _p(tpool), _p(t_off),
                                   G, _p(d2), _p(idx), _p(t2q))
    return d2, idx, t2q


def c_merge_top2(keys):
    keys = np.ascontiguousarray(keys, np.uint64)
    S, M, _ = keys.shape
    out = np.empty((M, 2), np.uint64)
    lib().oracle_merge_top2(_p(keys), S, M, _p(out))
    return out


def c_ratio(num_d2, den_d2=None, den_f32=None, tau=0.7):
    num_d2 = np.ascontiguousarray(num_d2, np.uint32)
    M = len(num_d2)
    ratio = np.empty(M, np.float64)
    mask = np.empty(M, np.uint8)
    if den_f32 is not None:
        den_f32 = np.ascontiguousarray(den_f32, np.float32)
    else:
        den_d2 = np.ascontiguousarray(den_d2, np.uint32)
    lib().oracle_ratio(_p(num_d2), 1, _p(den_d2), 1, _p(den_f32), M, float(tau), _p(ratio), _p(mask))
    return ratio, mask.astype(bool)
