"""Synthetic SIFT-like descriptor sets.

Uniform-random u8 vectors never pass a ratio test (min ratio ~0.9), so the bench
and the parity tests use descriptors with SIFT's statistics instead: Gamma(0.6)
magnitudes, L2-normalised, clipped at 0.2, renormalised, scaled to norm 512 and
rounded to u8 -- the same post-processing cv2's SIFT applies.  A fraction of the
targets are noisy copies of queries (so ratios spread over tau) and 1% of the
targets are exact duplicates of other targets (forces ties -> lowest index wins).
"""
import numpy as np


def siftlike(n, rng):
    x = rng.gamma(0.6, 1.0, size=(n, 128)).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    np.minimum(x, 0.2, out=x)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return np.clip(np.rint(512.0 * x), 0, 255).astype(np.uint8)


def make_pair(M, N, seed, p_match=0.5, dup_frac=0.01, sigma=(4.0, 40.0), chunk=1 << 16):
    """Query set [M,128] u8 and target set [N,128] u8 with planted matches and ties."""
    rng = np.random.default_rng(seed)
    q = np.empty((M, 128), np.uint8)
    t = np.empty((N, 128), np.uint8)
    for a in (q, t):
        for s in range(0, len(a), chunk):
            a[s:s + chunk] = siftlike(min(chunk, len(a) - s), rng)
    n_pl = int(p_match * min(M, N))
    if n_pl:
        qi = rng.choice(M, n_pl, replace=False)
        ti = rng.choice(N, n_pl, replace=False)
        for s in range(0, n_pl, chunk):
            sl = slice(s, s + chunk)
            sg = rng.uniform(sigma[0], sigma[1], size=(len(qi[sl]), 1)).astype(np.float32)
            noise = rng.standard_normal((len(qi[sl]), 128), dtype=np.float32) * sg
            t[ti[sl]] = np.clip(np.rint(q[qi[sl]].astype(np.float32) + noise), 0, 255).astype(np.uint8)
    n_dup = int(dup_frac * N)
    if n_dup and N > 1:
        dst = rng.choice(N, n_dup, replace=False)
        src = rng.integers(0, N, n_dup)
        t[dst] = t[src]
    return q, t


def make_groups(G, lo, hi, seed):
    """Grouped workload (config 4): G groups with n_q, n_t ~ U[lo, hi], packed contiguously."""
    rng = np.random.default_rng(seed)
    nq = rng.integers(lo, hi + 1, G)
    nt = rng.integers(lo, hi + 1, G)
    q_off = np.zeros(G + 1, np.int64)
    t_off = np.zeros(G + 1, np.int64)
    np.cumsum(nq, out=q_off[1:])
    np.cumsum(nt, out=t_off[1:])
    Q, T = int(q_off[-1]), int(t_off[-1])
    qpool = np.empty((Q, 128), np.uint8)
    tpool = np.empty((T, 128), np.uint8)
    chunk = 1 << 16
    for a in (qpool, tpool):
        for s in range(0, len(a), chunk):
            a[s:s + chunk] = siftlike(min(chunk, len(a) - s), rng)
    # plant: in every group, half of the smaller side are noisy copies; a few exact duplicates
    for g in range(G):
        k = int(min(nq[g], nt[g]) // 2)
        if k == 0:
            continue
        qi = q_off[g] + rng.choice(nq[g], k, replace=False)
        ti = t_off[g] + rng.choice(nt[g], k, replace=False)
        sg = rng.uniform(4.0, 40.0, size=(k, 1)).astype(np.float32)
        noise = rng.standard_normal((k, 128), dtype=np.float32) * sg
        tpool[ti] = np.clip(np.rint(qpool[qi].astype(np.float32) + noise), 0, 255).astype(np.uint8)
        if nt[g] >= 4:
            tpool[t_off[g] + rng.integers(0, nt[g])] = tpool[t_off[g] + rng.integers(0, nt[g])]
        if nq[g] >= 4:
            qpool[q_off[g] + rng.integers(0, nq[g])] = qpool[q_off[g] + rng.integers(0, nq[g])]
    return qpool, q_off, tpool, t_off
