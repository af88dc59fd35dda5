/*
 * oracle.c -- CPU restatement of Fast-Match's descriptor-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fast_match_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: the reference (arnfred/Fast-Match) ships no tests or golden
 * vectors for this path, and the arithmetic lives in a
 * third-party dependency that is NOT under /root/reference: OpenCV
 * (cv2.BFMatcher -> cv::batchDistance; version unpinned by the reference,
 * cv2 4.13.0 in this image).  The oracle is therefore pinned against outputs
 * of cv2.BFMatcher itself, generated in the build container by
 * tests/golden/make_golden.py (script and vectors committed).
 *
 * Algorithm restated (observable behaviour of cv::BFMatcher::knnMatchImpl as
 * called from the reference):
 *   matchutil.py:39-43      bf_match(dt1, dt2, k, crossCheck)
 *   cache.pyx:250-252       thumbnail self-match, k=2, keep r[1].distance
 *   cache.pyx:271-273       full-image self-match, k=2 (FLANN in the reference;
 *                           the exact top-2 is what it approximates)
 *   fastmatch.pyx:122-124   thumbnail mutual-NN + ratio vs cached self distance
 *   fastmatch.pyx:161-165   per-cell mutual-NN + ratio
 *   Classic Matching.ipynb cell 3 (JSON :59-73)  Ratio-Match top-2, d1/d2
 *
 * Descriptors are rows of 128 uint8 (SIFT values are integers 0..255), so the
 * squared L2 distance is an exact integer <= 128*255^2 = 8323200.  Selection
 * is lexicographic on (d2, index): ties go to the lowest index, which is what
 * batchDistance's strict "<" insertion produces.
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define FM_DIM 128
#define FM_NONE_D2 0xFFFFFFFFu

#if defined(__x86_64__) && defined(__GNUC__) && !defined(FM_ORACLE_NO_CLONES)
#define FM_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define FM_CLONES
#endif

/* squared L2 distance between two 128-byte descriptors (exact, int32) */
FM_CLONES
static uint32_t d2_u8(const uint8_t *a, const uint8_t *b) {
    int32_t acc = 0;
    for (int k = 0; k < FM_DIM; ++k) {
        int32_t d = (int32_t)a[k] - (int32_t)b[k];
        acc += d * d;
    }
    return (uint32_t)acc;
}

/* Strict-"<" insertion into a 2-slot list scanned in increasing index order:
 * equal distances keep the earlier (lower) index.  batchDistance semantics. */
static inline void insert2(uint32_t d, int32_t j, uint32_t *d0, int32_t *i0,
                           uint32_t *d1, int32_t *i1) {
    if (d < *d0) { *d1 = *d0; *i1 = *i0; *d0 = d; *i0 = j; }
    else if (d < *d1) { *d1 = d; *i1 = j; }
}

/*
 * Exact top-2 nearest neighbours of every query row among the target rows.
 * Follows bf_match(k=2) (matchutil.py:39-43) / match_bf (Classic Matching cell 3).
 * d2  [M][2] uint32  squared distances (0xFFFFFFFF when the slot is missing)
 * idx [M][2] int32   target index + t_index_base (-1 when missing)
 */
FM_CLONES
void oracle_top2_u8(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N,
                    int32_t t_index_base, uint32_t *d2, int32_t *idx) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < M; ++i) {
        const uint8_t *qi = q + i * FM_DIM;
        uint32_t d0 = FM_NONE_D2, d1 = FM_NONE_D2;
        int32_t i0 = -1, i1 = -1;
        for (int64_t j = 0; j < N; ++j) {
            const uint8_t *tj = t + j * FM_DIM;
            int32_t acc = 0;
            for (int k = 0; k < FM_DIM; ++k) {
                int32_t d = (int32_t)qi[k] - (int32_t)tj[k];
                acc += d * d;
            }
            insert2((uint32_t)acc, (int32_t)j, &d0, &i0, &d1, &i1);
        }
        d2[2 * i] = d0; d2[2 * i + 1] = d1;
        idx[2 * i] = i0 < 0 ? -1 : i0 + t_index_base;
        idx[2 * i + 1] = i1 < 0 ? -1 : i1 + t_index_base;
    }
}

/*
 * One mutual-nearest-neighbour round: BFMatcher(NORM_L2, crossCheck=True)
 * .knnMatch(q, t, k=1) as used at fastmatch.pyx:122-123 and :161-162.
 * q_gather (nullable) selects rows of q (Metric_Cache.get's fancy-index copy,
 * cache.pyx:188); local query i is row q_gather[i].
 * Outputs (all local indices):
 *   q2t_d2 [nq][2], q2t_idx [nq][2]  top-2 of each query among the targets
 *   t2q_idx [nt]                     nearest query of each target (-1 if nq==0)
 * The mutual set is { i : q2t_idx[i][0] >= 0 && t2q_idx[q2t_idx[i][0]] == i }.
 */
void oracle_mutual_u8(const uint8_t *q, const int32_t *q_gather, int64_t nq,
                      const uint8_t *t, int64_t nt, uint32_t *q2t_d2,
                      int32_t *q2t_idx, int32_t *t2q_idx) {
    uint32_t *best_t = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nt > 0 ? nt : 1));
    for (int64_t j = 0; j < nt; ++j) { best_t[j] = FM_NONE_D2; t2q_idx[j] = -1; }
    for (int64_t i = 0; i < nq; ++i) {
        const uint8_t *qi = q + (int64_t)(q_gather ? q_gather[i] : i) * FM_DIM;
        uint32_t d0 = FM_NONE_D2, d1 = FM_NONE_D2;
        int32_t i0 = -1, i1 = -1;
        for (int64_t j = 0; j < nt; ++j) {
            uint32_t d = d2_u8(qi, t + j * FM_DIM);
            insert2(d, (int32_t)j, &d0, &i0, &d1, &i1);
            if (d < best_t[j]) { best_t[j] = d; t2q_idx[j] = (int32_t)i; }
        }
        q2t_d2[2 * i] = d0; q2t_d2[2 * i + 1] = d1;
        q2t_idx[2 * i] = i0; q2t_idx[2 * i + 1] = i1;
    }
    free(best_t);
}

/*
 * Many independent mutual-NN rounds (the flood-fill's match_position calls,
 * fastmatch.pyx:145-169, batched).  Group g uses local queries
 * [q_off[g], q_off[g+1]) -- rows q_gather[.] of qpool when q_gather != NULL,
 * else rows of qpool directly -- against rows [t_off[g], t_off[g+1]) of tpool.
 * Outputs are indexed by local query / local target position in the
 * concatenation; indices are local to the group.
 */
void oracle_grouped_mutual_u8(const uint8_t *qpool, const int32_t *q_gather,
                              const int64_t *q_off, const uint8_t *tpool,
                              const int64_t *t_off, int32_t G, uint32_t *q2t_d2,
                              int32_t *q2t_idx, int32_t *t2q_idx) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int32_t g = 0; g < G; ++g) {
        int64_t q0 = q_off[g], nq = q_off[g + 1] - q0;
        int64_t t0 = t_off[g], nt = t_off[g + 1] - t0;
        const uint8_t *qb = q_gather ? qpool : qpool + q0 * FM_DIM;
        const int32_t *gi = q_gather ? q_gather + q0 : NULL;
        oracle_mutual_u8(qb, gi, nq, tpool + t0 * FM_DIM, nt, q2t_d2 + 2 * q0,
                         q2t_idx + 2 * q0, t2q_idx + t0);
    }
}

/*
 * Merge of per-shard top-2 candidates (new in the sharded path; it restates
 * "top-2 over the union of the shards").  keys [S][M][2] are packed
 * (d2 << 32 | global index); out [M][2] are the two smallest per query.
 * Unsigned order on the packed key is the lexicographic (d2, index) order.
 */
void oracle_merge_top2(const uint64_t *keys, int32_t S, int64_t M, uint64_t *out) {
    for (int64_t i = 0; i < M; ++i) {
        uint64_t a = ~0ull, b = ~0ull;
        for (int32_t s = 0; s < S; ++s)
            for (int k = 0; k < 2; ++k) {
                uint64_t v = keys[((int64_t)s * M + i) * 2 + k];
                if (v < a) { b = a; a = v; } else if (v < b) { b = v; }
            }
        out[2 * i] = a; out[2 * i + 1] = b;
    }
}

/*
 * Ratio test.  DMatch.distance is float32(sqrt(float32(d2))); the reference
 * divides two such values as Python floats (doubles) and compares "< tau":
 *   fastmatch.pyx:124, :165  ratio = m.distance / cached_self_distance[queryIdx]
 *   Classic Matching cell 3  ratio = m[0].distance / m[1].distance
 *   fastmatch.pyx:75, :82    ratios < tau
 * den_f32 (nullable): float32 denominators; if NULL, den_d2 is used the same
 * way as the numerator.  A missing numerator (0xFFFFFFFF) gives ratio = +inf
 * and mask 0; a zero denominator follows IEEE (x/0 = inf, 0/0 = nan -> mask 0).
 */
void oracle_ratio(const uint32_t *num_d2, int64_t num_stride, const uint32_t *den_d2,
                  int64_t den_stride, const float *den_f32, int64_t M, double tau,
                  double *ratio, uint8_t *mask) {
    for (int64_t i = 0; i < M; ++i) {
        uint32_t n = num_d2[i * num_stride];
        uint32_t dd = den_f32 ? 0u : den_d2[i * den_stride];
        double r;
        if (n == FM_NONE_D2 || dd == FM_NONE_D2) r = INFINITY;
        else {
            double num = (double)sqrtf((float)n);
            double den = den_f32 ? (double)den_f32[i] : (double)sqrtf((float)dd);
            r = num / den;
        }
        ratio[i] = r;
        mask[i] = (r < tau) ? 1 : 0;
    }
}

int oracle_version(void) { return 1; }
