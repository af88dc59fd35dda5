/*
 * fastmatch_b200.h -- C-ABI of libfmatch.so: the B200-native replacement for the
 * nearest-neighbour calls on Fast-Match's descriptor-matching hot path.
 *
 * The reference has no FFI of its own for this path: it calls OpenCV's
 * cv2.BFMatcher / cv2.FlannBasedMatcher from Python.  Each entry point below
 * names the reference call site(s) (file:line under the reference checkout)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer would add to matchutil.py / fastmatch.pyx.
 *
 * Conventions
 *   - descriptors: row-major uint8, FM_DIM (=128) bytes per row, base pointer
 *     16-byte aligned (SIFT descriptors are integers 0..255, so u8 is exact);
 *   - "device" entry points take device pointers owned by the caller, run
 *     asynchronously on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream), never free or retain caller memory and are
 *     re-entrant across streams;
 *   - every function returns FM_OK (0) or a negative FM_E* code; the message of
 *     the last failure on the calling thread is fm_last_error();
 *   - squared distances are exact integers (<= 128*255^2 = 8323200); a missing
 *     neighbour (fewer than 2 targets) has d2 = FM_NONE_D2 and idx = -1;
 *   - selection is lexicographic on (d2, index): ties go to the lowest index,
 *     as cv::batchDistance's strict "<" insertion does.
 * No exceptions and no torch/C++ types cross this boundary.
 */
#ifndef FASTMATCH_B200_H_
#define FASTMATCH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_DIM 128
#define FM_NONE_D2 0xFFFFFFFFu
#define FM_NONE_KEY 0xFFFFFFFFFFFFFFFFull

#define FM_OK 0
#define FM_EINVAL -1   /* bad argument (null pointer, negative size, misalignment) */
#define FM_ECUDA -2    /* a CUDA runtime / driver call failed                      */
#define FM_ENOSPACE -3 /* workspace too small                                      */
#define FM_EUNSUPPORTED -4 /* device is not sm_100 / requested algorithm unavailable */

/* dense algorithm selector for fm_top2_u8 */
#define FM_ALGO_AUTO 0
#define FM_ALGO_MMA_SYNC 1 /* portable warp-MMA kernel (small problems, differential check) */
#define FM_ALGO_TCGEN05 2  /* sm_100a tcgen05.mma kind::i8 + TMA + TMEM kernel               */

int fm_version(void);
const char *fm_last_error(void);

/* sm_major/sm_minor/sm_count of `device`; has_tcgen05 = 1 on sm_100 parts. */
int fm_device_caps(int device, int *sm_major, int *sm_minor, int *sm_count, int *has_tcgen05);

/*
 * Exact top-2 nearest neighbours of M query descriptors among N target
 * descriptors.  Replaces cv2.BFMatcher(NORM_L2).knnMatch(q, t, k=2) (and k=1):
 *   matchutil.py:39-43 (bf_match), cache.pyx:250-252 and :271-273 (self-match,
 *   q == t), Classic Matching.ipynb cell 3 JSON :63-65 (Ratio-Match).
 *   d2   [M][2] uint32   squared distances, ascending
 *   idx  [M][2] int32    target row + t_index_base
 *   keys [M][2] uint64   optional (may be NULL): d2 << 32 | idx, FM_NONE_KEY when
 *                        missing -- the form the sharded path exchanges
 * ws / ws_bytes: scratch of at least fm_top2_workspace_bytes(M, N) bytes.
 */
size_t fm_top2_workspace_bytes(int64_t M, int64_t N);
int fm_top2_u8(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, int32_t t_index_base,
               uint32_t *d2, int32_t *idx, uint64_t *keys, void *ws, size_t ws_bytes, int algo,
               void *stream);

/*
 * Ratio-Match in one call: exact top-2 with the Lowe ratio test fused into the kernel's
 * write-out (Classic Matching.ipynb cell 3: knnMatch(k=2), m[0].distance / m[1].distance,
 * compared with tau).  ratio [M] float64 and/or mask [M] uint8 (ratio < tau); one of the two
 * may be NULL.  A query with fewer than two targets gets ratio = +inf, mask = 0.
 */
int fm_ratio_match_u8(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, double tau,
                      uint32_t *d2, int32_t *idx, double *ratio, uint8_t *mask, void *ws,
                      size_t ws_bytes, int algo, void *stream);

/*
 * Ratio test on squared distances.  Replaces the Python-side arithmetic
 *   fastmatch.pyx:124 and :165  ratio = m.distance / cached_distance[queryIdx]
 *   Classic Matching.ipynb cell 3 JSON :65  ratio = m[0].distance / m[1].distance
 *   fastmatch.pyx:75, :82       ratios < tau
 * ratio[i] = (double)sqrtf((float)num_d2[i*num_stride]) / den, where den is
 * (double)den_f32[i] if den_f32 != NULL else (double)sqrtf((float)den_d2[i*den_stride]);
 * mask[i] = ratio[i] < tau.  A missing numerator/denominator gives +inf / 0.
 * ratio may be NULL (mask only) and mask may be NULL (ratio only).
 */
int fm_ratio_f32sqrt(const uint32_t *num_d2, int64_t num_stride, const uint32_t *den_d2,
                     int64_t den_stride, const float *den_f32, int64_t M, double tau,
                     double *ratio, uint8_t *mask, void *stream);

/*
 * G independent mutual-nearest-neighbour rounds in one launch.  Replaces the
 * per-round cv2.BFMatcher(NORM_L2, crossCheck=True).knnMatch(query_ds,
 * target_ds, k=1) of fastmatch.pyx:161-162 (match_position, one call per
 * flood-fill round) and :122-123 (match_thumbs), plus the fancy-index gather
 * of cache.pyx:188 when q_gather is given.
 *   group g: local queries [q_off[g], q_off[g+1]) -- rows q_gather[.] of qpool
 *   if q_gather != NULL, else rows of qpool -- against rows
 *   [t_off[g], t_off[g+1]) of tpool -- or, if t_base != NULL, the t_off[g+1]-t_off[g]
 *   rows of tpool starting at row t_base[g] (lets many groups share one resident cell;
 *   outputs stay indexed by t_off).  Empty groups / sides are legal.
 *   q2t_d2 [total_q][2], q2t_idx [total_q][2]: top-2 per local query (indices
 *   local to the group); t2q_idx [total_t]: nearest local query per target
 *   (-1 if the group has no queries).  crossCheck keeps local query i iff
 *   t2q_idx[t_off[g] + q2t_idx[i][0]] == i - q_off[g].
 *   mutual [total_q] uint8 (may be NULL): that predicate, fused.
 * total_q / total_t = q_off[G] / t_off[G]; tpool_rows = number of rows of tpool (only read
 * when t_base != NULL, else total_t is used); max_nq = an upper bound on the number of queries
 * of any group (host-known; avoids a device->host sync).  algo: FM_ALGO_AUTO picks the
 * tcgen05 kernel on sm_100; FM_ALGO_MMA_SYNC forces the warp-MMA kernel.
 * Every descriptor is read from HBM once (the squared norms are summed in the kernel).
 * A group is processed by ONE thread block (the path is built for thousands of small rounds):
 * for a single large pair of sets (match_thumbs, bf_match(crossCheck=True) on whole images) two
 * fm_top2_u8 calls, q->t and t->q, give the same mutual pairs an order of magnitude faster
 * (fast_match_b200.backend.mutual_single does that above 2^16 pairs).
 */
size_t fm_grouped_workspace_bytes(int64_t total_q, int64_t total_t, int64_t tpool_rows, int32_t G);
int fm_grouped_mutual_u8(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                         const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base,
                         int32_t G, int64_t total_q, int64_t total_t, int64_t tpool_rows,
                         int32_t max_nq, uint32_t *q2t_d2, int32_t *q2t_idx, int32_t *t2q_idx,
                         uint8_t *mutual, void *ws, size_t ws_bytes, int algo, void *stream);

/*
 * Merge per-shard top-2 candidates (new: the exchange step of the target-
 * sharded path; no reference counterpart because cv2.BFMatcher cannot hold
 * >= 2^18 train rows).  keys [S][M][2] packed d2 << 32 | global idx;
 * out_keys [M][2] the two smallest; d2/idx (nullable) the unpacked form.
 */
int fm_merge_top2(const uint64_t *keys, int32_t S, int64_t M, uint64_t *out_keys, uint32_t *d2,
                  int32_t *idx, void *stream);

/*
 * Host-buffer convenience: what matchutil.bf_match(dt1, dt2, k=2) binds when the
 * caller holds numpy arrays.  Copies q/t to the device (pinned staging owned by
 * the library), runs fm_top2_u8, copies the result back and synchronises.
 * dist (nullable) [M][2] float32 = sqrtf((float)d2), i.e. DMatch.distance.
 * mask (nullable) [M] uint8 = Lowe ratio test dist[i][0] / dist[i][1] < tau (Classic
 * Matching.ipynb cell 3 JSON :65), computed on the device in the same call.
 * Pinned / cudaHostRegister'ed caller buffers are used directly; pageable ones are staged.
 * One context (two streams, staging buffers) per device; the caller's current device is restored.
 * The library measures the upload rate of its previous calls on the device and, where the link is
 * slow, uploads the queries in two halves so that the second half travels under the first half's
 * compute (FM_HOST_HALVES=1|2 in the environment forces one form).
 */
int fm_top2_host_u8(const uint8_t *q_host, int64_t M, const uint8_t *t_host, int64_t N,
                    uint32_t *d2_host, int32_t *idx_host, float *dist_host, double tau,
                    uint8_t *mask_host, int device);

/*
 * Instrumentation (no reference counterpart).  fm_launch_count: kernels launched by this
 * library since load.  fm_profile_enable(1): bracket the dominant kernel of every following
 * fm_top2_u8 / fm_grouped_mutual_u8 call with CUDA events on the launching stream;
 * fm_profile_read sums their elapsed milliseconds (synchronises on the events).
 */
long long fm_launch_count(void);
int fm_profile_enable(int on);
int fm_profile_read(double *total_ms, int *launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* FASTMATCH_B200_H_ */
