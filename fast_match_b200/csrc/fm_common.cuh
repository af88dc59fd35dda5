// fm_common.cuh -- shared helpers for libfmatch.so (error reporting, packed keys).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fastmatch_b200.h"

namespace fm {

void set_error(const char *fmt, ...);

#define FM_CUDA_TRY(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            fm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                          __FILE__, __LINE__);                                         \
            return FM_ECUDA;                                                           \
        }                                                                              \
    } while (0)

// Packed candidate: d2 in the high word, index in the low word.  Unsigned order on
// the key is the lexicographic (d2, index) order the reference's matcher produces.
__host__ __device__ inline unsigned long long pack_key(uint32_t d2, uint32_t idx) {
    return ((unsigned long long)d2 << 32) | (unsigned long long)idx;
}

// Insert into an ascending 2-slot list.
__device__ __forceinline__ void insert2(unsigned long long k, unsigned long long &m1,
                                        unsigned long long &m2) {
    if (k < m2) {
        if (k < m1) { m2 = m1; m1 = k; }
        else m2 = k;
    }
}

// Merge two ascending 2-slot lists (a1<=a2, b1<=b2) into (a1, a2).
__device__ __forceinline__ void merge2(unsigned long long &a1, unsigned long long &a2,
                                       unsigned long long b1, unsigned long long b2) {
    unsigned long long lo = a1 < b1 ? a1 : b1;
    unsigned long long hi = a1 < b1 ? b1 : a1;
    unsigned long long h2 = a2 < b2 ? a2 : b2;
    a1 = lo;
    a2 = hi < h2 ? hi : h2;
}

// Lowe ratio test on squared distances, exactly as the reference computes it from DMatch
// distances (fastmatch.pyx:124,165; Classic Matching.ipynb cell 3): float32 sqrt, float64 divide.
__device__ __forceinline__ double ratio_f32sqrt(uint32_t num_d2, uint32_t den_d2) {
    if (num_d2 == FM_NONE_D2 || den_d2 == FM_NONE_D2) return __longlong_as_double(0x7FF0000000000000ll);
    return __ddiv_rn((double)__fsqrt_rn((float)num_d2), (double)__fsqrt_rn((float)den_d2));
}

struct RatioOut {            // optional fused ratio-test outputs of the dense kernels
    double tau;
    double *ratio;           // [M] or null
    uint8_t *mask;           // [M] or null
};

__device__ __forceinline__ void write_ratio(const RatioOut &r, int64_t row, uint32_t d2a, uint32_t d2b) {
    if (r.ratio == nullptr && r.mask == nullptr) return;
    const double v = ratio_f32sqrt(d2a, d2b);
    if (r.ratio) r.ratio[row] = v;
    if (r.mask) r.mask[row] = v < r.tau ? 1 : 0;
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// Instrumentation shared by the translation units (fm_api.cu owns the storage).
//  * every kernel launch of the library is counted (fm_launch_count);
//  * when profiling is on (fm_profile_enable) the dominant kernel of each call is bracketed by a
//    CUDA event pair on the launching stream; fm_profile_read sums the elapsed times.
void count_launch(int n = 1);
void prof_begin(cudaStream_t s);   // no-op unless profiling is enabled
void prof_end(cudaStream_t s);

// launchers implemented in the kernel translation units
int launch_sweep_mma_dense(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N,
                           int32_t t_index_base, uint32_t *d2, int32_t *idx, uint64_t *keys,
                           cudaStream_t s);
int launch_sweep_mma_grouped(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                             const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base,
                             int32_t G, int64_t total_q, int64_t total_t, int32_t max_nq, uint32_t *q2t_d2,
                             int32_t *q2t_idx, int32_t *t2q_idx, uint8_t *mutual,
                             unsigned long long *colkeys, cudaStream_t s);
int launch_top2_tc(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, int32_t t_index_base,
                   uint32_t *d2, int32_t *idx, uint64_t *keys, RatioOut rout, void *ws,
                   size_t ws_bytes, cudaStream_t s);
size_t top2_tc_workspace_bytes(int64_t M, int64_t N);
size_t grouped_tc_workspace_bytes(int64_t total_q, int64_t tpool_rows, bool gather);
int launch_grouped_tc(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                      const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base, int32_t G,
                      int64_t total_q, int64_t total_t, int64_t tpool_rows, uint32_t *q2t_d2,
                      int32_t *q2t_idx, int32_t *t2q_idx, uint8_t *mutual, void *ws, size_t ws_bytes,
                      cudaStream_t s);
bool tc_supported();

}  // namespace fm
