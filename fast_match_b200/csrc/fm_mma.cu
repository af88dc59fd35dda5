// fm_mma.cu -- warp-MMA (mma.sync m16n8k32 u8) sweep kernel.
//
// One kernel serves two entry points:
//   * fm_top2_u8 with FM_ALGO_MMA_SYNC: dense top-2 for small problems and as the
//     on-device differential check of the tcgen05 kernel;
//   * fm_grouped_mutual_u8: thousands of independent (query tile x cell) rounds in
//     one launch, row top-2 and column top-1 (crossCheck) from the same accumulators.
// Replaces cv2.BFMatcher(...).knnMatch at matchutil.py:42-43 and fastmatch.pyx:122-123,
// 161-162 of the reference.  d2 = |q|^2 + |t|^2 - 2 q.t in exact integer arithmetic.
//
// Layout: a CTA owns a slab of 128 query rows (8 warps x 16 rows).  Each warp keeps its
// A fragments (16 rows x 128 bytes) in registers for the whole sweep; target rows stream
// through a double-buffered cp.async ring of 64-row chunks (row stride padded to 144 B so
// the B-fragment loads are bank-conflict free).  The distance tile only ever exists as
// mma accumulators.
#include "fm_common.cuh"

namespace fm {

namespace {

constexpr int TM = 128;      // query rows per slab
constexpr int TN = 64;       // target rows per chunk
constexpr int ROWB = 144;    // padded smem row stride (bytes)
constexpr int NTHREADS = 256;
constexpr unsigned long long NONE = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                       uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;  // src-size 0 -> destination is zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// load one 64-row chunk of targets into smem (rows beyond nt are zero-filled)
__device__ __forceinline__ void load_chunk(uint8_t *sB, const uint8_t *tbase, int64_t row0,
                                           int64_t nt, int tid) {
#pragma unroll
    for (int i = 0; i < (TN * 8) / NTHREADS; ++i) {
        int piece = tid + i * NTHREADS;  // 0..511
        int r = piece >> 3, c16 = piece & 7;
        int64_t row = row0 + r;
        bool ok = row < nt;
        const uint8_t *src = tbase + (ok ? row : 0) * FM_DIM + c16 * 16;
        cp_async16(sB + r * ROWB + c16 * 16, src, ok);
    }
}

template <bool COLMIN>
__global__ void __launch_bounds__(NTHREADS)
k_sweep_mma(const uint8_t *__restrict__ qpool, const int32_t *__restrict__ q_gather,
            const int64_t *__restrict__ q_off, const uint8_t *__restrict__ tpool,
            const int64_t *__restrict__ t_off, const int64_t *__restrict__ t_base,
            int64_t M_dense, int64_t N_dense,
            int32_t t_index_base, uint32_t *__restrict__ out_d2, int32_t *__restrict__ out_idx,
            uint64_t *__restrict__ out_keys, unsigned long long *__restrict__ colkeys) {
    __shared__ __align__(16) uint8_t sB[2][TN * ROWB];
    __shared__ int sTn[2][TN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, tig = lane & 3;

    int64_t q0, nq, t0, nt;
    if (q_off) {
        q0 = q_off[blockIdx.x]; nq = q_off[blockIdx.x + 1] - q0;
        t0 = t_off[blockIdx.x]; nt = t_off[blockIdx.x + 1] - t0;
    } else {
        q0 = 0; nq = M_dense; t0 = 0; nt = N_dense;
    }
    const uint8_t *tbase = tpool + ((q_off && t_base) ? t_base[blockIdx.x] : t0) * FM_DIM;
    const int64_t nchunks = (nt + TN - 1) / TN;

    for (int64_t slab = blockIdx.y; slab * TM < nq; slab += gridDim.y) {
        // ---- A fragments: rows r0 (= g8) and r1 (= g8 + 8) of this warp's 16 rows
        const int64_t r0 = slab * TM + warp * 16 + g8, r1 = r0 + 8;
        const bool v0 = r0 < nq, v1 = r1 < nq;
        const uint8_t *p0 = qpool, *p1 = qpool;
        if (v0) p0 = qpool + (q_gather ? (int64_t)q_gather[q0 + r0] : q0 + r0) * FM_DIM;
        if (v1) p1 = qpool + (q_gather ? (int64_t)q_gather[q0 + r1] : q0 + r1) * FM_DIM;
        uint32_t a[4][4];
        int qn0 = 0, qn1 = 0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            a[ks][0] = v0 ? *(const uint32_t *)(p0 + ks * 32 + 4 * tig) : 0u;
            a[ks][1] = v1 ? *(const uint32_t *)(p1 + ks * 32 + 4 * tig) : 0u;
            a[ks][2] = v0 ? *(const uint32_t *)(p0 + ks * 32 + 16 + 4 * tig) : 0u;
            a[ks][3] = v1 ? *(const uint32_t *)(p1 + ks * 32 + 16 + 4 * tig) : 0u;
            qn0 = __dp4a(a[ks][0], a[ks][0], (unsigned)qn0);
            qn0 = __dp4a(a[ks][2], a[ks][2], (unsigned)qn0);
            qn1 = __dp4a(a[ks][1], a[ks][1], (unsigned)qn1);
            qn1 = __dp4a(a[ks][3], a[ks][3], (unsigned)qn1);
        }
        qn0 += __shfl_xor_sync(0xffffffffu, qn0, 1);
        qn0 += __shfl_xor_sync(0xffffffffu, qn0, 2);
        qn1 += __shfl_xor_sync(0xffffffffu, qn1, 1);
        qn1 += __shfl_xor_sync(0xffffffffu, qn1, 2);

        unsigned long long m1[2] = {NONE, NONE}, m2[2] = {NONE, NONE};

        __syncthreads();  // previous slab's readers are done with sB / sTn
        if (nchunks > 0) load_chunk(sB[0], tbase, 0, nt, tid);
        cp_async_commit();

        for (int64_t c = 0; c < nchunks; ++c) {
            const int buf = (int)(c & 1);
            if (c + 1 < nchunks) load_chunk(sB[buf ^ 1], tbase, (c + 1) * TN, nt, tid);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();  // chunk c is visible to every thread

            {   // |t|^2 of the chunk: 4 threads per row, 32 bytes each
                int r = tid >> 2, qtr = tid & 3;
                const uint4 *p = (const uint4 *)(sB[buf] + r * ROWB + qtr * 32);
                uint4 x = p[0], y = p[1];
                unsigned s = 0;
                s = __dp4a(x.x, x.x, s); s = __dp4a(x.y, x.y, s);
                s = __dp4a(x.z, x.z, s); s = __dp4a(x.w, x.w, s);
                s = __dp4a(y.x, y.x, s); s = __dp4a(y.y, y.y, s);
                s = __dp4a(y.z, y.z, s); s = __dp4a(y.w, y.w, s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (qtr == 0) sTn[buf][r] = (int)s;
            }

            int acc[TN / 8][4];
#pragma unroll
            for (int n8 = 0; n8 < TN / 8; ++n8) {
                acc[n8][0] = acc[n8][1] = acc[n8][2] = acc[n8][3] = 0;
                const uint8_t *brow = sB[buf] + (n8 * 8 + g8) * ROWB + 4 * tig;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t b0 = *(const uint32_t *)(brow + ks * 32);
                    uint32_t b1 = *(const uint32_t *)(brow + ks * 32 + 16);
                    mma_u8(acc[n8], a[ks], b0, b1);
                }
            }
            __syncthreads();  // sTn[buf] complete

            const int64_t jbase = c * TN;
#pragma unroll
            for (int n8 = 0; n8 < TN / 8; ++n8) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = n8 * 8 + 2 * tig + e;
                    const int64_t j = jbase + col;
                    const int tn = sTn[buf][col];
                    const bool jv = j < nt;
                    // rows r0 / r1
                    uint32_t d0 = (uint32_t)(qn0 + tn - 2 * acc[n8][e]);
                    uint32_t d1 = (uint32_t)(qn1 + tn - 2 * acc[n8][2 + e]);
                    if (jv) {
                        uint32_t jj = (uint32_t)((int32_t)j + t_index_base);
                        insert2(pack_key(d0, jj), m1[0], m2[0]);
                        insert2(pack_key(d1, jj), m1[1], m2[1]);
                    }
                    if (COLMIN) {
                        unsigned long long k0 = (jv && v0) ? pack_key(d0, (uint32_t)r0) : NONE;
                        unsigned long long k1 = (jv && v1) ? pack_key(d1, (uint32_t)r1) : NONE;
                        unsigned long long k = k0 < k1 ? k0 : k1;
#pragma unroll
                        for (int m = 4; m <= 16; m <<= 1) {
                            unsigned long long o = shfl_xor_u64(k, m);
                            k = o < k ? o : k;
                        }
                        if (g8 == 0 && k != NONE) atomicMin(colkeys + t0 + j, k);
                    }
                }
            }
        }
        cp_async_wait<0>();

        // ---- combine the 4 lanes that share a row, write out
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int m = 1; m <= 2; m <<= 1) {
                unsigned long long b1 = shfl_xor_u64(m1[r], m), b2 = shfl_xor_u64(m2[r], m);
                merge2(m1[r], m2[r], b1, b2);
            }
        }
        if (tig == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int64_t row = r ? r1 : r0;
                if (row >= nq) continue;
                const int64_t o = (q0 + row) * 2;
                const unsigned long long ka = m1[r], kb = m2[r];
                out_d2[o] = (uint32_t)(ka >> 32);
                out_d2[o + 1] = (uint32_t)(kb >> 32);
                out_idx[o] = ka == NONE ? -1 : (int32_t)(uint32_t)ka;
                out_idx[o + 1] = kb == NONE ? -1 : (int32_t)(uint32_t)kb;
                if (out_keys) { out_keys[o] = ka; out_keys[o + 1] = kb; }
            }
        }
    }
}

// per group: colkeys -> t2q_idx, then the fused crossCheck predicate
__global__ void k_grouped_finalize(const int64_t *__restrict__ q_off,
                                   const int64_t *__restrict__ t_off,
                                   const unsigned long long *__restrict__ colkeys,
                                   const int32_t *__restrict__ q2t_idx,
                                   int32_t *__restrict__ t2q_idx, uint8_t *__restrict__ mutual) {
    const int g = blockIdx.x;
    const int64_t q0 = q_off[g], nq = q_off[g + 1] - q0;
    const int64_t t0 = t_off[g], nt = t_off[g + 1] - t0;
    for (int64_t j = threadIdx.x; j < nt; j += blockDim.x) {
        unsigned long long k = colkeys[t0 + j];
        t2q_idx[t0 + j] = k == NONE ? -1 : (int32_t)(uint32_t)k;
    }
    if (!mutual) return;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < nq; i += blockDim.x) {
        int32_t j = q2t_idx[(q0 + i) * 2];
        mutual[q0 + i] = (j >= 0 && t2q_idx[t0 + j] == (int32_t)i) ? 1 : 0;
    }
}

}  // namespace

int launch_sweep_mma_dense(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N,
                           int32_t t_index_base, uint32_t *d2, int32_t *idx, uint64_t *keys,
                           cudaStream_t s) {
    if (M == 0) return FM_OK;
    int64_t slabs = (M + TM - 1) / TM;
    dim3 grid(1, (unsigned)(slabs < 65535 ? slabs : 65535));
    prof_begin(s);
    k_sweep_mma<false><<<grid, NTHREADS, 0, s>>>(q, nullptr, nullptr, t, nullptr, nullptr, M, N,
                                                 t_index_base, d2, idx, keys, nullptr);
    prof_end(s);
    FM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return FM_OK;
}

int launch_sweep_mma_grouped(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                             const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base,
                             int32_t G, int64_t total_q, int64_t total_t, int32_t max_nq, uint32_t *q2t_d2,
                             int32_t *q2t_idx, int32_t *t2q_idx, uint8_t *mutual,
                             unsigned long long *colkeys, cudaStream_t s) {
    if (G == 0) return FM_OK;
    if (total_t > 0) FM_CUDA_TRY(cudaMemsetAsync(colkeys, 0xFF, sizeof(unsigned long long) * total_t, s));
    int64_t slabs = (max_nq + TM - 1) / TM;
    if (slabs < 1) slabs = 1;
    if (slabs > 64) slabs = 64;  // kernel loops over the remaining slabs
    dim3 grid((unsigned)G, (unsigned)slabs);
    prof_begin(s);
    k_sweep_mma<true><<<grid, NTHREADS, 0, s>>>(qpool, q_gather, q_off, tpool, t_off, t_base, 0, 0, 0,
                                                q2t_d2, q2t_idx, nullptr, colkeys);
    prof_end(s);
    FM_CUDA_TRY(cudaGetLastError());
    k_grouped_finalize<<<G, 128, 0, s>>>(q_off, t_off, colkeys, q2t_idx, t2q_idx, mutual);
    FM_CUDA_TRY(cudaGetLastError());
    count_launch(2);
    return FM_OK;
}

}  // namespace fm
