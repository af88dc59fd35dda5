// fm_tc.cu -- dense exact top-2 on the 5th-generation tensor cores (sm_100a).
//
// Replaces cv2.BFMatcher(NORM_L2).knnMatch(q, t, k=2) (matchutil.py:39-43, Classic
// Matching.ipynb cell 3) for large descriptor sets.  d2 = |q|^2 + |t|^2 - 2 q.t, with the
// u8 x u8 -> s32 contraction on tcgen05.mma kind::i8; the M x N distance matrix only ever
// exists as TMEM accumulators.
//
// Work decomposition (details at k_top2_body)
//   * a worker is a CTA pair (cluster of 2, one TPC; the leader issues cta_group::2 MMAs of
//     M = 256 x N = 128) or, for small / ragged M, a single CTA (M = 128 x N = 256 MMAs);
//   * a CTA keeps 256 query rows = two 128-row sub-tiles resident in shared memory (TMA, 128B
//     swizzle: one descriptor = one swizzle row); target tiles of 256 descriptors stream through
//     a TMA/mbarrier ring and every tile is multiplied against both sub-tiles;
//   * accumulators live in TMEM: four 128-column buffers (sub-tile x column half) per CTA of a
//     pair, two 256-column buffers in a single CTA; the epilogue drains one while the tensor core
//     fills the others;
//   * persistent stream-K schedule: the (M-block, tile) steps are cut into equal contiguous ranges,
//     one per worker; a worker leaves the top-2 of each of its runs in a partial-key slot and
//     k_merge_partial folds the slots of a row (and applies the ratio test).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..19 = epilogue: (sub-tile, column half) x TMEM lane quarter (= warp % 4).
//
// The epilogue is the bottleneck (K is only 128: 512 tensor cycles per 32768 outputs, and the
// integer min/max pipe retires 64 lanes/clk/SM), so the per-output work is ~0.5 instruction:
//   * |t_j|^2 is folded into the contraction.  A fifth K-step multiplies a constant query-side
//     block by 32 "digit" bytes per target that encode E_j = ceil((C - |t_j|^2)/2) (C = global
//     constant), so the accumulator is acc'_ij = q_i.t_j + E_j and
//         |t_j|^2 - 2 q_i.t_j  >=  C - 2 acc'_ij          (equality up to rounding of E_j)
//     i.e. the largest accumulator of a row is its nearest neighbour (+25% MMA work buys a
//     filter that needs no per-column term);
//   * per 16 columns a 3-input max tree (0.5 op/output) is compared with the row's bound
//     (C - m2)/2, m2 = current second-best partial distance (shared between the two warps
//     that sweep the same row and, via global memory, between CTAs on other target slices); only chunks that pass recompute exact integer keys
//     (partial*256 + column) and update the row's top-2 with strict "<" in increasing column
//     order, so ties keep the lowest index.  Results are exact for any u8 input.
#include <cuda.h>
#include <math.h>

#include <mutex>
#include <type_traits>

#include "fm_common.cuh"
#include "fm_tc_ptx.cuh"

namespace fm {

namespace tc {

constexpr int BM = 128;                 // rows per sub-tile (UMMA M)
constexpr int SUBS = 2;                 // sub-tiles per CTA
constexpr int BN = 256;                 // targets per tile (UMMA N)
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;   // 640
constexpr int COLS_PER_WARP = BN / 2;    // 128: a warp sweeps one column half of its sub-tile
constexpr int A_BYTES = BM * FM_DIM;    // 16 KB per sub-tile
constexpr int B_BYTES = BN * FM_DIM;    // 32 KB per stage
constexpr int TMEM_COLS = 512;

struct __align__(8) Bars {
    unsigned long long full[8], empty[8], a_full, a_empty, tmem_full[4], tmem_empty[4];
    uint32_t tmem_base;
    uint32_t pad;
};
constexpr int AX_BYTES = BM * 32;       // constant query-side block of the norm K-step (4 KB)
constexpr int BX_BYTES = BN * 32;       // per-target digit block of the norm K-step (8 KB / stage)
constexpr int STAGE_BYTES = B_BYTES + BX_BYTES;
constexpr int SMEM_A = 0;
constexpr int SMEM_AX = SMEM_A + SUBS * A_BYTES;
constexpr int SMEM_B = SMEM_AX + AX_BYTES;
constexpr int SMEM_KEYS = SMEM_B + STAGES * STAGE_BYTES;        // [256 rows][2 column halves][2] u64
constexpr int SMEM_M2 = SMEM_KEYS + SUBS * BM * 2 * 2 * 8;      // [256 rows] shared second-best bound
constexpr int SMEM_CK = SMEM_M2 + SUBS * BM * 4;                // [16 warps][2 slots][128] exact-key constants
constexpr int SMEM_BARS = SMEM_CK + EPI_WARPS * 2 * COLS_PER_WARP * 4;
constexpr int SMEM_TOTAL = SMEM_BARS + (int)sizeof(Bars);
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;                   // slack for 1024-B alignment
static_assert(SMEM_ALLOC <= 232448, "exceeds the 227 KB of shared memory a CTA can use");

constexpr int I32_MAX = 0x7FFFFFFF;
constexpr int NONE_P = 0x7FFFFF;   // "no candidate": above every real partial distance (<= 8323200)

// ---------------------------------------------------------------------------------------------
// pre-pass (one launch): |q_i|^2 for the queries; for every target j the norm digits and the
// exact-key constant.
//   E_j = 2*ceil(max(C - |t_j|^2, 0)/4)  with the fixed constant C = 2*EMAX, written as 16
//   base-255 digits duplicated into both 16-byte halves of a 32-byte row (so the operand is
//   indifferent to the 32B swizzle).  2 E_j >= C - |t_j|^2 always (targets with |t_j|^2 > C,
//   i.e. mean byte value > 174, just get E_j = 0: the filter stays conservative, never wrong).
//   ckey_j = (|t_j|^2 + 2 E_j) * 256 + (j & 255)  (wrapping int32; INT_MAX on the tile padding).
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: the three kernels of a call (pre-pass, sweep, slot merge) are
// chained so that each one is scheduled -- and runs its prologue -- while its predecessor drains;
// pdl_wait() returns once the predecessor has completed and its writes are visible (no-op for a
// kernel that was launched without the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int EHALF_MAX = 255 * 255 * 15 + 254;   // digits 0..14 weigh 255, digit 15 weighs 1
constexpr int EMAX = 2 * EHALF_MAX;
constexpr int CG = 2 * EMAX;                      // the constant C of the filter
__global__ void k_prepass(const uint8_t *__restrict__ q, int64_t M, int64_t m_padded,
                          const uint8_t *__restrict__ t, int64_t N, int64_t n_padded,
                          int *__restrict__ qn, int *__restrict__ gbound, int *__restrict__ ckey,
                          uint4 *__restrict__ digits) {
    pdl_launch_dependents();                      // the sweep may be scheduled; it waits before it reads
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = gt >> 3;                  // 8 threads (16 B each) per descriptor
    const int sub = (int)(gt & 7);
    const bool is_q = row < m_padded;
    const int64_t j = row - m_padded;
    const bool live = is_q ? row < M : j < N;
    unsigned s = 0;
    if (live) {
        const uint8_t *src = is_q ? q + row * FM_DIM : t + j * FM_DIM;
        const uint4 x = *(const uint4 *)(src + sub * 16);
        s = __dp4a(x.x, x.x, s); s = __dp4a(x.y, x.y, s);
        s = __dp4a(x.z, x.z, s); s = __dp4a(x.w, x.w, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (is_q) {
        if (sub == 0) {
            if (row < M) qn[row] = (int)s;
            if (gbound) gbound[row] = NONE_P;     // cross-CTA bound on the second-best distance
        }
        return;
    }
    if (j >= n_padded || sub > 1) return;
    if (j >= N) { if (sub == 0) ckey[j] = I32_MAX; return; }
    const int tn = (int)s;
    const int e = CG - tn;
    const int half = e > 0 ? (e + 3) >> 2 : 0;    // E_j / 2  (<= EHALF_MAX because e <= 2*EMAX)
    if (sub == 0) ckey[j] = (int)(((unsigned)(tn + 4 * half) << 8) | (unsigned)(j & 255));
    int sdig = half / 255;
    const unsigned r = (unsigned)(half - sdig * 255);
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const int d = sdig > 255 ? 255 : sdig;
        sdig -= d;
        w[k >> 2] |= (unsigned)d << (8 * (k & 3));
    }
    w[3] |= r << 24;
    digits[2 * j + sub] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
#ifdef FM_TC_PROF
__device__ unsigned long long g_prof[16];
__device__ unsigned long long g_tl[160 * 8];      // per CTA: globaltimer stamps of the kernel's phases
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(slot) _pacc[slot] += (unsigned long long)(clock64() - _t0)
#define TL_STAMP(slot) do { if (blockIdx.x < 160) g_tl[blockIdx.x * 8 + (slot)] = gtimer(); } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot)
#define TL_STAMP(slot)
#endif

// The persistent schedule: worker w owns the (M-block, tile) steps [begin[w], begin[w + 1]) of the
// M-block-major step sequence.  The host cuts the sequence by a cost model (make_plan), so the
// ranges travel as a kernel parameter.
constexpr int MAX_WORKERS = 160;
struct Sched {
    long long begin[MAX_WORKERS + 1];
};

// the worker whose range holds `step` (largest w with begin[w] <= step; ranges are never empty)
__host__ __device__ __forceinline__ int owner_of_step(const Sched &sc, int nworkers, long long step) {
    int lo = 0, hi = nworkers - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sc.begin[mid] <= step) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// A CTA's work range as a sequence of segments: runs of target tiles of one M-block.  Only the
// first segment can start in the middle of an M-block (and so be a later slot of it).
struct Segments {
    int mblock, tile_begin, remaining, slot, ntiles_row;
    __device__ __forceinline__ bool more() const { return remaining > 0; }
    __device__ __forceinline__ int ntiles() const {
        const int room = ntiles_row - tile_begin;
        return remaining < room ? remaining : room;
    }
    __device__ __forceinline__ void advance() {
        remaining -= ntiles();
        ++mblock;
        tile_begin = 0;
        slot = 0;
    }
};

struct RowState {
    int m1, i1, m2, i2;   // best / second-best partial distance (|t|^2 - 2 q.t) and target index
};

__device__ __forceinline__ int min3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ unsigned umin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u32(a, b, c); }

__device__ __forceinline__ int max16(const int *v) {
    const int a = max3(v[0], v[1], v[2]), b = max3(v[3], v[4], v[5]), c = max3(v[6], v[7], v[8]);
    const int d = max3(v[9], v[10], v[11]), e = max3(v[12], v[13], v[14]);
    return max3(max3(a, b, c), max3(d, e, v[15]), a);
}

// Exact update of a row's top-2 from 16 accumulator columns that passed the filter.
// key = partial * 256 + column (unique inside a tile, so integer order on keys is the
// lexicographic (distance, index) order).  `bound` <= s.m2, so a key below it enters the top-2;
// the chunk's runner-up only matters when the chunk also produced a new best.
__device__ __forceinline__ void slow16(const int *v, uint32_t ck_saddr, int jtile, int bound,
                                       RowState &s) {
    int k[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int4 c = ld_shared_v4(ck_saddr + i * 16);
        k[4 * i + 0] = c.x - 512 * v[4 * i + 0];
        k[4 * i + 1] = c.y - 512 * v[4 * i + 1];
        k[4 * i + 2] = c.z - 512 * v[4 * i + 2];
        k[4 * i + 3] = c.w - 512 * v[4 * i + 3];
    }
    const int a = min3(k[0], k[1], k[2]), b = min3(k[3], k[4], k[5]), c = min3(k[6], k[7], k[8]);
    const int d = min3(k[9], k[10], k[11]), e = min3(k[12], k[13], k[14]);
    const int kmin = min3(min3(a, b, c), min3(d, e, k[15]), a);
    const int p1 = kmin >> 8;
    if (p1 < bound) {
        const int j1 = jtile + (kmin & 255);
        if (p1 < s.m1) {
            // runner-up of the chunk: smallest key above kmin (keys are distinct)
            const int base = kmin + 1;
            unsigned u[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = (unsigned)(k[i] - base);   // kmin -> 0xFFFFFFFF
            const unsigned ua = umin3(u[0], u[1], u[2]), ub = umin3(u[3], u[4], u[5]);
            const unsigned uc = umin3(u[6], u[7], u[8]), ud = umin3(u[9], u[10], u[11]);
            const unsigned ue = umin3(u[12], u[13], u[14]);
            const unsigned um = umin3(umin3(ua, ub, uc), umin3(ud, ue, u[15]), ua);
            const int k2 = (int)(um + (unsigned)base);
            const int p2 = k2 >> 8;
            if (p2 < s.m1) { s.m2 = p2; s.i2 = jtile + (k2 & 255); }
            else { s.m2 = s.m1; s.i2 = s.i1; }
            s.m1 = p1; s.i1 = j1;
        } else {
            s.m2 = p1; s.i2 = j1;
        }
    }
}

// PAIR = false: one CTA per work range, M = 128 MMAs, two 256-column accumulators (one per
//                sub-tile), 4-stage ring of 256-target tiles.
// PAIR = true:  a cluster of two CTAs (one TPC) per work range.  The leader issues M = 256,
//                N = 128 MMAs (cta_group::2): each CTA contributes its own 128 query rows per
//                sub-tile and half of the 128 target rows, and receives a 128 x 128 accumulator.
//                That gives four accumulator buffers per CTA (sub-tile x column half) at the
//                shared-memory operand traffic of the N = 256 single-CTA instruction, so the
//                epilogue warps hold a buffer only while they read it.
template <bool PAIR>
struct Cfg {
    static constexpr int NBUF = PAIR ? 4 : 2;                       // TMEM accumulator buffers
    static constexpr int STAGES = PAIR ? 8 : 4;
    static constexpr int STAGE_BYTES = PAIR ? (B_BYTES + BX_BYTES) / 2 : (B_BYTES + BX_BYTES);
    static constexpr int MBLOCK_ROWS = PAIR ? 2 * SUBS * BM : SUBS * BM;
    static constexpr int HALF_B = 64 * FM_DIM;                      // PAIR: 64 target rows per column half
    static constexpr int HALF_X = 64 * 32;
    // Who returns a ring slot to the producer: a tcgen05.commit (costs the tensor pipe a ~50 clk
    // bubble while the epilogue reads TMEM) or the epilogue warp that sees the tile's last
    // accumulator complete.  Measured: the second is 5 % faster without PAIR, no gain with it.
    static constexpr bool EPILOGUE_FREES_SLOT = !PAIR;
};
static_assert(Cfg<true>::STAGES * Cfg<true>::STAGE_BYTES == STAGES * STAGE_BYTES, "same ring size");
static_assert(Cfg<true>::STAGES <= 8 && STAGES <= 8, "Bars holds 8 ring slots");

template <bool PAIR>
__device__ __forceinline__ void k_top2_body(const CUtensorMap &map_q, const CUtensorMap &map_t,
                                            const CUtensorMap &map_x, int64_t M, int64_t N,
                                            int32_t t_index_base, int ntiles_row, const Sched &sched,
                                            const int *__restrict__ qn, const int *__restrict__ ckey,
                                            int *__restrict__ gbound,
                                            unsigned long long *__restrict__ partial) {
    using C = Cfg<PAIR>;
#ifdef FM_TC_PROF
    const long long _tk0 = clock64();
    unsigned long long _pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (threadIdx.x == 0) TL_STAMP(0);
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Bars *bars = (Bars *)(smem + SMEM_BARS);
    unsigned long long *skeys = (unsigned long long *)(smem + SMEM_KEYS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;        // 0 = leader (issues the MMAs)
    // Persistent, stream-K style schedule: the work is the sequence of (M-block, target tile)
    // steps in M-block-major order; a worker (CTA, or CTA pair) owns a contiguous range of it,
    // i.e. a few *segments*, each a run of tiles of one M-block.  A segment's candidates go to
    // the partial-key slot `worker - first worker that touches the M-block`.
    const unsigned worker = PAIR ? blockIdx.x >> 1 : blockIdx.x;
    const unsigned nworkers = PAIR ? gridDim.x >> 1 : gridDim.x;
    Segments seg0;
    {
        const long long w_begin = sched.begin[worker], w_end = sched.begin[worker + 1];
        seg0.ntiles_row = ntiles_row;
        seg0.mblock = (int)(w_begin / ntiles_row);
        seg0.tile_begin = (int)(w_begin - (long long)seg0.mblock * ntiles_row);
        seg0.remaining = (int)(w_end - w_begin);          // < 2^31: checked by the host
        seg0.slot = seg0.tile_begin == 0 ? 0
            : (int)worker - owner_of_step(sched, (int)nworkers, w_begin - seg0.tile_begin);
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
        mbar_init(smem_u32(&bars->a_full), 1);
        mbar_init(smem_u32(&bars->a_empty), 1);
        for (int i = 0; i < C::NBUF; ++i) {
            mbar_init(smem_u32(&bars->tmem_full[i]), 1);
            // 8 epilogue warps release a buffer: one group here, or 4 warps in each CTA of the pair
            mbar_init(smem_u32(&bars->tmem_empty[i]), EPI_WARPS / SUBS);
        }
        fence_barrier_init();
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
        tma_prefetch_desc(&map_x);
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(smem_u32(&bars->tmem_base), TMEM_COLS);
        else tmem_alloc(smem_u32(&bars->tmem_base), TMEM_COLS);
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + (AX_BYTES / 16)) {
        // constant query-side block of the norm K-step: weights 255 (x15), 1 in both 16-B halves
        *(uint4 *)(smem + SMEM_AX + (threadIdx.x - 128) * 16) =
            make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x01FFFFFFu);
        fence_proxy_async();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // the peer's barriers must exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // everything above overlapped the pre-pass; its outputs (digits, key constants, norms, bounds)
    // are read from here on
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) TL_STAMP(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // PAIR: both CTAs load (their own query rows, their half of every target tile) and
            // every load signals the leader's barrier, which expects the bytes of both.
            const uint32_t abar = PAIR ? mapa_rank(smem_u32(&bars->a_full), 0) : smem_u32(&bars->a_full);
            int u = 0;      // tiles issued so far (ring position)
            auto load_b = [&](int tile, int uu) {
                const int stage = uu % C::STAGES;
                const uint32_t ph = (uu / C::STAGES) & 1;
                { PROF_T0(); mbar_wait(smem_u32(&bars->empty[stage]), ph ^ 1); PROF_ADD(2); }
                const uint32_t fb_local = smem_u32(&bars->full[stage]);
                uint8_t *st = smem + SMEM_B + stage * C::STAGE_BYTES;
                if (PAIR) {
                    if (rank == 0) mbar_expect_tx(fb_local, 2 * C::STAGE_BYTES);
                    const uint32_t fb = mapa_rank(fb_local, 0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // column half h of the tile = targets [h*128, h*128+128): 64 from each CTA
                        const int row = tile * BN + h * COLS_PER_WARP + (int)rank * 64;
                        tma_load_2d_pair(smem_u32(st + h * C::HALF_B), &map_t, 0, row, fb);
                        tma_load_2d_pair(smem_u32(st + 2 * C::HALF_B + h * C::HALF_X), &map_x, 0, row, fb);
                    }
                } else {
                    mbar_expect_tx(fb_local, B_BYTES + BX_BYTES);
                    tma_load_2d(smem_u32(st), &map_t, 0, tile * BN, fb_local);
                    tma_load_2d(smem_u32(st + B_BYTES), &map_x, 0, tile * BN, fb_local);
                }
            };
            Segments sg = seg0;
            int seg = 0;
            for (; sg.more(); ++seg, sg.advance()) {
                const int nt = sg.ntiles();
                // The first target tiles of a segment only need ring slots (freed by the previous
                // segment's MMAs), so they are requested before the single-buffered query tile,
                // which has to wait until the previous segment's last MMA is done.
                const int pre = nt < C::STAGES ? nt : C::STAGES;
                for (int it = 0; it < pre; ++it, ++u) load_b(sg.tile_begin + it, u);
                mbar_wait(smem_u32(&bars->a_empty), (seg & 1) ^ 1);
                if (PAIR) {
                    if (rank == 0) mbar_expect_tx(smem_u32(&bars->a_full), 2 * SUBS * A_BYTES);
                    for (int s = 0; s < SUBS; ++s)
                        tma_load_2d_pair(smem_u32(smem + SMEM_A + s * A_BYTES), &map_q, 0,
                                         sg.mblock * C::MBLOCK_ROWS + s * (2 * BM) + (int)rank * BM, abar);
                } else {
                    mbar_expect_tx(abar, SUBS * A_BYTES);
                    for (int s = 0; s < SUBS; ++s)
                        tma_load_2d(smem_u32(smem + SMEM_A + s * A_BYTES), &map_q, 0,
                                    sg.mblock * C::MBLOCK_ROWS + s * BM, abar);
                }
                for (int it = pre; it < nt; ++it, ++u) load_b(sg.tile_begin + it, u);
            }
            // tail: the last thing the MMA warp commits is a_empty of the last segment; once it is
            // here, no multicast arrive is still on its way to this CTA's barriers
            if (PAIR) mbar_wait(smem_u32(&bars->a_empty), (seg & 1) ^ 1);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (PAIR: the leader CTA only) =====================
        // The whole warp walks the loop (so every address below is warp-uniform and lives in
        // uniform registers); one elected lane issues.  The issue path must stay short: with
        // four busy epilogue warps on the same scheduler a long one cannot keep up with 64-clk MMAs.
        if (rank == 0) {
            constexpr uint32_t idesc = PAIR ? make_idesc(2 * BM, COLS_PER_WARP) : make_idesc(BM, BN);
            const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t axdesc = make_desc_sw32(sbase + SMEM_AX);
            const uint64_t adesc0 = make_desc(sbase + SMEM_A);
            auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            int u = 0;
            Segments sg = seg0;
            for (int seg = 0; sg.more(); ++seg, sg.advance()) {
                { PROF_T0(); mbar_wait(smem_u32(&bars->a_full), seg & 1); PROF_ADD(3); }
                tc_fence_after();
                const int nt = sg.ntiles();
                for (int it = 0; it < nt; ++it, ++u) {
                    const int stage = u % C::STAGES;
                    const uint32_t ph = (u / C::STAGES) & 1;
                    { PROF_T0(); mbar_wait(smem_u32(&bars->full[stage]), ph); PROF_ADD(0); }
                    tc_fence_after();
#ifdef FM_TC_PROF
                    if (u == 0 && lane == 0) TL_STAMP(2);
#endif
                    const uint32_t sb = sbase + SMEM_B + stage * C::STAGE_BYTES;
#pragma unroll
                    for (int b = 0; b < C::NBUF; ++b) {
                        const int s = PAIR ? b >> 1 : b, h = PAIR ? b & 1 : 0;
                        { PROF_T0(); mbar_wait(smem_u32(&bars->tmem_empty[b]), (u & 1) ^ 1); PROF_ADD(1); }
                        tc_fence_after();
                        const uint64_t adesc = adesc0 + (uint64_t)(s * (A_BYTES >> 4));
                        const uint64_t bdesc = make_desc(PAIR ? sb + h * C::HALF_B : sb);
                        const uint64_t bxdesc = make_desc_sw32(PAIR ? sb + 2 * C::HALF_B + h * C::HALF_X : sb + B_BYTES);
                        const uint32_t d = tbase + s * BN + h * COLS_PER_WARP;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < FM_DIM / 32; ++k) {
                                if (PAIR) umma_i8_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                                else umma_i8(d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                            }
                            if (PAIR) umma_i8_pair(d, axdesc, bxdesc, idesc, 1);     // + E_j
                            else umma_i8(d, axdesc, bxdesc, idesc, 1);
#ifdef FM_PAIR_COMMIT_PER_SUB   /* variant: one "accumulator ready" signal per sub-tile (both column halves) */
                            if (!PAIR || h == 1) commit(smem_u32(&bars->tmem_full[PAIR ? s * 2 : b]));
#else
                            commit(smem_u32(&bars->tmem_full[b]));
#endif
                        }
                        __syncwarp();
                    }
                    if (!C::EPILOGUE_FREES_SLOT) {
                        if (elect_one()) commit(smem_u32(&bars->empty[stage]));
                        __syncwarp();
                    }
#ifdef FM_TC_PROF
                    if (lane == 0) atomicAdd(&g_prof[11], 1ull);
#endif
                }
                if (elect_one()) commit(smem_u32(&bars->a_empty));      // the query tile may be replaced
                __syncwarp();
            }
#ifdef FM_TC_PROF
            if (lane == 0) TL_STAMP(3);
#endif
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // A warp owns 32 rows (its TMEM lane quarter) x 128 columns (a column half) of one
        // sub-tile and sweeps them in two passes of 64 columns; the accumulator goes back to the
        // MMA warp as soon as the second pass has been read.  Without PAIR the 8 warps of a
        // sub-tile share one 256-column buffer and the two groups run half a period apart; with
        // PAIR every (sub-tile, column half) is a buffer of its own.
        const int ew = warp - 4;
        const int s = ew >> 3;               // sub-tile
        const int lq = warp & 3;             // TMEM lane quarter this warp may touch
        const int ch = (ew >> 2) & 1;        // column half
        const int bi = PAIR ? s * 2 + ch : s;
        const int row_in_sub = lq * 32 + lane;
        const bool releases_stage = s == SUBS - 1 && lq == 0 && (PAIR ? ch == 1 : ch == 0);
        constexpr int cg = CG;
        // sm2[row]: (second-best partial distance + 1) published by the two warps that sweep the
        // same row (and by other CTAs, below) -- "+1" because a sibling's candidate may carry a
        // higher index (non-strict bound).
        const uint32_t sm2_a = smem_u32(smem + SMEM_M2) + (s * BM + row_in_sub) * 4;
        // exact-key constants of this warp's 128 columns: warp-private, double-buffered in smem
        const uint32_t ck_a = smem_u32(smem + SMEM_CK) + ew * (2 * COLS_PER_WARP * 4);
        // per-warp TMEM address of its 32 lanes x 128 columns (warp-uniform)
        const uint32_t taddr0 = __shfl_sync(0xffffffffu, tmem_base + ((uint32_t)(lq * 32) << 16) + s * BN + ch * COLS_PER_WARP, 0);
#ifdef FM_PAIR_COMMIT_PER_SUB
        const uint32_t full_a = smem_u32(&bars->tmem_full[PAIR ? s * 2 : bi]);
#else
        const uint32_t full_a = smem_u32(&bars->tmem_full[bi]);
#endif
        const uint32_t empty_a = PAIR ? mapa_rank(smem_u32(&bars->tmem_empty[bi]), 0) : smem_u32(&bars->tmem_empty[bi]);
        int u = 0;          // tiles consumed so far (barrier phase)
        for (Segments sg = seg0; sg.more(); sg.advance()) {
        const int tile_begin = sg.tile_begin, ntiles = sg.ntiles();
        RowState st;
        st.m1 = st.m2 = NONE_P; st.i1 = st.i2 = -1;
        if (ch == 0) st_shared_s32(sm2_a, NONE_P);
        // (also orders the previous segment's reads of the exchange area before this segment's writes)
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        const int *ckg = ckey + (int64_t)tile_begin * BN + ch * COLS_PER_WARP + lane * 4;
        cp_async16(ck_a + lane * 16, ckg);
        cp_async_commit();
        const int64_t grow = (int64_t)sg.mblock * C::MBLOCK_ROWS +
                             (PAIR ? s * (2 * BM) + (int)rank * BM : s * BM) + row_in_sub;
        int gnext = NONE_P, gpub = NONE_P;

        for (int it = 0; it < ntiles; ++it, ++u) {
            const int jtile = (tile_begin + it) * BN;
            if (it + 1 < ntiles)
                cp_async16(ck_a + ((it + 1) & 1) * (COLS_PER_WARP * 4) + lane * 16, ckg + (int64_t)(it + 1) * BN);
            cp_async_commit();
            cp_async_wait1();
            __syncwarp();
            const uint32_t ck = ck_a + (it & 1) * (COLS_PER_WARP * 4);
            if (gbound != nullptr && ch == 0) {
                // Bounds also travel between the CTAs that sweep other target slices for the same
                // rows (global memory, every 4th tile, same non-strict "+1" convention): read one
                // tile ahead of use, publish with a fire-and-forget reduction.
                const int ph = it & 3;
                if (ph == 1) red_shared_min_s32(sm2_a, gnext);
                else if (ph == 0) gnext = ld_global_relaxed(gbound + grow);
                else if (ph == 2) {
                    const int v = ld_shared_s32(sm2_a);
                    if (v < gpub) { red_global_min_s32(gbound + grow, v); gpub = v; }
                }
            }
            const int shared_b = ld_shared_s32(sm2_a);
#ifdef FM_TC_PROF
            const long long _te0 = clock64();
#endif
            mbar_wait(full_a, u & 1);
#ifdef FM_TC_PROF
            const long long _te1 = clock64();
            if (u == 0 && warp == 4 && lane == 0) TL_STAMP(4);
#endif
            tc_fence_after();
            // The last accumulator of a tile being complete means every MMA that read the tile's
            // ring slot is done: one warp hands the slot back to the producer.
            if (C::EPILOGUE_FREES_SLOT && releases_stage && lane == 0) mbar_arrive(smem_u32(&bars->empty[u % C::STAGES]));
            int bound = min(st.m2, shared_b);
            int thr = (cg - bound) >> 1;        // acc' > thr  <=>  C - 2 acc' < bound
            // The sweep of one accumulator, in two flavours: per-16-column filter branches while the
            // rows' bounds are still loose (first tiles of a segment: most chunks pass), one branch
            // per 32 columns once they are tight (the exact update stays per 16 columns).
            auto sweep = [&](auto warm) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                int v0[32], v1[32];
                tmem_ld32(taddr0 + pass * 64, v0);
#ifndef FM_EXPERIMENT_HALF_LD
                tmem_ld32(taddr0 + pass * 64 + 32, v1);
#endif
                tmem_ld_wait();
#ifdef FM_EXPERIMENT_EARLY_RELEASE   /* timing experiment only (results are wrong): the accumulator goes
                                        back before it is read, so the MMA warp never waits for the epilogue */
                if (pass == 0 && lane == 0) { if (PAIR) mbar_arrive_cluster(empty_a); else mbar_arrive(empty_a); }
#else
                if (pass == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_cluster(empty_a); else mbar_arrive(empty_a); }
                }
#endif
#ifdef FM_EXPERIMENT_NO_SLOW   /* timing experiment only: results are wrong */
#define FM_TRIG(x) ((x) > 0x7FFFFFF0)
#else
#define FM_TRIG(x) ((x) > thr)
#endif
#define FM_CHUNK(V, COL)                                                                  \
                if (FM_TRIG(max16(V))) {                                                     \
                    const int m2_before = st.m2;                                             \
                    slow16(V, ck + (pass * 64 + (COL)) * 4, jtile, bound, st);               \
                    if (st.m2 < m2_before) {                                                 \
                        red_shared_min_s32(sm2_a, st.m2 + 1);                                \
                        bound = min(st.m2, bound);                                           \
                        thr = (cg - bound) >> 1;                                             \
                    }                                                                        \
                }
#if defined(FM_EXPERIMENT_NO_ALU)      /* timing experiment only: results are wrong */
                if ((v0[0] ^ v0[31] ^ v1[0] ^ v1[31]) == 0x7FFFFFF1) st.m2 = 0;
#elif defined(FM_EXPERIMENT_HALF_LD)   /* timing experiment only: results are wrong */
                FM_CHUNK(v0, 0) FM_CHUNK(v0 + 16, 16)
#else
#define FM_SLOW(V, COL, MX)                                                               \
                if (FM_TRIG(MX)) {                                                           \
                    const int m2_before = st.m2;                                             \
                    slow16(V, ck + (pass * 64 + (COL)) * 4, jtile, bound, st);               \
                    if (st.m2 < m2_before) {                                                 \
                        red_shared_min_s32(sm2_a, st.m2 + 1);                                \
                        bound = min(st.m2, bound);                                           \
                        thr = (cg - bound) >> 1;                                             \
                    }                                                                        \
                }
                if constexpr (decltype(warm)::value) {
                    // warm rows (most chunks fail the filter): one branch per 32 columns
                    const int ma = max16(v0), mb = max16(v0 + 16);
                    if (FM_TRIG(max(ma, mb))) { FM_SLOW(v0, 0, ma) FM_SLOW(v0 + 16, 16, mb) }
                    const int mc = max16(v1), md = max16(v1 + 16);
                    if (FM_TRIG(max(mc, md))) { FM_SLOW(v1, 32, mc) FM_SLOW(v1 + 16, 48, md) }
                } else {
                    FM_CHUNK(v0, 0) FM_CHUNK(v0 + 16, 16) FM_CHUNK(v1, 32) FM_CHUNK(v1 + 16, 48)
                }
#undef FM_SLOW
#endif
#undef FM_CHUNK
#undef FM_TRIG
            }
            };
            // Measured (tools/variant_batch.py, round 2): one branch per 32 columns everywhere is
            // as fast as the per-16 form at 50k x 50k and 2 % faster at 200k x 200k; switching
            // flavour after the first 16 / 32 / 64 tiles of a segment (FM_VARIANT_HYBRID) sits in between.
#ifdef FM_VARIANT_HYBRID
            if (it >= FM_VARIANT_HYBRID) sweep(std::true_type{}); else sweep(std::false_type{});
#elif defined(FM_VARIANT_PER16)
            sweep(std::false_type{});
#else
            sweep(std::true_type{});
#endif
#ifdef FM_TC_PROF
            if (lane == 0) {
                const long long _te3 = clock64();
                _pacc[4] += (unsigned long long)(_te1 - _te0);   // wait tmem_full
                _pacc[6] += (unsigned long long)(_te3 - _te1);   // loads + filter + updates
                _pacc[7] += 1ull;
            }
#endif
        }
        // ---- merge the 2 column halves of every row through shared memory
        {
            const int r = s * BM + row_in_sub;
            const int qnr = grow < M ? __ldg(qn + grow) : 0;
            unsigned long long k1 = st.i1 < 0 ? FM_NONE_KEY
                : pack_key((uint32_t)(st.m1 + qnr), (uint32_t)(st.i1 + t_index_base));
            unsigned long long k2 = st.i2 < 0 ? FM_NONE_KEY
                : pack_key((uint32_t)(st.m2 + qnr), (uint32_t)(st.i2 + t_index_base));
            skeys[(r * 2 + ch) * 2] = k1;
            skeys[(r * 2 + ch) * 2 + 1] = k2;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (ch == 0) {
            const int r = s * BM + row_in_sub;
            unsigned long long a = skeys[r * 4], b = skeys[r * 4 + 1];
            merge2(a, b, skeys[r * 4 + 2], skeys[r * 4 + 3]);
            if (grow < M) {
                partial[((int64_t)sg.slot * M + grow) * 2] = a;
                partial[((int64_t)sg.slot * M + grow) * 2 + 1] = b;
            }
        }
        }   // segments
#ifdef FM_TC_PROF
        if (warp == 4 && lane == 0) TL_STAMP(5);
#endif
    }

    tc_fence_before();
    // PAIR: neither CTA may leave (or free its TMEM) while the peer can still signal its barriers
    // or the tensor core can still read its shared memory
    if (PAIR) cluster_sync_all();
    else __syncthreads();
#ifdef FM_TC_PROF
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 8; ++i) if (_pacc[i]) atomicAdd(&g_prof[i], _pacc[i]);
    if (threadIdx.x == 0) { atomicAdd(&g_prof[8], (unsigned long long)(clock64() - _tk0)); atomicAdd(&g_prof[9], 1ull); atomicAdd(&g_prof[10], (unsigned long long)seg0.remaining); TL_STAMP(6); }
#endif
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_top2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
          const __grid_constant__ CUtensorMap map_x, int64_t M, int64_t N, int32_t t_index_base,
          int ntiles_row, const __grid_constant__ Sched sched, const int *__restrict__ qn,
          const int *__restrict__ ckey, int *__restrict__ gbound,
          unsigned long long *__restrict__ partial) {
    k_top2_body<false>(map_q, map_t, map_x, M, N, t_index_base, ntiles_row, sched, qn, ckey, gbound, partial);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
k_top2_tc_pair(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
               const __grid_constant__ CUtensorMap map_x, int64_t M, int64_t N, int32_t t_index_base,
               int ntiles_row, const __grid_constant__ Sched sched, const int *__restrict__ qn,
               const int *__restrict__ ckey, int *__restrict__ gbound,
               unsigned long long *__restrict__ partial) {
    k_top2_body<true>(map_q, map_t, map_x, M, N, t_index_base, ntiles_row, sched, qn, ckey, gbound, partial);
}

// merge of the partial keys the CTAs that swept one M-block left in its slots (same semantics as
// fm_merge_top2); the slot count of a block follows from the schedule.
__global__ void k_merge_partial(const unsigned long long *__restrict__ partial, int ntiles_row,
                                const __grid_constant__ Sched sched, int nworkers, int mblock_rows, int64_t M,
                                uint32_t *__restrict__ d2, int32_t *__restrict__ idx,
                                unsigned long long *__restrict__ keys, const RatioOut rout) {
    pdl_wait();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const long long first = (i / mblock_rows) * ntiles_row;
    const int slots = owner_of_step(sched, nworkers, first + ntiles_row - 1) - owner_of_step(sched, nworkers, first) + 1;
    unsigned long long a = FM_NONE_KEY, b = FM_NONE_KEY;
    for (int s = 0; s < slots; ++s) {
        const ulonglong2 v = *(const ulonglong2 *)(partial + ((int64_t)s * M + i) * 2);
        insert2(v.x, a, b);
        insert2(v.y, a, b);
    }
    d2[2 * i] = (uint32_t)(a >> 32);
    d2[2 * i + 1] = (uint32_t)(b >> 32);
    idx[2 * i] = a == FM_NONE_KEY ? -1 : (int32_t)(uint32_t)a;
    idx[2 * i + 1] = b == FM_NONE_KEY ? -1 : (int32_t)(uint32_t)b;
    if (keys) { keys[2 * i] = a; keys[2 * i + 1] = b; }
    write_ratio(rout, i, (uint32_t)(a >> 32), (uint32_t)(b >> 32));
}

// all slots missing (no targets): the ratio test sees +inf
__global__ void k_ratio_none(int64_t M, const RatioOut rout) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < M) write_ratio(rout, i, FM_NONE_D2, FM_NONE_D2);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Plan {
    bool pair;                  // CTA-pair kernel (cta_group::2)
    int mblock_rows;            // query rows per M-block: 256, or 512 for a pair
    int64_t mblocks, mpad, ntiles, npad, work;
    int workers, slots;         // CTAs (or pairs) in the persistent grid; partial-key slots per row
    int aligned;                // whole M-blocks per worker (targets larger than L2)
    Sched sched;                // first step of every worker
    size_t off_ckey, off_digits, off_qn, off_gbound, off_partial, total;
};

// Per-device facts and one-time set-up, keyed by the current device: a process may drive several
// GPUs, and function attributes / SM counts / cluster support are per device.
struct DevInfo {
    bool init = false, attrs_set = false;
    int sms = 148, major = 0, pair_ok = 0;
    size_t l2 = 64u << 20;
};
constexpr int MAX_DEVICES = 64;
static DevInfo &dev_info() {
    static DevInfo info[MAX_DEVICES];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
    DevInfo &d = info[dev];
    std::lock_guard<std::mutex> lock(mu);
    if (!d.init) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) d.sms = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev) == cudaSuccess && v > 0) d.l2 = (size_t)v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess) d.major = v;
        // Can a 2-CTA cluster of the pair kernel be resident at all (it cannot on e.g. a MIG slice
        // with single-SM TPCs)?  "No" or any error selects the single-CTA kernel.
        if (d.major == 10 &&
            cudaFuncSetAttribute(k_top2_tc_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC) == cudaSuccess &&
            cudaFuncSetAttribute(k_top2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC) == cudaSuccess) {
            d.attrs_set = true;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2);
            cfg.blockDim = dim3(NTHREADS);
            cfg.dynamicSmemBytes = SMEM_ALLOC;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_top2_tc_pair, &cfg) == cudaSuccess && n > 0) d.pair_ok = 1;
        }
        (void)cudaGetLastError();
        d.init = true;
    }
    return d;
}

// FM_TC_PAIR=0 / 1 forces the single-CTA / CTA-pair kernel (for A/B timing and tests); default: pair
// whenever a 512-row M-block is not mostly padding.  FM_TC_ALIGNED=0 / 1 likewise forces the
// stream-K / whole-M-block schedule.
static int pair_override() {
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("FM_TC_PAIR");
        v = e && *e ? atoi(e) : -1;
    }
    return v;
}

// Cost model of the stream-K cut (measured with tools/timeline.py): besides one unit per tile step a
// worker pays for every exact update of a row's top-2 ("event", ~0.1 tile steps each), and those
// concentrate where a row is swept first.  A block's steps that run at the start of a worker's range
// -- the tail piece [x, L) of a split block, or a whole block -- start from no bound at all:
//   events(s tiles) = 8 s for s < 4, else 32 + 64 ln(s / 4)   (per warp: 32 rows x 2 ln(targets))
// while the head piece [0, x) of a split block runs at the very end of the previous worker's range,
// seeded through `gbound` by the tail piece that ran first, and only pays the remainder.  So the cost
// of the step sequence up to tile x of a block is  head(x) = block_cost - tail_cost(L - x), and the
// cuts are placed at equal increments of that cumulative cost.
static double tail_cost(double s, double evcost) {
    if (s <= 0) return 0;
    return s + evcost * (s < 4 ? 8 * s : 32 + 64 * log(s / 4));
}

static void cut_schedule(Plan &p, double evcost) {
    const int W = p.workers;
    const long long L = p.ntiles;
    if (p.aligned) {
        for (int w = 0; w <= W; ++w) p.sched.begin[w] = p.mblocks * w / W * L;
    } else {
        const double cb = tail_cost((double)L, evcost), total = cb * (double)p.mblocks;
        p.sched.begin[0] = 0;
        for (int w = 1; w < W; ++w) {
            const double c = total * w / W;
            long long blk = (long long)(c / cb);
            if (blk >= p.mblocks) blk = p.mblocks - 1;
            const double rem = c - blk * cb;
            long long lo = 0, hi = L;           // smallest x with head(x) >= rem
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (cb - tail_cost((double)(L - mid), evcost) >= rem) hi = mid; else lo = mid + 1;
            }
            p.sched.begin[w] = blk * L + lo;
        }
        p.sched.begin[W] = p.work;
        for (int w = 1; w < W; ++w)             // never an empty range
            if (p.sched.begin[w] <= p.sched.begin[w - 1]) p.sched.begin[w] = p.sched.begin[w - 1] + 1;
        for (int w = W - 1; w >= 1; --w)
            if (p.sched.begin[w] >= p.sched.begin[w + 1]) p.sched.begin[w] = p.sched.begin[w + 1] - 1;
    }
    for (int w = W + 1; w <= MAX_WORKERS; ++w) p.sched.begin[w] = p.work;
    // partial-key slots: only a worker's first run can start inside an M-block
    int slots = 1;
    for (int w = 1; w < W; ++w) {
        const long long b = p.sched.begin[w];
        if (b % L == 0) continue;
        const int slot = w - owner_of_step(p.sched, W, b - b % L);
        if (slot + 1 > slots) slots = slot + 1;
    }
    p.slots = slots;
}

static Plan make_plan(int64_t M, int64_t N) {
    Plan p;
    const DevInfo &dv = dev_info();
    const int ov = pair_override();
    // pair by default unless its 512-row M-blocks would add a (relatively) large block of padding
    p.pair = (ov >= 0 ? ov != 0 : (M > 8 * SUBS * BM || (M > SUBS * BM && (M - 1) % (2 * SUBS * BM) >= SUBS * BM))) &&
             dv.pair_ok;
    p.mblock_rows = p.pair ? 2 * SUBS * BM : SUBS * BM;
    p.mblocks = (M + p.mblock_rows - 1) / p.mblock_rows;
    p.mpad = p.mblocks * p.mblock_rows;
    p.ntiles = (N + BN - 1) / BN;
    p.npad = p.ntiles * BN;
    // persistent grid: every worker gets a contiguous share of the (M-block, tile) steps
    p.work = p.mblocks * p.ntiles;
    int cap = p.pair ? dv.sms / 2 : dv.sms;
    if (cap > MAX_WORKERS) cap = MAX_WORKERS;
    p.workers = (int)(p.work < cap ? p.work : cap);
    if (p.workers < 1) p.workers = 1;
    // Targets (+ digits) that do not fit L2 are streamed from HBM once per M-block unless the workers
    // sweep them in step: give every worker whole M-blocks then (<= 1/8 imbalance by the condition).
    static const int aligned_override = [] { const char *e = getenv("FM_TC_ALIGNED"); return e && *e ? atoi(e) : -1; }();
    p.aligned = aligned_override >= 0 ? aligned_override != 0
        : (size_t)p.npad * (FM_DIM + 32) > dv.l2 / 2 && p.mblocks >= 8 * (int64_t)p.workers;
    if (p.aligned && p.mblocks < p.workers) p.aligned = 0;       // (forced by the override on a small problem)
    static const double evcost = [] { const char *e = getenv("FM_TC_EVCOST"); return e && *e ? atof(e) : 0.15; }();
    cut_schedule(p, evcost);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    p.off_ckey = 0;
    p.off_digits = up(p.off_ckey + (size_t)p.npad * 4);
    p.off_qn = up(p.off_digits + (size_t)p.npad * 32);
    p.off_gbound = up(p.off_qn + (size_t)p.mpad * 4);
    p.off_partial = up(p.off_gbound + (size_t)p.mpad * 4);
    p.total = up(p.off_partial + (size_t)p.slots * M * 16);
    return p;
}

}  // namespace tc

#ifdef FM_TC_PROF
extern "C" int fm_debug_prof(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, tc::g_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_prof, z, sizeof(z)); }
    return 0;
}
// globaltimer stamps of the last launch: out[cta * 8 + phase], 160 CTAs
extern "C" int fm_debug_timeline(unsigned long long *out1280) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out1280, tc::g_tl, sizeof(unsigned long long) * 160 * 8) == cudaSuccess ? 0 : -1;
}
#endif

// Test hook (not part of include/fastmatch_b200.h): the schedule the dense tcgen05 kernel would use.
// out = {pair, mblock_rows, mblocks, ntiles, workers, slots, aligned, workspace bytes >> 10}
extern "C" int fm_debug_plan(int64_t M, int64_t N, long long *out8) {
    if (M <= 0 || N <= 0 || out8 == nullptr) return FM_EINVAL;
    const tc::Plan p = tc::make_plan(M, N);
    out8[0] = p.pair; out8[1] = p.mblock_rows; out8[2] = p.mblocks; out8[3] = p.ntiles;
    out8[4] = p.workers; out8[5] = p.slots; out8[6] = p.aligned; out8[7] = (long long)(p.total >> 10);
    return FM_OK;
}

// Test hook: first step of every worker of that schedule; out[0..workers] (cap >= workers + 1).
extern "C" int fm_debug_plan_begins(int64_t M, int64_t N, long long *out, int cap) {
    if (M <= 0 || N <= 0 || out == nullptr) return FM_EINVAL;
    const tc::Plan p = tc::make_plan(M, N);
    if (cap < p.workers + 1) return FM_ENOSPACE;
    for (int w = 0; w <= p.workers; ++w) out[w] = p.sched.begin[w];
    return FM_OK;
}

bool tc_supported() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return tc::dev_info().major == 10 && tc::encode_fn() != nullptr;
}

size_t top2_tc_workspace_bytes(int64_t M, int64_t N) {
    if (M <= 0 || N <= 0) return 0;
    return tc::make_plan(M, N).total;
}

int launch_top2_tc(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, int32_t t_index_base,
                   uint32_t *d2, int32_t *idx, uint64_t *keys, RatioOut rout, void *ws,
                   size_t ws_bytes, cudaStream_t s) {
    using namespace tc;
    if (N == 0) {  // nothing to match against: every slot is missing
        FM_CUDA_TRY(cudaMemsetAsync(d2, 0xFF, (size_t)M * 8, s));
        FM_CUDA_TRY(cudaMemsetAsync(idx, 0xFF, (size_t)M * 8, s));
        if (keys) FM_CUDA_TRY(cudaMemsetAsync(keys, 0xFF, (size_t)M * 16, s));
        if (rout.ratio || rout.mask) {
            k_ratio_none<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(M, rout);
            FM_CUDA_TRY(cudaGetLastError());
            count_launch();
        }
        return FM_OK;
    }
    const Plan p = make_plan(M, N);
    if (ws_bytes < p.total) { set_error("tcgen05 path: workspace too small"); return FM_ENOSPACE; }
    for (int w = 0; w < p.workers; ++w)
        if (p.sched.begin[w + 1] - p.sched.begin[w] >= (long long)0x7FFFFFF0) {
            set_error("tcgen05 path: M x N too large for one launch, shard it");
            return FM_EINVAL;
        }
    uint8_t *w = (uint8_t *)ws;
    int *ckey = (int *)(w + p.off_ckey), *qn = (int *)(w + p.off_qn);
    uint8_t *digits = w + p.off_digits;
    unsigned long long *partial = (unsigned long long *)(w + p.off_partial);

    // a pair loads every target tile as four 64-row boxes (column half x CTA)
    const int tbox = p.pair ? 64 : BN;
    CUtensorMap map_q, map_t, map_x;
    int rc;
    if ((rc = make_map(&map_q, q, M, FM_DIM, BM, CU_TENSOR_MAP_SWIZZLE_128B)) != FM_OK) return rc;
    if ((rc = make_map(&map_t, t, N, FM_DIM, tbox, CU_TENSOR_MAP_SWIZZLE_128B)) != FM_OK) return rc;
    if ((rc = make_map(&map_x, digits, N, 32, tbox, CU_TENSOR_MAP_SWIZZLE_NONE)) != FM_OK) return rc;

    int *gbound = p.slots > 1 ? (int *)(w + p.off_gbound) : nullptr;
    k_prepass<<<(unsigned)(((p.mpad + p.npad) * 8 + 255) / 256), 256, 0, s>>>(q, M, p.mpad, t, N, p.npad, qn,
                                                                              gbound, ckey, (uint4 *)digits);
    FM_CUDA_TRY(cudaGetLastError());
    count_launch();

    if (!dev_info().attrs_set) { set_error("tcgen05 path: could not opt in to %d bytes of shared memory", SMEM_ALLOC); return FM_ECUDA; }
    // FM_TC_PDL=0 launches the three kernels back to back without programmatic dependent launch
    static const bool use_pdl = [] { const char *e = getenv("FM_TC_PDL"); return !(e && *e && atoi(e) == 0); }();
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = s;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_ALLOC;
    const int ntiles_i = (int)p.ntiles;
    prof_begin(s);
    if (p.pair) {
        cfg.gridDim = dim3(2 * p.workers);
        FM_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_top2_tc_pair, map_q, map_t, map_x, M, N, t_index_base, ntiles_i, p.sched,
                                       (const int *)qn, (const int *)ckey, gbound, partial));
    } else {
        cfg.gridDim = dim3(p.workers);
        FM_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_top2_tc, map_q, map_t, map_x, M, N, t_index_base, ntiles_i, p.sched,
                                       (const int *)qn, (const int *)ckey, gbound, partial));
    }
    prof_end(s);
    count_launch();
    cfg.gridDim = dim3((unsigned)((M + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    FM_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_merge_partial, (const unsigned long long *)partial, ntiles_i, p.sched,
                                   p.workers, p.mblock_rows, M, d2, idx, (unsigned long long *)keys, rout));
    count_launch();
    return FM_OK;
}

}  // namespace fm
