// fm_grouped_tc.cu -- thousands of small mutual-nearest-neighbour rounds in ONE launch on the
// tcgen05 tensor cores (sm_100a).
//
// Replaces the per-round cv2.BFMatcher(NORM_L2, crossCheck=True).knnMatch(query_ds, target_ds, k=1)
// of fastmatch.pyx:161-162 (match_position, one call per flood-fill round) and :122-123
// (match_thumbs).  A round is tiny (tens to hundreds of descriptors per side), so the path is
// launch- and HBM-bound: every descriptor is read once (TMA, 32-row boxes = only the rows a round
// needs), the distance tile lives in TMEM, and only (d2, index) pairs go back to HBM.
//
// Persistent kernel, one CTA per SM; groups are claimed from a global counter (sizes vary 256x).
// For a group with nq queries and nt targets the CTA runs two passes of "units"
//   pass 0: A = 128-query slab, B = up to 256 targets  -> per query the two smallest (d2, idx)
//   pass 1: A = 128-target slab, B = up to 256 queries -> per target the nearest query (crossCheck)
// so both directions are a per-row reduction over columns (no cross-lane reductions, no atomics);
// the second pass costs tensor time only, which this path has to spare.  Units flow through the
// same three pipelines as the dense kernel: TMA ring (A 16 KB + B 32 KB per stage) -> single-thread
// tcgen05.mma kind::i8 (M = 128, N = round_up(valid B rows, 16)) -> two ping-pong 256-column
// TMEM accumulators -> 16 epilogue warps.  The epilogue forms exact integer keys
//   key = (|b_j|^2 - 2 a_i.b_j) * 256 + column      (one IMAD; |b_j|^2*256+column comes from smem)
// and keeps the running minimum (pass 1) or two minima (pass 0) per row; integer order on keys is
// the lexicographic (distance, index) order, chunks are visited in increasing index with strict
// "<", so ties go to the lowest local index exactly as cv::batchDistance does.
#include <cuda.h>

#include <mutex>

#include "fm_common.cuh"
#include "fm_tc_ptx.cuh"

namespace fm {

namespace gtc {

using namespace tc;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;
constexpr int COLS_PER_WARP = BN / 2;    // a warp sweeps one column half of a unit
constexpr int A_BYTES = BM * FM_DIM, B_BYTES = BN * FM_DIM, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BOX_ROWS = 32, BOX_BYTES = BOX_ROWS * FM_DIM;
constexpr int RING = 8;
constexpr int NORM_WARPS = 2;            // warps 2 and 3
constexpr int I32_MAX = 0x7FFFFFFF;
constexpr int NONE_P = 0x7FFFFF;

enum { F_PASS1 = 1, F_FIRST = 2, F_LAST = 4, F_STOP = 8 };

struct __align__(16) Slot {
    int a_row0, a_valid, b_row0, b_valid, a_out0, b_local0, flags, n_mma;
    int ck[BN];            // |b_c|^2 * 256 + c for the unit's B rows (written by the norm warps)
    int an[BM];            // |a_i|^2 of the unit's A rows (only for the last unit of a slab)
};

struct __align__(8) Bars {
    unsigned long long full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    unsigned long long norm_full[2 * RING];   // per ring slot: its ck / an entries are in place
    uint32_t tmem_base, pad;
    int stage_slot[STAGES + 1];     // ring slot of the unit that travels in each TMA stage
};

constexpr int SMEM_STAGES = 0;
constexpr int SMEM_RING = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_KEYS = SMEM_RING + 2 * RING * (int)sizeof(Slot);   // [2 groups][2 bufs][128 rows][2 halves][2] u64
constexpr int SMEM_BARS = SMEM_KEYS + 2 * 2 * BM * 2 * 2 * 8;
constexpr int SMEM_TOTAL = SMEM_BARS + (int)sizeof(Bars);
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;

struct Params {
    const int64_t *q_off, *t_off, *t_base;   // t_base nullable
    int32_t G;
    int *counter;
    uint32_t *q2t_d2;
    int32_t *q2t_idx, *t2q_idx;
};

#ifdef FM_TC_PROF
__device__ unsigned long long g_gprof[16];
#define GP_T(var) const long long var = clock64()
#define GP_ACC(slot, a, b) _gp[slot] += (unsigned long long)((b) - (a))
#else
#define GP_T(var)
#define GP_ACC(slot, a, b)
#endif

__device__ __forceinline__ int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ unsigned umin3i(unsigned a, unsigned b, unsigned c) { return __vimin3_u32(a, b, c); }

// minimum of 32 keys (3-input tree, 16 ops)
__device__ __forceinline__ int min32(const int (&v)[32]) {
    int t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = min3i(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    t[10] = min(v[30], v[31]);
    const int a = min3i(t[0], t[1], t[2]), b = min3i(t[3], t[4], t[5]), c = min3i(t[6], t[7], t[8]);
    return min3i(min3i(a, b, c), t[9], t[10]);
}
// unsigned minimum of key - base over 32 keys (the key equal to base - 1 wraps to the maximum)
__device__ __forceinline__ unsigned umin32(const int (&v)[32], int base) {
#ifdef FM_GROUPED_UMIN_TREE
    unsigned t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        t[i] = umin3i((unsigned)(v[3 * i] - base), (unsigned)(v[3 * i + 1] - base), (unsigned)(v[3 * i + 2] - base));
    t[10] = min((unsigned)(v[30] - base), (unsigned)(v[31] - base));
    const unsigned a = umin3i(t[0], t[1], t[2]), b = umin3i(t[3], t[4], t[5]), c = umin3i(t[6], t[7], t[8]);
    return umin3i(umin3i(a, b, c), t[9], t[10]);
#else
    // fused add + min (VIADDMNMX.U32: one ALU-pipe op per key instead of a subtraction plus half a
    // 3-input min); four independent chains keep the dependent-issue latency out of the way
    const unsigned nb = (unsigned)(-base);
    unsigned r0 = 0xFFFFFFFFu, r1 = 0xFFFFFFFFu, r2 = 0xFFFFFFFFu, r3 = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r0 = __viaddmin_u32((unsigned)v[i], nb, r0);
        r1 = __viaddmin_u32((unsigned)v[8 + i], nb, r1);
        r2 = __viaddmin_u32((unsigned)v[16 + i], nb, r2);
        r3 = __viaddmin_u32((unsigned)v[24 + i], nb, r3);
    }
    return umin3i(umin3i(r0, r1, r2), r3, r0);
#endif
}

// Queries given by index (q_gather) are packed once so that TMA can fetch a round's rows as
// contiguous boxes.  16 bytes per thread, four loads in flight.
__global__ void k_pack(const uint8_t *__restrict__ qpool, const int32_t *__restrict__ gather, int64_t nq,
                       uint8_t *__restrict__ qpacked) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int sub = (int)(gt & 7);                // 8 threads (16 B each) per descriptor
    constexpr int U = 4;
    const int64_t quarter = (nq + U - 1) / U;
    uint4 x[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const int64_t r = (gt >> 3) + k * quarter;
        x[k] = make_uint4(0, 0, 0, 0);
        if ((gt >> 3) < quarter && r < nq) x[k] = *(const uint4 *)(qpool + (int64_t)gather[r] * FM_DIM + sub * 16);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const int64_t r = (gt >> 3) + k * quarter;
        if ((gt >> 3) < quarter && r < nq) *(uint4 *)(qpacked + r * FM_DIM + sub * 16) = x[k];
    }
}

// |row|^2 of one staged descriptor (128 B, 128B-swizzled tile: the swizzle only permutes the 16-byte
// chunks inside the row, so the sum is taken in whatever order avoids bank conflicts)
__device__ __forceinline__ int row_norm(uint32_t tile_saddr, int r) {
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int4 x = ld_shared_v4(tile_saddr + r * FM_DIM + ((c ^ (r & 7)) << 4));
        s = __dp4a((unsigned)x.x, (unsigned)x.x, s); s = __dp4a((unsigned)x.y, (unsigned)x.y, s);
        s = __dp4a((unsigned)x.z, (unsigned)x.z, s); s = __dp4a((unsigned)x.w, (unsigned)x.w, s);
    }
    return (int)s;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_grouped_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
             const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Bars *bars = (Bars *)(smem + SMEM_BARS);
    Slot *ring = (Slot *)(smem + SMEM_RING);            // [2 accumulator buffers][RING]
    unsigned long long *skeys = (unsigned long long *)(smem + SMEM_KEYS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FM_TC_PROF
    unsigned long long _gp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long _gk0 = clock64();
#endif

    if (threadIdx.x == 0) {
        // a stage is free again when the MMAs that read it have completed (tcgen05.commit) and both norm warps are done with it
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1 + NORM_WARPS); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tmem_full[i]), 1); mbar_init(smem_u32(&bars->tmem_empty[i]), EPI_WARPS / 2); }
        for (int i = 0; i < 2 * RING; ++i) mbar_init(smem_u32(&bars->norm_full[i]), NORM_WARPS);
        fence_barrier_init();
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_t);
    }
    if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== producer: claims groups, emits units =====================
        // All B-chunks of a slab go to the same accumulator buffer (= the same epilogue group, which
        // carries the rows' running minima across the chunks).  Two slabs are in flight, one per
        // buffer, and their units are emitted alternately, so that the tensor core fills one
        // accumulator while the other group drains its own (the MMAs are issued in unit order: two
        // consecutive units on the same buffer would serialise MMA and epilogue).
        // Groups are claimed two steps ahead (atomic, then offsets) so that neither latency is exposed.
        struct Stream { bool active; int pass, na, nbr, a0, b0; int64_t a_src, b_src, a_out; };
        Stream st[2];
        st[0].active = st[1].active = false;
        bool stopped[2] = {false, false};
        // claim pipeline
        int claimed = lane == 0 ? atomicAdd(P.counter, 1) : 0;                 // stage 1: atomic in flight
        int g1 = __shfl_sync(0xffffffffu, claimed, 0);                         // stage 2: offsets in flight
        int64_t o1_q0 = 0, o1_q1 = 0, o1_t0 = 0, o1_t1 = 0, o1_tb = 0;
        if (g1 < P.G) {
            o1_q0 = P.q_off[g1]; o1_q1 = P.q_off[g1 + 1]; o1_t0 = P.t_off[g1]; o1_t1 = P.t_off[g1 + 1];
            o1_tb = P.t_base ? P.t_base[g1] : o1_t0;
        }
        claimed = lane == 0 ? atomicAdd(P.counter, 1) : 0;
        // current group / slab iterator
        bool have_group = false, exhausted = false;
        int64_t q0 = 0, t0 = 0, tsrc = 0;
        int nq = 0, nt = 0, it_pass = 0, it_a0 = 0;
        int u = 0, nb[2] = {0, 0};

        auto next_slab = [&](Stream &x) -> bool {
            for (;;) {
                if (!have_group) {
                    if (g1 >= P.G) { exhausted = true; return false; }
                    q0 = o1_q0; nq = (int)(o1_q1 - o1_q0); t0 = o1_t0; nt = (int)(o1_t1 - o1_t0); tsrc = o1_tb;
                    g1 = __shfl_sync(0xffffffffu, claimed, 0);
                    if (g1 < P.G) {
                        o1_q0 = P.q_off[g1]; o1_q1 = P.q_off[g1 + 1]; o1_t0 = P.t_off[g1]; o1_t1 = P.t_off[g1 + 1];
                        o1_tb = P.t_base ? P.t_base[g1] : o1_t0;
                    }
                    claimed = lane == 0 ? atomicAdd(P.counter, 1) : 0;
                    if (nq == 0 || nt == 0) continue;          // nothing to multiply (k_mutual fills these rows)
                    have_group = true; it_pass = 0; it_a0 = 0;
                }
                const int na = it_pass == 0 ? nq : nt;
                if (it_a0 >= na) {
                    if (it_pass == 0) { it_pass = 1; it_a0 = 0; } else have_group = false;
                    continue;
                }
                x.active = true; x.pass = it_pass; x.na = na; x.nbr = it_pass == 0 ? nt : nq;
                x.a0 = it_a0; x.b0 = 0;
                x.a_src = it_pass == 0 ? q0 : tsrc; x.b_src = it_pass == 0 ? tsrc : q0;
                x.a_out = it_pass == 0 ? q0 : t0;
                it_a0 += BM;
                return true;
            }
        };

        while (!(stopped[0] && stopped[1])) {
#pragma unroll
            for (int buf = 0; buf < 2; ++buf) {
                if (stopped[buf]) continue;
                Stream &x = st[buf];
                bool stop = false;
                if (!x.active && (exhausted || !next_slab(x))) stop = true;
                const int stage = u % STAGES;
                const int slot_idx = buf * RING + (nb[buf]++ % RING);
                Slot *sl = &ring[slot_idx];
                GP_T(_p0);
                mbar_wait(smem_u32(&bars->empty[stage]), ((u / STAGES) & 1) ^ 1);
                GP_T(_p1);
                GP_ACC(0, _p0, _p1);
                const uint32_t fb = smem_u32(&bars->full[stage]);
                ++u;
                if (stop) {                 // after the last slab: one stop unit per accumulator buffer
                    if (lane == 0) {
                        sl->flags = F_STOP;
                        bars->stage_slot[stage] = slot_idx;
                        mbar_expect_tx(fb, 0);
                    }
                    __syncwarp();
                    stopped[buf] = true;
                    continue;
                }
                const int a_valid = min(BM, x.na - x.a0), b_valid = min(BN, x.nbr - x.b0);
                const int abox = (a_valid + BOX_ROWS - 1) / BOX_ROWS, bbox = (b_valid + BOX_ROWS - 1) / BOX_ROWS;
                const int arow = (int)(x.a_src + x.a0), brow = (int)(x.b_src + x.b0);
                // unit header (the norm warps fill in ck / an once the tiles have landed)
                if (lane == 0) {
                    sl->a_row0 = arow; sl->a_valid = a_valid;
                    sl->b_row0 = brow; sl->b_valid = b_valid;
                    sl->a_out0 = (int)(x.a_out + x.a0); sl->b_local0 = x.b0;
                    sl->flags = (x.pass ? F_PASS1 : 0) | (x.b0 == 0 ? F_FIRST : 0) | (x.b0 + BN >= x.nbr ? F_LAST : 0);
                    sl->n_mma = (b_valid + 15) & ~15;
                    bars->stage_slot[stage] = slot_idx;
                }
                __syncwarp();
                // 32-row boxes (only the rows the unit needs are read): up to 4 for the A slab,
                // up to 8 for the B rows; everything below is warp-uniform, one elected lane issues
                const uint32_t sa = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + SMEM_STAGES + stage * STAGE_BYTES;
                const CUtensorMap *amap = x.pass == 0 ? &map_q : &map_t, *bmap = x.pass == 0 ? &map_t : &map_q;
                if (elect_one()) {
                    mbar_expect_tx(fb, (abox + bbox) * BOX_BYTES);
                    for (int i = 0; i < abox; ++i)
                        tma_load_2d(sa + i * BOX_BYTES, amap, 0, arow + i * BOX_ROWS, fb);
                    for (int i = 0; i < bbox; ++i)
                        tma_load_2d(sa + A_BYTES + i * BOX_BYTES, bmap, 0, brow + i * BOX_ROWS, fb);
                }
                __syncwarp();
                x.b0 += BN;
                if (x.b0 >= x.nbr) x.active = false;
                GP_T(_p2);
                GP_ACC(1, _p1, _p2);
#ifdef FM_TC_PROF
                _gp[2] += 1;
#endif
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the loop (operands stay in uniform registers), one elected lane issues.
        {
            const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
            int nb[2] = {0, 0}, stops = 0;
            for (int u = 0; stops < 2; ++u) {
                const int stage = u % STAGES;
                GP_T(_m0);
                mbar_wait(smem_u32(&bars->full[stage]), (u / STAGES) & 1);
                GP_T(_m1);
                GP_ACC(3, _m0, _m1);
                tc_fence_after();
                const int slot_idx = __shfl_sync(0xffffffffu, bars->stage_slot[stage], 0);
                const Slot *sl = &ring[slot_idx];
                const int buf = slot_idx / RING;
                const int kb = nb[buf]++;
                // the accumulator buffer must have been drained by the epilogue before its "full"
                // barrier is signalled again -- for the stop unit too, or the barrier could run two
                // phases ahead of a slow epilogue warp (parity aliasing)
                mbar_wait(smem_u32(&bars->tmem_empty[buf]), (kb & 1) ^ 1);
                GP_T(_m2);
                GP_ACC(4, _m1, _m2);
                tc_fence_after();
                const int flags = __shfl_sync(0xffffffffu, sl->flags, 0);
                const int n_mma = __shfl_sync(0xffffffffu, sl->n_mma, 0);
                if (flags & F_STOP) {
                    // tcgen05.commit arrives only after every MMA issued above has completed, so the
                    // stop signal cannot overtake the accumulators still in flight
                    if (elect_one()) {
                        umma_commit(smem_u32(&bars->tmem_full[buf]));
                        umma_commit(smem_u32(&bars->empty[stage]));
                    }
                    __syncwarp();
                    ++stops;
                    continue;
                }
                const uint32_t idesc = make_idesc(BM, n_mma);
                const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES;
                const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < FM_DIM / 32; ++k)
                        umma_i8(tbase + buf * BN, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
                    umma_commit(smem_u32(&bars->tmem_full[buf]));
                    umma_commit(smem_u32(&bars->empty[stage]));
                }
                __syncwarp();
                GP_T(_m3);
                GP_ACC(5, _m2, _m3);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== norm warps =====================
        // |b_c|^2 of every B row (-> exact-key constants) and, for the last unit of a slab, |a_i|^2 of
        // the A rows, summed from the staged tiles while the tensor core multiplies them.
        const int t64 = (warp - 2) * 32 + lane;
        const uint32_t sbase = smem_u32(smem);
        int stops = 0;
        for (int u = 0; stops < 2; ++u) {
            const int stage = u % STAGES;
            mbar_wait(smem_u32(&bars->full[stage]), (u / STAGES) & 1);
            const int slot_idx = bars->stage_slot[stage];
            Slot *sl = &ring[slot_idx];
            const int flags = sl->flags;
            if (!(flags & F_STOP)) {
                const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES;
                const int b_valid = sl->b_valid;
                for (int r = t64; r < b_valid; r += NORM_WARPS * 32)
                    sl->ck[r] = (int)(((unsigned)row_norm(sa + A_BYTES, r) << 8) | (unsigned)r);
                if (flags & F_LAST) {
                    const int a_valid = sl->a_valid;
                    for (int r = t64; r < a_valid; r += NORM_WARPS * 32) sl->an[r] = row_norm(sa, r);
                }
            } else {
                ++stops;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&bars->norm_full[slot_idx]));
                mbar_arrive(smem_u32(&bars->empty[stage]));
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // Two groups of 8 warps, one per accumulator buffer; a warp owns 32 rows (its TMEM lane
        // quarter) x 128 columns (a column half), swept in 32-column pieces.
        const int ew = warp - 4, grp = ew >> 3, lq = warp & 3, ch = (ew >> 2) & 1;
        const int row = lq * 32 + lane;
        const uint32_t taddr_grp = __shfl_sync(0xffffffffu, tmem_base + ((uint32_t)(lq * 32) << 16) + grp * BN, 0);
        const uint32_t full_a = smem_u32(&bars->tmem_full[grp]), empty_a = smem_u32(&bars->tmem_empty[grp]);
        int m1 = NONE_P, i1 = -1, m2 = NONE_P, i2 = -1, nslab = 0;
        for (int k = 0;; ++k) {
            GP_T(_e0);
            mbar_wait(full_a, k & 1);
            GP_T(_e1);
            GP_ACC(6, _e0, _e1);
            tc_fence_after();
            const Slot *sl = &ring[grp * RING + (k % RING)];
            mbar_wait(smem_u32(&bars->norm_full[grp * RING + (k % RING)]), (k / RING) & 1);
            const int flags = sl->flags;
            if (flags & F_STOP) break;
            const int b_valid = sl->b_valid, b_local0 = sl->b_local0;
            // the unit's valid columns are split evenly (in 32-column pieces) between the two warps
            // that sweep a row, so that short units do not leave the second one idle
            const int half = ((b_valid + 63) >> 6) << 5;                                   // <= 128
            // (a warp whose 32 rows all lie beyond the slab's last valid row has nothing to sweep)
            const int ncols = lq * 32 >= sl->a_valid ? 0 : (ch ? max(b_valid - half, 0) : min(half, b_valid));   // warp-uniform
            const uint32_t taddr0 = taddr_grp + ch * half;
            const bool top1 = (flags & F_PASS1) != 0;                                      // warp-uniform
            // |a_i|^2 is only needed when the slab is written out
            int an = 0;
            if ((flags & F_LAST) && row < sl->a_valid) an = sl->an[row];
            if (flags & F_FIRST) { m1 = m2 = NONE_P; i1 = i2 = -1; }
            const uint32_t cka = smem_u32(&sl->ck[0]) + ch * half * 4;
            int k1 = I32_MAX, k2 = I32_MAX;
            if (ncols == 0) {        // nothing to read: hand the accumulator back right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_a);
            }
            // Per 32 columns: exact keys in place of the accumulators (columns past the last B row ->
            // INT_MAX), smallest key by a 3-input min tree and, for the top-2 pass, the runner-up =
            // smallest key above it (keys are distinct or INT_MAX), obtained as the unsigned minimum
            // of key - (kmin + 1).  The accumulator goes back to the MMA warp after the last read.
#pragma unroll 1
            for (int base = 0; base < ncols; base += 32) {
                int V[32];
                GP_T(_q0);
                tmem_ld32(taddr0 + base, V);
                tmem_ld_wait();
                GP_T(_q1);
                GP_ACC(13, _q0, _q1);
                if (base + 32 >= ncols) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_a);
                }
                if (base + 32 <= ncols) {       // full piece (the common case): no per-column checks
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const int4 ck = ld_shared_v4(cka + (base + c4 * 4) * 4);
                        V[c4 * 4 + 0] = ck.x - 512 * V[c4 * 4 + 0];
                        V[c4 * 4 + 1] = ck.y - 512 * V[c4 * 4 + 1];
                        V[c4 * 4 + 2] = ck.z - 512 * V[c4 * 4 + 2];
                        V[c4 * 4 + 3] = ck.w - 512 * V[c4 * 4 + 3];
                    }
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const int cb = base + c4 * 4;
                        if (cb + 4 <= ncols) {
                            const int4 ck = ld_shared_v4(cka + cb * 4);
                            V[c4 * 4 + 0] = ck.x - 512 * V[c4 * 4 + 0];
                            V[c4 * 4 + 1] = ck.y - 512 * V[c4 * 4 + 1];
                            V[c4 * 4 + 2] = ck.z - 512 * V[c4 * 4 + 2];
                            V[c4 * 4 + 3] = ck.w - 512 * V[c4 * 4 + 3];
                        } else if (cb < ncols) {
                            const int4 ck = ld_shared_v4(cka + cb * 4);
                            V[c4 * 4 + 0] = ck.x - 512 * V[c4 * 4 + 0];
                            V[c4 * 4 + 1] = cb + 1 < ncols ? ck.y - 512 * V[c4 * 4 + 1] : I32_MAX;
                            V[c4 * 4 + 2] = cb + 2 < ncols ? ck.z - 512 * V[c4 * 4 + 2] : I32_MAX;
                            V[c4 * 4 + 3] = I32_MAX;
                        } else {
                            V[c4 * 4 + 0] = V[c4 * 4 + 1] = V[c4 * 4 + 2] = V[c4 * 4 + 3] = I32_MAX;
                        }
                    }
                }
                GP_T(_q2);
                GP_ACC(14, _q1, _q2);
#ifdef FM_TC_PROF
                _gp[15] += 1;
#endif
                const int h1 = min32(V);
                int h2 = I32_MAX;
                if (!top1) {
                    const int hb = h1 + 1;
                    const unsigned um = umin32(V, hb);
                    h2 = um >= (unsigned)(I32_MAX - hb) ? I32_MAX : (int)(um + (unsigned)hb);
                }
                k2 = min3i(max(k1, h1), k2, h2);
                k1 = min(k1, h1);
            }
            GP_T(_e2);
            GP_ACC(7, _e1, _e2);
            // merge the unit's minima into the row state (chunks arrive in increasing index)
            if (ncols > 0) {
                const int p1 = k1 >> 8;
                if (k1 != I32_MAX && p1 < m2) {
                    const int j1 = b_local0 + (k1 & 255);
                    if (p1 < m1) {
                        const int p2 = k2 >> 8;
                        if (k2 != I32_MAX && p2 < m1) { m2 = p2; i2 = b_local0 + (k2 & 255); }
                        else { m2 = m1; i2 = i1; }
                        m1 = p1; i1 = j1;
                    } else {
                        m2 = p1; i2 = j1;
                    }
                }
            }
            GP_T(_e3);
            GP_ACC(8, _e2, _e3);
            if (flags & F_LAST) {
                // combine the two column halves of each row (the 2 warps of this group and lane
                // quarter: own named barrier, double-buffered exchange area), add |a_i|^2, write out
                const int a_valid = sl->a_valid, a_out0 = sl->a_out0;
                unsigned long long *sk = skeys + ((grp * 2 + (nslab & 1)) * BM + row) * 4;
                ++nslab;
                sk[ch * 2] = i1 < 0 ? FM_NONE_KEY : pack_key((uint32_t)(m1 + an), (uint32_t)i1);
                sk[ch * 2 + 1] = i2 < 0 ? FM_NONE_KEY : pack_key((uint32_t)(m2 + an), (uint32_t)i2);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + lq) : "memory");
                if (ch == 0 && row < a_valid) {
                    unsigned long long a = sk[0], b = sk[1];
                    merge2(a, b, sk[2], sk[3]);
                    const int64_t o = (int64_t)a_out0 + row;
                    if (top1) {
                        P.t2q_idx[o] = a == FM_NONE_KEY ? -1 : (int32_t)(uint32_t)a;
                    } else {
                        *(uint2 *)(P.q2t_d2 + o * 2) = make_uint2((uint32_t)(a >> 32), (uint32_t)(b >> 32));
                        *(int2 *)(P.q2t_idx + o * 2) = make_int2(a == FM_NONE_KEY ? -1 : (int32_t)(uint32_t)a,
                                                                 b == FM_NONE_KEY ? -1 : (int32_t)(uint32_t)b);
                    }
                }
            }
            GP_T(_e4);
            GP_ACC(9, _e3, _e4);
#ifdef FM_TC_PROF
            _gp[10] += 1;
#endif
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef FM_TC_PROF
    if (lane == 0 && (warp == 0 || warp == 1 || warp == 4))
        for (int i = 0; i < 16; ++i) if (i != 11 && i != 12 && _gp[i]) atomicAdd(&g_gprof[i], _gp[i]);
    if (threadIdx.x == 0) { atomicAdd(&g_gprof[11], (unsigned long long)(clock64() - _gk0)); atomicAdd(&g_gprof[12], 1ull); }
#endif
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// crossCheck predicate per local query (needs both passes of the whole grid)
// Per group: the rows of a group with an empty side (no unit ever touches them) get "no neighbour";
// the others get their crossCheck flag.
__global__ void k_mutual(const int64_t *__restrict__ q_off, const int64_t *__restrict__ t_off,
                         uint32_t *__restrict__ q2t_d2, int32_t *__restrict__ q2t_idx,
                         int32_t *__restrict__ t2q_idx, uint8_t *__restrict__ mutual) {
    const int g = blockIdx.x;
    const int64_t q0 = q_off[g], nq = q_off[g + 1] - q0, t0 = t_off[g], nt = t_off[g + 1] - t0;
    if (nq == 0 || nt == 0) {
        for (int64_t i = threadIdx.x; i < nq; i += blockDim.x) {
            *(uint2 *)(q2t_d2 + (q0 + i) * 2) = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
            *(int2 *)(q2t_idx + (q0 + i) * 2) = make_int2(-1, -1);
            if (mutual) mutual[q0 + i] = 0;
        }
        for (int64_t j = threadIdx.x; j < nt; j += blockDim.x) t2q_idx[t0 + j] = -1;
        return;
    }
    if (mutual == nullptr) return;
    for (int64_t i = threadIdx.x; i < nq; i += blockDim.x) {
        const int32_t j = q2t_idx[(q0 + i) * 2];
        mutual[q0 + i] = (j >= 0 && t2q_idx[t0 + j] == (int32_t)i) ? 1 : 0;
    }
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace gtc

#ifdef FM_TC_PROF
extern "C" int fm_debug_gprof(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, gtc::g_gprof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(gtc::g_gprof, z, sizeof(z)); }
    return 0;
}
#endif

size_t grouped_tc_workspace_bytes(int64_t total_q, int64_t tpool_rows, bool gather) {
    using namespace gtc;
    (void)tpool_rows;
    return up256(256) + (gather ? up256((size_t)total_q * FM_DIM) : 0) + 256;
}

int launch_grouped_tc(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                      const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base, int32_t G,
                      int64_t total_q, int64_t total_t, int64_t tpool_rows, uint32_t *q2t_d2,
                      int32_t *q2t_idx, int32_t *t2q_idx, uint8_t *mutual, void *ws, size_t ws_bytes,
                      cudaStream_t s) {
    using namespace gtc;
    if (ws_bytes < grouped_tc_workspace_bytes(total_q, tpool_rows, q_gather != nullptr)) {
        set_error("grouped tcgen05 path: workspace too small");
        return FM_ENOSPACE;
    }
    if (!(total_q > 0 && total_t > 0 && tpool_rows > 0)) {     // nothing to match: every slot is "no neighbour"
        if (total_q > 0) {
            FM_CUDA_TRY(cudaMemsetAsync(q2t_d2, 0xFF, (size_t)total_q * 8, s));
            FM_CUDA_TRY(cudaMemsetAsync(q2t_idx, 0xFF, (size_t)total_q * 8, s));
            if (mutual) FM_CUDA_TRY(cudaMemsetAsync(mutual, 0, (size_t)total_q, s));
        }
        if (total_t > 0) FM_CUDA_TRY(cudaMemsetAsync(t2q_idx, 0xFF, (size_t)total_t * 4, s));
        return FM_OK;
    }
    {
        uint8_t *w = (uint8_t *)ws;
        int *counter = (int *)w; w += up256(256);
        uint8_t *qpack = q_gather ? w : nullptr;
        FM_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), s));          // the kernel's group counter
        if (q_gather) {
            k_pack<<<(unsigned)((((total_q + 3) / 4) * 8 + 255) / 256), 256, 0, s>>>(qpool, q_gather, total_q, qpack);
            FM_CUDA_TRY(cudaGetLastError());
            count_launch();
        }
        CUtensorMap map_q, map_t;
        int rc;
        if ((rc = make_map(&map_q, qpack ? qpack : qpool, total_q, FM_DIM, BOX_ROWS, CU_TENSOR_MAP_SWIZZLE_128B)) != FM_OK) return rc;
        if ((rc = make_map(&map_t, tpool, tpool_rows, FM_DIM, BOX_ROWS, CU_TENSOR_MAP_SWIZZLE_128B)) != FM_OK) return rc;
        // function attributes and SM counts are per device: a process may drive several GPUs
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        static std::mutex attr_mu;
        static bool attr_set[64] = {};
        static int sm_count[64] = {};
        {
            std::lock_guard<std::mutex> lock(attr_mu);
            const int slot = dev >= 0 && dev < 64 ? dev : 0;
            if (!attr_set[slot] || slot != dev) {
                FM_CUDA_TRY(cudaFuncSetAttribute(k_grouped_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
                cudaDeviceGetAttribute(&sm_count[slot], cudaDevAttrMultiProcessorCount, dev);
                attr_set[slot] = true;
            }
            if (sm_count[slot] > 0) sms = sm_count[slot];
        }
        Params P{q_off, t_off, t_base, G, counter, q2t_d2, q2t_idx, t2q_idx};
        const int grid = G < sms ? G : sms;
        prof_begin(s);
        k_grouped_tc<<<grid, NTHREADS, SMEM_ALLOC, s>>>(map_q, map_t, P);
        prof_end(s);
        FM_CUDA_TRY(cudaGetLastError());
        count_launch();
    }
    // rows of groups with an empty side + (optionally) the crossCheck flags
    k_mutual<<<G, 128, 0, s>>>(q_off, t_off, q2t_d2, q2t_idx, t2q_idx, mutual);
    FM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return FM_OK;
}

}  // namespace fm
