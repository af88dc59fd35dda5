// Micro-benchmark: issue rate of tcgen05.mma kind::i8 (u8 x u8 -> s32) for the shapes the dense
// kernel could use: single CTA or CTA pair (cta_group::2), N = 128 / 256, K-steps issued
// back-to-back on one accumulator or interleaved over several.  Operands are whatever is in
// shared memory (zeros); only the timing matters.  One issuing thread per CTA, every SM busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fast_match_b200/csrc \
//        -I../../include umma_rate.cu -o umma_rate -lcuda
#include <cstdio>
#include <cuda_runtime.h>

#include "fm_tc_ptx.cuh"

using namespace fm::tc;

constexpr int KSTEPS = 5;       // the dense kernel issues 4 descriptor K-steps + the norm step
constexpr int ITERS = 400;

struct Bar { unsigned long long done; uint32_t tmem; uint32_t pad; unsigned long long scratch[4]; };

// MODE 0: for acc { for k: mma(acc, k) }      (the dense kernel's order)
// MODE 1: for k   { for acc: mma(acc, k) }    (K-steps interleaved over the accumulators)
// LDW: that many extra warps keep reading TMEM (tcgen05.ld 32x32b.x32, like the epilogue) while the
// MMAs run; X5: the 5th K-step uses 32-byte-swizzle descriptors (the norm K-step of the dense kernel).
// CMT: commit to a scratch barrier after every accumulator's K-steps (as the dense kernel does);
// ALUW: that many extra warps run a dependent integer-max loop (issue-slot competition).
template <bool PAIR, int N, int NACC, int MODE, int LDW = 0, bool X5 = false, bool CMT = false, int ALUW = 0, int RD_BASE = 0, int RD_SPAN = 512, int MMA_ON = 1, int CMT_EVERY = 1, int LDX = 32>
__device__ __forceinline__ void body(long long *out, int slot) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Bar *bar = (Bar *)(smem + 96 * 1024);
    volatile int *stop = (volatile int *)(smem + 96 * 1024 + 128);
    if (threadIdx.x == 0) *stop = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar->done), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar->scratch[i]), 1); fence_barrier_init(); }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(smem_u32(&bar->tmem), 512); else tmem_alloc(smem_u32(&bar->tmem), 512); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem;
    long long t0 = 0, t1 = 0;
    if (warp == 0 && lane == 0) {
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc(PAIR ? 256 : 128, N);
            const uint64_t adesc = make_desc(smem_u32(smem));
            const uint64_t bdesc = make_desc(smem_u32(smem + 32 * 1024));
            const uint64_t axdesc = make_desc_sw32(smem_u32(smem + 80 * 1024));
            const uint64_t bxdesc = make_desc_sw32(smem_u32(smem + 84 * 1024));
            auto mma = [&](int a, int k) {
                const uint64_t ad = X5 && k == 4 ? axdesc : adesc + 2 * (k & 3);
                const uint64_t bd = X5 && k == 4 ? bxdesc + a * 128 : bdesc + 2 * (k & 3) + a * 512;
                if (PAIR) umma_i8_pair(tmem + a * N, ad, bd, idesc, k > 0);
                else umma_i8(tmem + a * N, ad, bd, idesc, k > 0);
            };
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < (MMA_ON ? ITERS : 0); ++it) {
                if (MODE == 0) {
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k) mma(a, k);
                        if (CMT && (a % CMT_EVERY) == CMT_EVERY - 1) { if (PAIR) umma_commit_pair(smem_u32(&bar->scratch[a])); else umma_commit(smem_u32(&bar->scratch[a])); }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
                        for (int a = 0; a < NACC; ++a) mma(a, k);
                }
            }
            if (MMA_ON) { if (PAIR) umma_commit_pair(smem_u32(&bar->done)); else umma_commit(smem_u32(&bar->done)); }
        }
        if (MMA_ON) mbar_wait(smem_u32(&bar->done), 0);
        else { while (clock64() - t0 < 200000) {} }
        t1 = clock64();
        if (blockIdx.x == 0) out[slot] = t1 - t0;
        *stop = 1;
    } else if (warp >= 4 && warp < 4 + LDW) {
        // epilogue-like readers: every warp sweeps 128 columns of its lane quarter, over and over
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + RD_BASE + ((warp - 4) >> 2) * (RD_SPAN / 4);
        int acc = 0;
        long long nld = 0;
        while (!*stop) {
            nld += RD_SPAN / 4 / 32;   /* counted in 32-column units */
#pragma unroll
            for (int c = 0; c < RD_SPAN / 4; c += LDX) {
                if (LDX == 32) {
                    int v[32];
                    tmem_ld32(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc ^= v[i];
                } else {
                    int v[64];
                    tmem_ld64(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 64; ++i) acc ^= v[i];
                }
            }
        }
        if (acc == 0x12345) out[63] = acc;
        if (lane == 0 && blockIdx.x == 0) atomicAdd((unsigned long long *)&out[32 + slot], (unsigned long long)nld);
    } else if (warp >= 4 && warp < 4 + ALUW) {
        int x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * (i + 1);
        while (!*stop) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __vimax3_s32(x[i], x[(i + 1) & 7] ^ r, i);
        }
        int acc = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= x[i];
        if (acc == 0x12345) out[63] = acc;
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, int MODE, int LDW = 0, bool X5 = false, bool CMT = false, int ALUW = 0>
__global__ void __launch_bounds__(640, 1) k_single(long long *out, int slot) { body<false, N, NACC, MODE, LDW, X5, CMT, ALUW>(out, slot); }
template <bool PAIR, int N, int NACC, int RD_BASE, int RD_SPAN, int MMA_ON>
__global__ void __launch_bounds__(640, 1) k_rd(long long *out, int slot) { body<false, N, NACC, 0, 16, false, true, 0, RD_BASE, RD_SPAN, MMA_ON>(out, slot); }
template <int CMT_EVERY, int LDX, bool CMT>
__global__ void __launch_bounds__(640, 1) k_cs(long long *out, int slot) { body<false, 128, 4, 0, 16, true, CMT, 0, 0, 512, 1, CMT_EVERY, LDX>(out, slot); }
template <int N, int NACC, int MODE, int LDW = 0, bool X5 = false, bool CMT = false, int ALUW = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) k_pair(long long *out, int slot) { body<true, N, NACC, MODE, LDW, X5, CMT, ALUW>(out, slot); }

template <typename K>
static void run(const char *name, K kern, int n, int nacc, bool pair, long long *d_out, int slot) {
    const int smem = 100 * 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) kern<<<148, 640, smem>>>(d_out, slot);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out + slot, 8, cudaMemcpyDeviceToHost);
    const double per_instr = (double)cyc / (ITERS * KSTEPS * nacc);
    // MACs per clk per SM: a pair instruction covers 256 rows over two SMs
    const double macs = 128.0 * n * 32 / per_instr;
    unsigned long long nld = 0;
    cudaMemcpy(&nld, d_out + 32 + slot, 8, cudaMemcpyDeviceToHost);
    cudaMemset(d_out + 32 + slot, 0, 8);
    // both launches add their loads; each 32-column load of a warp moves 4 KB
    printf("%-44s %s  %7.1f clk/instr  %7.0f MAC/clk/SM   TMEM read %6.1f B/clk/SM\n", name, e == cudaSuccess ? "ok " : cudaGetErrorString(e), per_instr, macs,
           (double)nld * 4096.0 / 2 / (double)cyc);
}

int main(int argc, char **argv) {
    long long *d_out;
    cudaMalloc(&d_out, 64 * 8);
    cudaMemset(d_out, 0, 64 * 8);
    int s = 0;
    if (argc > 3) {
        printf("-- N=128 x4 acc, 16 reader warps: commit spacing and reader load width\n");
        run("no commits, x32 readers", k_cs<1, 32, false>, 128, 4, false, d_out, s++);
        run("commit every acc, x32 readers", k_cs<1, 32, true>, 128, 4, false, d_out, s++);
        run("commit every 2 acc, x32 readers", k_cs<2, 32, true>, 128, 4, false, d_out, s++);
        run("commit every 4 acc, x32 readers", k_cs<4, 32, true>, 128, 4, false, d_out, s++);
        run("commit every acc, x64 readers", k_cs<1, 64, true>, 128, 4, false, d_out, s++);
        run("no commits, x64 readers", k_cs<1, 64, false>, 128, 4, false, d_out, s++);
        return 0;
    }
    if (argc > 2) {
        printf("-- TMEM read rate of 16 reader warps (readers on [base, base+span)) vs concurrent MMAs\n");
        run("no MMA, readers on [256,512)", k_rd<false, 256, 1, 256, 256, 0>, 256, 1, false, d_out, s++);
        run("no MMA, readers on [0,512)", k_rd<false, 256, 1, 0, 512, 0>, 256, 1, false, d_out, s++);
        run("N=256 acc [0,256), readers [256,512)", k_rd<false, 256, 1, 256, 256, 1>, 256, 1, false, d_out, s++);
        run("N=128 acc [0,128), readers [256,512)", k_rd<false, 128, 1, 256, 256, 1>, 128, 1, false, d_out, s++);
        run("N=128 acc [0,128), readers [128,256)", k_rd<false, 128, 1, 128, 128, 1>, 128, 1, false, d_out, s++);
        run("N=128 acc [0,256) x2, readers [256,512)", k_rd<false, 128, 2, 256, 256, 1>, 128, 2, false, d_out, s++);
        run("N=128 acc x4 all, readers all", k_rd<false, 128, 4, 0, 512, 1>, 128, 4, false, d_out, s++);
        run("N=256 acc x2 all, readers all", k_rd<false, 256, 2, 0, 512, 1>, 256, 2, false, d_out, s++);
        return 0;
    }
    if (argc > 1) {
        printf("-- commit after every accumulator (5 K-steps)\n");
        run("single N=256 2 acc + commits", k_single<256, 2, 0, 0, true, true>, 256, 2, false, d_out, s++);
        run("single N=128 4 acc + commits", k_single<128, 4, 0, 0, true, true>, 128, 4, false, d_out, s++);
        run("pair   N=256 2 acc + commits", k_pair<256, 2, 0, 0, true, true>, 256, 2, true, d_out, s++);
        run("pair   N=128 4 acc + commits", k_pair<128, 4, 0, 0, true, true>, 128, 4, true, d_out, s++);
        printf("-- 16 integer-busy warps beside the issuing thread\n");
        run("single N=256 2 acc + alu", k_single<256, 2, 0, 0, true, false, 16>, 256, 2, false, d_out, s++);
        run("pair   N=128 4 acc + alu", k_pair<128, 4, 0, 0, true, false, 16>, 128, 4, true, d_out, s++);
        run("single N=256 2 acc + commits + alu", k_single<256, 2, 0, 0, true, true, 16>, 256, 2, false, d_out, s++);
        run("pair   N=128 4 acc + commits + alu", k_pair<128, 4, 0, 0, true, true, 16>, 128, 4, true, d_out, s++);
        return 0;
    }
    run("single N=256 1 acc", k_single<256, 1, 0>, 256, 1, false, d_out, s++);
    run("single N=256 2 acc, acc-major", k_single<256, 2, 0>, 256, 2, false, d_out, s++);
    run("single N=256 2 acc, k-major", k_single<256, 2, 1>, 256, 2, false, d_out, s++);
    run("single N=128 1 acc", k_single<128, 1, 0>, 128, 1, false, d_out, s++);
    run("single N=128 4 acc, acc-major", k_single<128, 4, 0>, 128, 4, false, d_out, s++);
    run("single N=128 4 acc, k-major", k_single<128, 4, 1>, 128, 4, false, d_out, s++);
    run("single N=64 4 acc, k-major", k_single<64, 4, 1>, 64, 4, false, d_out, s++);
    run("pair   N=256 1 acc", k_pair<256, 1, 0>, 256, 1, true, d_out, s++);
    run("pair   N=256 2 acc, acc-major", k_pair<256, 2, 0>, 256, 2, true, d_out, s++);
    run("pair   N=256 2 acc, k-major", k_pair<256, 2, 1>, 256, 2, true, d_out, s++);
    run("pair   N=128 1 acc", k_pair<128, 1, 0>, 128, 1, true, d_out, s++);
    run("pair   N=128 4 acc, acc-major", k_pair<128, 4, 0>, 128, 4, true, d_out, s++);
    run("pair   N=128 4 acc, k-major", k_pair<128, 4, 1>, 128, 4, true, d_out, s++);
    run("pair   N=64 4 acc, k-major", k_pair<64, 4, 1>, 64, 4, true, d_out, s++);
    printf("-- 5th K-step through 32B-swizzle descriptors\n");
    run("single N=256 2 acc, x5", k_single<256, 2, 0, 0, true>, 256, 2, false, d_out, s++);
    run("single N=128 4 acc, x5", k_single<128, 4, 0, 0, true>, 128, 4, false, d_out, s++);
    run("pair   N=128 4 acc, x5", k_pair<128, 4, 0, 0, true>, 128, 4, true, d_out, s++);
    printf("-- with 16 warps reading TMEM all the time\n");
    run("single N=256 2 acc + 16 ld warps", k_single<256, 2, 0, 16>, 256, 2, false, d_out, s++);
    run("single N=128 4 acc + 16 ld warps", k_single<128, 4, 0, 16>, 128, 4, false, d_out, s++);
    run("pair   N=256 2 acc + 16 ld warps", k_pair<256, 2, 0, 16>, 256, 2, true, d_out, s++);
    run("pair   N=128 4 acc + 16 ld warps", k_pair<128, 4, 0, 16>, 128, 4, true, d_out, s++);
    run("single N=256 2 acc + 8 ld warps", k_single<256, 2, 0, 8>, 256, 2, false, d_out, s++);
    run("single N=128 4 acc + 8 ld warps", k_single<128, 4, 0, 8>, 128, 4, false, d_out, s++);
    run("single N=256 2 acc + 16 ld warps, x5", k_single<256, 2, 0, 16, true>, 256, 2, false, d_out, s++);
    run("pair   N=128 4 acc + 16 ld warps, x5", k_pair<128, 4, 0, 16, true>, 128, 4, true, d_out, s++);
    return 0;
}
