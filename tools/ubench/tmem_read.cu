// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM sub-partition, and whether it
// overlaps with integer work issued by other warps of the same sub-partition (sm_100a).
//   LW warps per sub-partition loop over "tcgen05.ld 32x32b.xW; tcgen05.wait::ld" (no arithmetic);
//   AW warps per sub-partition loop over dependent-free VIMNMX3 (the epilogue's filter op).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fast_match_b200/csrc \
//        -I../../include tmem_read.cu -o tmem_read -lcuda
#include <cstdio>
#include <cuda_runtime.h>

#include "fm_tc_ptx.cuh"

using namespace fm::tc;

constexpr int LOOPS = 2000;

template <int LW, int AW, int W>
__global__ void __launch_bounds__(32 * 4 * (LW + AW) + 0, 1) k(long long *out, int slot) {
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done_cnt;
    if (threadIdx.x == 0) done_cnt = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int q = warp & 3, idx = warp >> 2;          // sub-partition, index inside it
    long long t0 = clock64(), t1 = t0;
    int sink = 0;
    if (idx < LW) {
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (idx * 128) % 512;
#pragma unroll 1
        for (int it = 0; it < LOOPS; ++it) {
            if (W == 32) { int v[32]; tmem_ld32(taddr + (it & 1) * 32, v); tmem_ld_wait(); sink ^= v[0] ^ v[31]; }
            else { int v[64]; tmem_ld64(taddr + (it & 1) * 64, v); tmem_ld_wait(); sink ^= v[0] ^ v[63]; }
        }
        t1 = clock64();
        if (lane == 0) atomicAdd((int *)&done_cnt, 1);
        if (blockIdx.x == 0 && lane == 0) atomicMax((unsigned long long *)&out[slot], (unsigned long long)(t1 - t0));
    } else {
        int x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * (i + 3);
        // with loaders present: run for as long as they do; alone: a fixed count
        long long n = 0;
#pragma unroll 1
        for (int it = 0; LW ? done_cnt < 4 * LW : it < LOOPS; ++it, ++n) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __vimax3_s32(x[i], x[(i + 3) & 7] ^ r, it);
        }
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long *)&out[48], (unsigned long long)n);
#pragma unroll
        for (int i = 0; i < 8; ++i) sink ^= x[i];
        if (blockIdx.x == 0 && lane == 0) atomicMax((unsigned long long *)&out[32 + slot], (unsigned long long)(t1 - t0));
    }
    if (sink == 0x1234567) out[63] = sink;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int LW, int AW, int W>
static void run(long long *d_out, int slot) {
    k<LW, AW, W><<<148, 32 * 4 * (LW + AW)>>>(d_out, slot);
    k<LW, AW, W><<<148, 32 * 4 * (LW + AW)>>>(d_out, slot);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[64];
    cudaMemcpy(c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
    cudaMemset(d_out, 0, sizeof(c));
    printf("%d ld warps (x%d) + %d alu warps per sub-partition: %s", LW, W, AW, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (LW) printf("  TMEM read %6.1f B/clk/sub-partition (%5.1f clk per load)", (double)LW * LOOPS * W * 128.0 / c[slot], (double)c[slot] / LOOPS);
    // c[48] = iterations of all alu warps of block 0 over both launches
    if (AW) printf("  VIMNMX3 %5.2f warp-instr/clk/sub-partition", (double)c[48] / 2 / 4 * 32.0 / c[32 + slot]);
    printf("\n");
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 64 * 8);
    cudaMemset(d_out, 0, 64 * 8);
    run<1, 0, 32>(d_out, 0);
    run<2, 0, 32>(d_out, 1);
    run<4, 0, 32>(d_out, 2);
    run<1, 0, 64>(d_out, 3);
    run<2, 0, 64>(d_out, 4);
    run<4, 0, 64>(d_out, 5);
    run<0, 1, 32>(d_out, 6);
    run<0, 2, 32>(d_out, 7);
    run<0, 4, 32>(d_out, 8);
    run<1, 1, 64>(d_out, 9);
    run<2, 2, 64>(d_out, 10);
    run<1, 3, 64>(d_out, 11);
    run<2, 2, 32>(d_out, 12);
    return 0;
}
