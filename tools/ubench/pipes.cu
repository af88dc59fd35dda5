// Micro-benchmark: per-SM throughput of the integer ops the epilogues are built from (sm_100a).
// Each kernel runs NW warps per block on every SM with 8 independent dependency chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void k(int *out, int a0, int b0, int c0) {
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a0 + threadIdx.x * (i + 1);
    unsigned u = (unsigned)b0;
    int y = c0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = __vimax3_s32(x[i], y, b0 + i);                 // VIMNMX3
            if (OP == 1) x[i] = max(x[i], y + i);                              // VIMNMX (2-input) (+ add folded?)
            if (OP == 2) x[i] = x[i] * b0 + y;                                 // IMAD
            if (OP == 3) x[i] = (int)__umulhi((unsigned)x[i], u) + y;          // IMAD.HI
            if (OP == 4) x[i] = (x[i] & b0) ^ (y | i);                         // LOP3
            if (OP == 5) { x[i] = __vimax3_s32(x[i], y, b0 + i); }             // mixed: set below
            if (OP == 6) x[i] = __vimax3_u16x2(x[i], y, b0 + i);               // VIMNMX3.U16x2
            if (OP == 7) x[i] = __float_as_int(fmaxf(__int_as_float(x[i]), __int_as_float(y + i)));  // FMNMX
        }
        if (OP == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y = (int)__umulhi((unsigned)(x[i] ^ y), u) + y;  // dependent IMAD.HI chain interleaved
        }
    }
    long long t1 = clock64();
    int s = y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (int)(t1 - t0);
}

template <int OP>
void run(const char *name, int nw, int opsPerIter) {
    int *d; cudaMalloc(&d, 16); int h[2];
    k<OP><<<148, nw * 32>>>(d, 3, 7, 11); cudaDeviceSynchronize();
    k<OP><<<148, nw * 32>>>(d, 3, 7, 11); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    double cyc = h[1];
    double warp_inst = (double)ITERS * opsPerIter * nw;      // per SM
    printf("%-22s warps/SM=%2d  cycles=%8.0f  warp-inst/clk/SM=%.2f  (lanes/clk/SM=%.0f)\n", name, nw, cyc, warp_inst / cyc, 32 * warp_inst / cyc);
    cudaFree(d);
}
int main() {
    for (int nw : {4, 8, 16}) {
        run<0>("VIMNMX3.S32", nw, 8); run<1>("VIMNMX.S32(+add)", nw, 8); run<2>("IMAD", nw, 8); run<3>("IMAD.HI.U32(+add)", nw, 8);
        run<4>("LOP3 x2", nw, 8); run<6>("VIMNMX3.U16x2", nw, 8); run<7>("FMNMX", nw, 8); run<5>("VIMNMX3 + IMAD.HI mix", nw, 16);
        printf("\n");
    }
    return 0;
}
