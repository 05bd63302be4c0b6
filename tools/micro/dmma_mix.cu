// micro-benchmark: do DMMA.8x8x4 and DFMA share one FP64 datapath?  Warps 0..W/2-1 of every CTA issue DMMA, the others DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(double* out, int iters, double a, double b, int mode) {   // mode 0: both, 1: DMMA warps only, 2: DFMA warps only
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double c[8][2], s = 0;
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    if (w < nw / 2) {
        if (mode != 2)
            for (int it = 0; it < iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    } else {
        if (mode != 1)
            for (int it = 0; it < iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
    }
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* o; cudaMalloc(&o, 8 * 148 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 13;
    for (int warps : {8, 16}) for (int mode = 0; mode < 3; ++mode) {
        float ms;
        mix<<<148, 32 * warps>>>(o, 16, 1.0000001, 1e-9, mode);
        cudaEventRecord(e0); mix<<<148, 32 * warps>>>(o, iters, 1.0000001, 1e-9, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dm = mode != 2 ? 148.0 * (warps / 2) * 8.0 * iters * 256.0 : 0, df = mode != 1 ? 148.0 * (warps / 2) * 32 * 16.0 * iters : 0;
        printf("warps/SM %2d mode %d (%s): %7.3f ms  DMMA %6.1f + DFMA %6.1f = %6.1f FMA/clk/SM\n", warps, mode, mode == 0 ? "both" : mode == 1 ? "DMMA only" : "DFMA only", ms,
               dm / (ms * 1e-3) / 148 / 1.965e9, df / (ms * 1e-3) / 148 / 1.965e9, (dm + df) / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
