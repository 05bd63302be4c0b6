// micro-benchmark: throughput of the FP64 tensor-core instruction DMMA.8x8x4 (mma.sync m8n8k4 f64) next to plain DFMA,
// for W warps per SM (developer tool; built by hand: nvcc -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu)
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void dmma(double* out, int iters, double a, double b) {
    double c[CH][2];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __global__ void dfma(double* out, int iters, double a, double b) {
    double x[CH];
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* o; cudaMalloc(&o, 8 * 148 * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 13;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        float ms;
        cudaEventRecord(e0); dmma<8><<<148, 32 * warps>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fma_ = 148.0 * warps * 8.0 * iters * 256.0;          // 8x8x4 = 256 FMAs per instruction
        printf("warps/SM %2d  DMMA: %7.3f ms  %6.2f TFLOP/s  %6.1f FMA/clk/SM  %5.2f cycles per warp instruction\n", warps, ms, 2 * fma_ / ms / 1e9,
               fma_ / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (8.0 * iters));
        cudaEventRecord(e0); dfma<16><<<148, 32 * warps>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fma_ = 148.0 * warps * 32 * 16.0 * iters;
        printf("             DFMA: %7.3f ms  %6.2f TFLOP/s  %6.1f FMA/clk/SM\n", ms, 2 * fma_ / ms / 1e9, fma_ / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
// second experiment (dmma mix): do DMMA and DFMA share one datapath?  Half of the warps of every SM issue DMMA, the other half DFMA.
