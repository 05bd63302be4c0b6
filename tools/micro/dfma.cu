// micro-benchmark: FP64 FMA throughput and shuffle/FMA latency on this GPU (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void dfma(double* out, int iters, double a, double b) {
    double x[CH];
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma(float* out, int iters, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void lat(double* out, long long* cyc, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = fma(x, a, b);
    long long t1 = clock64();
    double y = x;
    for (int it = 0; it < iters; ++it) y += __shfl_xor_sync(0xffffffffu, y, 1);
    long long t2 = clock64();
    out[threadIdx.x] = y; cyc[0] = t1 - t0; cyc[1] = t2 - t1;
}
int main() {
    double* o; cudaMalloc(&o, 8 * 148 * 8 * 1024); long long* c; cudaMalloc(&c, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 14;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma<8><<<148 * 4, 256>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 4 * 256 * 8.0 * iters;
        printf("DFMA: %.3f ms, %.2f TFLOP/s (2 flop/fma), %.1f FMA/clk/SM @1.9GHz\n", ms, 2 * fl / ms / 1e9, fl / (ms * 1e-3) / 148 / 1.9e9);
        cudaEventRecord(e0); ffma<<<148 * 4, 256>>>((float*)o, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM\n", ms, 2 * fl / ms / 1e9, fl / (ms * 1e-3) / 148 / 1.9e9);
    }
    lat<<<1, 32>>>(o, c, 4096, 1.0000001, 1e-9); long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency %.1f cyc, shfl+dadd dependent %.1f cyc\n", h[0] / 4096.0, h[1] / 4096.0);
    return 0;
}
