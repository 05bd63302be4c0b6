// micro-benchmark: DMMA.8x8x4 throughput when every instruction reads different operand registers, and with the operands
// loaded from shared memory (LDS.128 per instruction pair) as in the sweep kernel
#include <cstdio>
#include <cuda_runtime.h>
#define MMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))
__global__ void k_regs(double* out, int iters, const double* in) {
    double c[12][2], a[12], b[12];
    for (int i = 0; i < 12; ++i) { c[i][0] = c[i][1] = 0; a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 12; ++i) MMA(c[i], a[i], b[i]);
    double s = 0; for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_smem(double* out, int iters, const double* in) {
    __shared__ double2 sa[1024], sb[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sa[i] = make_double2(in[i], in[i + 1]); sb[i] = make_double2(in[i + 2], in[i + 3]); }
    __syncthreads();
    double c[12][2];
    for (int i = 0; i < 12; ++i) c[i][0] = c[i][1] = 0;
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 a = sa[(w * 64 + ((it * 4 + j) & 15) * 4 + (l >> 2) * 125 + (l & 3)) & 1023];
            const double2 b = sb[((it * 4 + j) * 4 + (l >> 2) * 132 + (l & 3)) & 1023];
            MMA(c[3 * j], a.x, b.x); MMA(c[3 * j + 1], a.y, b.y); MMA(c[3 * j + 2], a.x + a.y, b.x + b.y);
        }
    }
    double s = 0; for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *o, *in; cudaMalloc(&o, 8 * 148 * 1024); cudaMalloc(&in, 8 * 4096); cudaMemset(in, 0, 8 * 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 12;
    for (int warps : {1, 2, 4, 8, 12}) {
        float ms;
        k_regs<<<148, 32 * warps>>>(o, 8, in);
        cudaEventRecord(e0); k_regs<<<148, 32 * warps>>>(o, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("warps/SM %2d  distinct register operands: %6.1f FMA/clk/SM  %5.1f cycles per warp DMMA\n", warps,
               148.0 * warps * 12.0 * iters * 256 / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (12.0 * iters));
        k_smem<<<148, 32 * warps>>>(o, 8, in);
        cudaEventRecord(e0); k_smem<<<148, 32 * warps>>>(o, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("             operands from shared memory:  %6.1f FMA/clk/SM  %5.1f cycles per warp DMMA\n",
               148.0 * warps * 12.0 * iters * 256 / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (12.0 * iters));
    }
    return 0;
}
