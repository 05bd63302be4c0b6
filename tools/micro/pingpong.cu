// micro-benchmark: latency of a flag-free hand-over through L2 between two CTAs (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pp(volatile unsigned long long* buf, int iters, long long* out) {
    // CTA 0 writes buf[0] = i, waits buf[32] == i ; CTA 1 waits buf[0] == i, writes buf[32] = i
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    for (unsigned long long i = 1; i <= (unsigned long long)iters; ++i) {
        if (blockIdx.x == 0) { buf[0] = i; while (buf[32] != i) {} }
        else if (blockIdx.x == 1) { while (buf[0] != i) {} buf[32] = i; }
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
}
int main() {
    unsigned long long* b; cudaMalloc(&b, 4096); cudaMemset(b, 0, 4096);
    long long* o; cudaMalloc(&o, 8);
    void* args[] = {&b, nullptr, &o};
    int iters = 20000; args[1] = &iters;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(b, 0, 4096);
        cudaLaunchCooperativeKernel((const void*)pp, dim3(148), dim3(32), args, 0, 0);
        long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
        printf("ping-pong round trip (2 hand-overs): %.0f cycles -> one hand-over %.0f cycles\n", (double)h / iters, (double)h / iters / 2);
    }
    return 0;
}
