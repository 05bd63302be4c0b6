// micro-benchmark: period of an all-gather through self-validating words in L2 (developer tool).
// NP producer CTAs publish W words each per round; NC consumer CTAs (the first NC) read all NP*W words; round r+1 of a
// producer starts when it has (as a consumer) seen round r complete.  Ring of 4 slots, re-arm two rounds later.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ bool tryget(const unsigned long long* p, unsigned long long& v) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    v = lo; return lo != SENT && hi != SENT;
}
__global__ void ag(unsigned long long* ring, int NP, int W, int rounds, long long* out, int mode) {
    // ring[slot][p][w] 16 bytes each
    const int g = blockIdx.x, tid = threadIdx.x;
    const size_t slot_stride = (size_t)NP * W * 2;
    long long t0 = clock64();
    __shared__ unsigned long long sink[1024];
    for (int r = 0; r < rounds; ++r) {
        unsigned long long* slot = ring + (size_t)(r & 3) * slot_stride;
        unsigned long long* arm = ring + (size_t)((r + 2) & 3) * slot_stride;
        if (g < NP && tid < W) { put(slot + ((size_t)g * W + tid) * 2, (unsigned long long)r + 1); put(arm + ((size_t)g * W + tid) * 2, SENT); }
        // consume
        int total = NP * W;
        if (mode == 0) {            // every thread polls its own words
            for (int e = tid; e < total; e += blockDim.x) {
                unsigned long long v;
                while (!tryget(slot + (size_t)e * 2, v)) {}
                sink[tid] = v;
            }
        } else {                    // one word per producer polled, then bulk read
            for (int p = tid; p < NP; p += blockDim.x) { unsigned long long v; while (!tryget(slot + ((size_t)p * W + W - 1) * 2, v)) {} }
            __syncthreads();
            for (int e = tid; e < total; e += blockDim.x) { unsigned long long v; while (!tryget(slot + (size_t)e * 2, v)) {} sink[tid] = v; }
        }
        __syncthreads();
    }
    if (tid == 0) out[g] = clock64() - t0;
}
int main(int argc, char** argv) {
    int rounds = 4000;
    unsigned long long* ring; long long* out;
    cudaMalloc(&ring, 4 * 148 * 64 * 16); cudaMalloc(&out, 148 * 8);
    int cfgs[][4] = {{148, 24, 256, 0}, {37, 24, 256, 0}, {37, 24, 256, 1}, {37, 24, 128, 0}, {148, 3, 256, 0}, {8, 24, 256, 0}, {2, 24, 32, 0}};
    for (auto& c : cfgs) {
        int NP = c[0], W = c[1], thr = c[2], mode = c[3];
        cudaMemset(ring, 0xFF, 4 * 148 * 64 * 16);
        void* args[] = {&ring, &NP, &W, &rounds, &out, &mode};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)ag, dim3(NP), dim3(thr), args, 0, 0);
        long long h[148]; cudaMemcpy(h, out, 8 * NP, cudaMemcpyDeviceToHost);
        printf("all-gather NP=%3d CTAs x %2d words, %3d threads, mode %d: %.0f cycles per round (%s)\n", NP, W, thr, mode, (double)h[0] / rounds, cudaGetErrorString(e));
    }
    return 0;
}
