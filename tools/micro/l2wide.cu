// micro-benchmark: latency of warp-wide L1-bypassing loads from L2 by footprint (developer tool).
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void writer(unsigned long long* buf, size_t n) {
    for (size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < n; i += (size_t)blockDim.x * gridDim.x) buf[i] = i;
}
// kind: 0 one lane 8 B; 1 32 lanes x 8 B; 2 32 lanes x 16 B; 3 three loads of 32 x 16 B (different 512-byte rows); strong: ld.relaxed.gpu or ld.cg
__global__ void reader(const unsigned long long* buf, int nrep, int kind, int strong, int* lat) {
    const int lane = threadIdx.x;
    for (int r = 0; r < nrep; ++r) {
        const unsigned long long* p = buf + (size_t)r * 4096;      // 32 KB apart
        unsigned long long a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0;
        long long t0, t1;
        __syncwarp();
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
        if (kind == 0) { if (lane == 0) { if (strong) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a0) : "l"(p) : "memory"); else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(a0) : "l"(p) : "memory"); } }
        else if (kind == 1) { if (strong) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a0) : "l"(p + lane) : "memory"); else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(a0) : "l"(p + lane) : "memory"); }
        else {
            if (strong) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p + 2 * lane) : "memory");
            else asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p + 2 * lane) : "memory");
            if (kind == 3) {
                if (strong) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "l"(p + 1024 + 2 * lane) : "memory");
                              asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "l"(p + 2048 + 2 * lane) : "memory"); }
                else { asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "l"(p + 1024 + 2 * lane) : "memory");
                       asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "l"(p + 2048 + 2 * lane) : "memory"); }
            }
        }
        unsigned long long acc = a0 ^ a1 ^ b0 ^ b1 ^ c0 ^ c1;
        bool all = __any_sync(0xffffffffu, acc == 0x123456789ull);
        asm volatile("{.reg .pred p; setp.eq.u32 p, %1, 1; @p trap; mov.u64 %0, %%clock64;}" : "=l"(t1) : "r"((int)all) : "memory");
        if (lane == 0) lat[r] = (int)(t1 - t0);
    }
}
int main() {
    const int nrep = 512; size_t n = (size_t)nrep * 4096 + 8192;
    unsigned long long* buf; int* lat; cudaMalloc(&buf, n * 8); cudaMalloc(&lat, nrep * 4);
    int h[nrep];
    const char* kn[] = {"1 lane x 8 B       ", "32 lanes x 8 B     ", "32 lanes x 16 B    ", "3 x 32 lanes x 16 B"};
    for (int strong = 1; strong >= 0; --strong)
        for (int kind = 0; kind < 4; ++kind) {
            writer<<<296, 256>>>(buf, n); cudaDeviceSynchronize();
            reader<<<1, 32>>>(buf, nrep, kind, strong, lat);
            { cudaError_t e = cudaMemcpy(h, lat, sizeof(h), cudaMemcpyDeviceToHost); if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e)); }
            std::sort(h, h + nrep);
            printf("%s %s: min %d p25 %d p50 %d p75 %d p90 %d max %d\n", strong ? "ld.relaxed.gpu" : "ld.cg         ", kn[kind], h[0], h[nrep / 4], h[nrep / 2], h[3 * nrep / 4], h[9 * nrep / 10], h[nrep - 1]);
        }
    return 0;
}
