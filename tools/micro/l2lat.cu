// micro-benchmark: latency of single L2 loads (one thread) over the 2 KB grains of a buffer that another SM has just
// written - are there near and far addresses for coherent (L1-bypassing) loads? (developer tool)
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void writer(unsigned long long* buf, int ngr) {
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < ngr; i += blockDim.x * gridDim.x)
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(buf + (size_t)i * 256), "l"((unsigned long long)i) : "memory");
}
__global__ void reader(const unsigned long long* buf, int ngr, int kind, int* lat, int target_sm) {
    unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid != target_sm || threadIdx.x != 0) return;
    if (atomicAdd(&lat[ngr], 1) != 0) return;        // first CTA on that SM only
    for (int i = 0; i < ngr; ++i) {
        const unsigned long long* p = buf + (size_t)i * 256;
        unsigned long long v;
        long long t0 = clock64();
        if (kind == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        else if (kind == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        else asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        long long t1;                                                 // the clock is read after an instruction that needs v
        asm volatile("{.reg .pred p; setp.eq.u64 p, %1, 0x123456789; @p trap; mov.u64 %0, %%clock64;}" : "=l"(t1) : "l"(v) : "memory");
        lat[i] = (int)(t1 - t0);
    }
}
int main() {
    const int ngr = 2048;          // 4 MB
    unsigned long long* buf; int* lat;
    cudaMalloc(&buf, (size_t)ngr * 2048); cudaMalloc(&lat, (ngr + 1) * 4);
    int h[ngr + 1];
    for (int sm : {0, 1, 74, 147})
        for (int kind = 0; kind < 3; ++kind) {
            writer<<<64, 128>>>(buf, ngr);
            cudaDeviceSynchronize();
            cudaMemset(lat, 0, (ngr + 1) * 4);
            reader<<<148 * 8, 32>>>(buf, ngr, kind, lat, sm);
            cudaMemcpy(h, lat, sizeof(h), cudaMemcpyDeviceToHost);
            std::sort(h, h + ngr);
            printf("SM %3d %s: min %d p10 %d p25 %d p50 %d p75 %d p90 %d p99 %d max %d\n", sm, kind == 0 ? "ld.relaxed.gpu" : (kind == 1 ? "ld.cg        " : "ld.ca        "),
                   h[0], h[ngr / 10], h[ngr / 4], h[ngr / 2], h[3 * ngr / 4], h[9 * ngr / 10], h[99 * ngr / 100], h[ngr - 1]);
        }
    return 0;
}
