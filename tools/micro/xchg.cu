// micro-benchmark: the per-strip hand-over of the cluster sweep kernel in isolation (developer tool).
// 128 producer CTAs (32 "clusters" x 4) publish 96 self-validating 16-byte words each (rows of N rho for their separator
// columns); every CTA then gathers 9 entries x 32 producers, and starts the next round when it has them (plus an optional
// compute delay).  Layouts of the exchange slot:
//   0  entry-major [e][PP] 16-byte words (the kernel's): scattered stores, one coalesced 512-byte poll per entry
//   1  entry-major, one word per 32-byte sector
//   2  producer-major [p][NS]: coalesced stores, polls gather 32 sectors
//   3  entry-major, one word per 128-byte line
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
#define NSE 384
#define PPW 33
__device__ int g_st = 0, g_ld = 0;      // instruction flavours (set by the host)
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v) {
    switch (g_st) {
        case 0: asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); break;
        case 1: asm volatile("st.global.cg.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); break;
        case 2: asm volatile("st.volatile.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); break;
        default: asm volatile("st.release.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); break;
    }
}
__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& lo, unsigned long long& hi) {
    switch (g_ld) {
        case 0: asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory"); break;
        case 1: asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory"); break;
        case 2: asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory"); break;
        default: asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory"); break;
    }
}
__device__ __forceinline__ size_t widx(int mode, int e, int p) {      // index in 8-byte units
    switch (mode) {
        case 0: return ((size_t)e * PPW + p) * 2;
        case 1: return ((size_t)e * PPW + p) * 4;
        case 2: return ((size_t)p * NSE + e) * 2;
        default: return ((size_t)e * PPW + p) * 16;
    }
}
__global__ void __cluster_dims__(4, 1, 1) xchg(unsigned long long* ring, size_t slot_stride, int rounds, int mode, int delay, int nostore_arm, long long* out, int pk, int nent) {
    const int g = blockIdx.x, l = g / 4, k = g % 4, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    long long t0 = clock64(), tpoll = 0, nr = 0, textra = 0;
    for (int r = 0; r < rounds; ++r) {
        unsigned long long* slot = ring + (size_t)(r & 3) * slot_stride;
        unsigned long long* nxt = ring + (size_t)((r + 1) & 3) * slot_stride;
        if (l < 32) {
            put(slot + widx(mode, 96 * k + tid, l), (unsigned long long)r + 1);
            if (!nostore_arm) put(nxt + widx(mode, 96 * k + tid, l), SENT);
        }
        long long tp = clock64();
        int ent[3];
        for (int o = 0; o < 3; ++o) { int e = (l - 1) * 12 + 9 * k + w + 3 * o; ent[o] = (e >= 0 && e < NSE && k < pk && o < nent) ? e : -1; }
        int spin = 0;
        for (;;) {
            unsigned long long lo[3], hi[3];
            bool ok = true;
            for (int o = 0; o < 3; ++o) { lo[o] = hi[o] = (unsigned long long)r + 1; if (ent[o] >= 0) ld2(slot + widx(mode, ent[o], lane), lo[o], hi[o]); }
            for (int o = 0; o < 3; ++o) ok = ok && (nostore_arm ? (lo[o] == (unsigned long long)r + 1) : (lo[o] != SENT && hi[o] != SENT));
            ++nr;
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spin > 2000000) { out[3 * 132] = 1; break; }
        }
        tpoll += clock64() - tp;
        {   // one more identical round on words that are all valid by now
            long long tq; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tq) :: "memory");
            unsigned long long lo[3], hi[3], acc = 0;
            for (int o = 0; o < 3; ++o) { lo[o] = hi[o] = 1; if (ent[o] >= 0) ld2(slot + widx(mode, ent[o], lane), lo[o], hi[o]); }
            for (int o = 0; o < 3; ++o) acc |= lo[o] ^ hi[o];
            long long tq1;
            asm volatile("{.reg .pred p; setp.eq.u64 p, %1, 0x123456789; @p trap; mov.u64 %0, %%clock64;}" : "=l"(tq1) : "l"(acc) : "memory");
            textra += tq1 - tq;
        }
        // the CTAs of a cluster exchange the gathered entries (DSMEM in the kernel): nobody is more than one round ahead
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (delay) { long long t1 = clock64(); while (clock64() - t1 < delay) {} }
    }
    if (tid == 0) { out[3 * g] = clock64() - t0; out[3 * g + 1] = tpoll; out[3 * g + 2] = nr; out[3 * 132 + 1 + g] = textra; }
}
int main(int argc, char** argv) {
    int rounds = 4000;
    size_t slot_stride = (size_t)NSE * PPW * 16 + 1024;        // 8-byte units, enough for every layout
    unsigned long long* ring; long long* out;
    cudaMalloc(&ring, 4 * slot_stride * 8); cudaMalloc(&out, (132 * 4 + 1) * 8); cudaMemset(out, 0, (132 * 4 + 1) * 8);
    int nostore_arm = 0, delay = 0, mode = 0, stk = 0, ldk = 0;
    cudaMemcpyToSymbol(g_st, &stk, 4); cudaMemcpyToSymbol(g_ld, &ldk, 4);
    for (int pk = 4; pk >= 1; pk >>= 1)
    for (int nent = 3; nent >= 1; nent -= 2) {
        cudaMemset(ring, 0xFF, 4 * slot_stride * 8);
        void* args[] = {&ring, &slot_stride, &rounds, &mode, &delay, &nostore_arm, &out, &pk, &nent};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)xchg, dim3(132), dim3(96), args, 0, 0);
        long long h[132 * 4 + 1]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); if (h[132 * 3]) { printf("runaway spin\n"); return 1; }
        double per = 0, pol = 0, nr = 0, ex = 0; int cnt = 0;
        for (int g = 4; g < 124; ++g) if (g % 4 < pk) { per += h[3 * g]; pol += h[3 * g + 1]; nr += h[3 * g + 2]; ex += h[3 * 132 + 1 + g]; ++cnt; }
        printf("%d of 4 CTAs per cluster poll %d entries per warp: period %.0f cycles, poll %.0f cycles, %.2f polling rounds -> %.0f cycles per polling round; a round on valid words %.0f (%s)\n",
               pk, nent, per / cnt / rounds, pol / cnt / rounds, nr / cnt / rounds, pol / nr, ex / cnt / rounds, cudaGetErrorString(e));
    }
    return 0;
}
