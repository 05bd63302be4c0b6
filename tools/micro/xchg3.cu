// micro-benchmark: who should poll?  The per-strip hand-over of the cluster sweep kernel (see xchg.cu) with the polls issued
//   0  by the warps that have just stored (the kernel's arrangement)       1  by one extra warp (9 entries)
//   2  by three extra warps (3 entries each)                                3  as 0, first poll delayed by 600 cycles
//   4  as 2 with a second set of three warps polling the same entries half a round later (two rounds in flight)
//   5  by the TMA: one bulk copy of the 9 entry rows (4.75 KB) into shared memory per polling round, checked there
// (developer tool)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
#define NSE 384
#define PPW 33
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ size_t widx(int e, int p) { return ((size_t)e * PPW + p) * 2; }
__global__ void __cluster_dims__(4, 1, 1) xchg(unsigned long long* ring, size_t slot_stride, int rounds, int mode, long long* out, int NC) {
    const int g = blockIdx.x, l = g / 4, k = g % 4, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    __shared__ volatile int done_round[3];
    __shared__ __align__(128) unsigned long long stage[9 * PPW * 2 + 16];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int all_ok;
    if (mode == 5 && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3) done_round[tid] = -1;
    __syncthreads();
    long long t0 = clock64(), tpoll = 0, nr = 0;
    const bool storer = w < 3;
    const bool poller = (mode == 0 || mode == 3) ? w < 3 : (mode == 1 ? w == 3 : (mode == 2 ? (w >= 3 && w < 6) : w >= 3));
    const int pw = (mode == 0 || mode == 3) ? w : (w - 3) % 3;          // polling warp index 0..2
    for (int r = 0; r < rounds; ++r) {
        unsigned long long* slot = ring + (size_t)(r & 3) * slot_stride;
        unsigned long long* nxt = ring + (size_t)((r + 1) & 3) * slot_stride;
        if (storer && l < NC) { put(slot + widx(96 * k + tid, l), (unsigned long long)r + 1); put(nxt + widx(96 * k + tid, l), SENT); }
        if (mode == 5) {
            long long tp = clock64();
            const int e0 = (l - 1) * 12 + 9 * k;                      // entries e0 .. e0+8 (clipped to 0..NSE-1)
            const int ea = max(e0, 0), eb = min(e0 + 9, NSE);
            const unsigned bytes = eb > ea ? (unsigned)((eb - ea) * PPW * 16) : 0u;
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar), st = (unsigned)__cvta_generic_to_shared(stage);
            int spin = 0;
            for (;;) {
                if (bytes) {
                    if (tid == 0) {
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(st), "l"(slot + widx(ea, 0)), "r"(bytes), "r"(mb) : "memory");
                    }
                    const unsigned par = (unsigned)(nr & 1);
                    asm volatile("{\n.reg .pred P1;\nW5:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D5;\nbra W5;\nD5:\n}" ::"r"(mb), "r"(par) : "memory");
                }
                ++nr;
                bool ok = true;
                for (int i = tid; i < (eb - ea) * PPW; i += 96) {
                    if (i % PPW < 32) ok = ok && stage[2 * i] != SENT && stage[2 * i + 1] != SENT;
                }
                int cnt = __syncthreads_and(ok);
                if (cnt) break;
                if (++spin > 2000000) { out[3 * 132] = 1; break; }
            }
            tpoll += clock64() - tp;
        } else if (poller) {
            long long tp = clock64();
            if (mode == 3) { while (clock64() - tp < 600) {} }
            if (mode == 4 && w >= 6) { while (clock64() - tp < 500) {} }
            const int ne = mode == 1 ? 9 : 3;
            int ent[9];
#pragma unroll
            for (int o = 0; o < 9; ++o) { int e = (l - 1) * 12 + 9 * k + (mode == 1 ? o : pw + 3 * o); ent[o] = (o < ne && e >= 0 && e < 12 * NC) ? e : -1; }
            int spin = 0;
            for (;;) {
                unsigned long long lo[9], hi[9];
                bool ok = true;
#pragma unroll
                for (int o = 0; o < 9; ++o) { lo[o] = hi[o] = 0; if (ent[o] >= 0 && lane < NC) ld2(slot + widx(ent[o], lane), lo[o], hi[o]); }
#pragma unroll
                for (int o = 0; o < 9; ++o) ok = ok && lo[o] != SENT && hi[o] != SENT;
                ++nr;
                if (mode == 4 && done_round[pw] >= r) break;             // the other warp with the same entries has them
                if (__all_sync(0xffffffffu, ok)) break;
                if (++spin > 2000000) { out[3 * 132] = 1; break; }
            }
            if (mode == 4 && lane == 0) done_round[pw] = r;
            tpoll += clock64() - tp;
        }
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (lane == 0 && (poller || mode == 5) && pw == 0 && (mode != 4 || w == 3)) { out[3 * g] = clock64() - t0; out[3 * g + 1] = tpoll; out[3 * g + 2] = nr; }
}
int main() {
    int rounds = 4000;
    size_t slot_stride = (size_t)NSE * PPW * 2 + 1024;
    unsigned long long* ring; long long* out;
    cudaMalloc(&ring, 4 * slot_stride * 8); cudaMalloc(&out, (132 * 3 + 1) * 8);
    for (int NC : {32, 24, 16, 8, 4, 2}) {
        int mode = 0;
        cudaMemset(ring, 0xFF, 4 * slot_stride * 8); cudaMemset(out, 0, (132 * 3 + 1) * 8);
        void* args[] = {&ring, &slot_stride, &rounds, &mode, &out, &NC};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)xchg, dim3(4 * (NC + 1)), dim3(96), args, 0, 0);
        long long h[132 * 3 + 1]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        if (h[132 * 3]) { printf("NC %d: runaway spin\n", NC); fflush(stdout); continue; }
        double per = 0, pol = 0, nr = 0; int cnt = 0; for (int g = 4; g < 4 * NC - 4; ++g) { per += h[3 * g]; pol += h[3 * g + 1]; nr += h[3 * g + 2]; ++cnt; }
        if (!cnt) { cnt = 1; per = h[3 * 4]; pol = h[3 * 4 + 1]; nr = h[3 * 4 + 2]; }
        printf("%2d producer clusters (%3d CTAs): period %.0f cycles, poll %.0f cycles, %.2f polling rounds -> %.0f cycles per polling round (%s)\n", NC, 4 * (NC + 1),
               per / cnt / rounds, pol / cnt / rounds, nr / cnt / rounds, pol / nr, cudaGetErrorString(e));
        fflush(stdout);
    }
    return 0;
}
