// micro-benchmark: the per-strip hand-over of the cluster sweep kernel with DIE-LOCAL replicas of the exchange slot
// (developer tool).  B200 is two dies; an L2 line lives on one of them (2 KB address grains, hashed), and a load from the
// other die costs ~700 instead of ~300 cycles (tools/micro/l2lat.cu).  Calibration: one SM classifies every grain of a
// pool as near/far, every SM then classifies itself against known grains.  The exchange is run (a) with one copy of
// every word (grains taken as they come), (b) with two copies, one per die, every reader polling the copy on its own die.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
#define NSE 384
#define GPS 128                 // grains per slot and replica (96 used: 4 entries of 32 words per grain)
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ int timed_load(const unsigned long long* p) {
    unsigned long long v; long long t0, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    asm volatile("{.reg .pred p; setp.eq.u64 p, %1, 0x123456789; @p trap; mov.u64 %0, %%clock64;}" : "=l"(t1) : "l"(v) : "memory");
    return (int)(t1 - t0);
}
__global__ void calib_grains(unsigned long long* pool, int ngr, int* lat, int* smid_out) {
    if (threadIdx.x) return;
    unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); *smid_out = (int)smid;
    for (int i = 0; i < ngr; ++i) pool[(size_t)i * 256] = 0;
    for (int rep = 0; rep < 4; ++rep)
        for (int i = 0; i < ngr; ++i) { int t = timed_load(pool + (size_t)i * 256); if (rep == 1 || t < lat[i]) lat[i] = t; }
}
__global__ void calib_sms(const unsigned long long* pool, const int* near8, const int* far8, int* sm_die, int* taken) {
    unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x || atomicAdd(&taken[smid], 1)) return;
    int tn = 1 << 30, tf = 1 << 30;
    for (int rep = 0; rep < 4; ++rep)
        for (int i = 0; i < 8; ++i) { int a = timed_load(pool + (size_t)near8[i] * 256), b = timed_load(pool + (size_t)far8[i] * 256); if (rep) { tn = min(tn, a); tf = min(tf, b); } }
    sm_die[smid] = tn < tf ? 0 : 1;
}
// tbl[rep][slot][grain] -> grain of the pool
__global__ void __cluster_dims__(4, 1, 1) xchg(unsigned long long* pool, const int* tbl, const int* sm_die, int nrep, int rounds, long long* out) {
    const int g = blockIdx.x, l = g / 4, k = g % 4, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    const int my = nrep == 2 ? sm_die[smid] : 0;
    long long t0 = clock64(), tpoll = 0, nr = 0;
    for (int r = 0; r < rounds; ++r) {
        const int s = r & 3, sn = (r + 1) & 3;
        if (l < 32) {
            const int e = 96 * k + tid;
            for (int rep = 0; rep < nrep; ++rep) {
                put(pool + (size_t)tbl[(rep * 4 + s) * GPS + (e >> 2)] * 256 + ((e & 3) * 32 + l) * 2, (unsigned long long)r + 1);
                put(pool + (size_t)tbl[(rep * 4 + sn) * GPS + (e >> 2)] * 256 + ((e & 3) * 32 + l) * 2, SENT);
            }
        }
        long long tp = clock64();
        const unsigned long long* src[3];
        for (int o = 0; o < 3; ++o) {
            int e = (l - 1) * 12 + 9 * k + w + 3 * o;
            src[o] = (e >= 0 && e < NSE) ? pool + (size_t)tbl[(my * 4 + s) * GPS + (e >> 2)] * 256 + ((e & 3) * 32 + lane) * 2 : nullptr;
        }
        int spin = 0;
        for (;;) {
            unsigned long long lo[3], hi[3];
            bool ok = true;
            for (int o = 0; o < 3; ++o) { lo[o] = hi[o] = 0; if (src[o]) ld2(src[o], lo[o], hi[o]); }
            for (int o = 0; o < 3; ++o) ok = ok && lo[o] != SENT && hi[o] != SENT;
            ++nr;
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spin > 2000000) { out[3 * 132] = 1; break; }
        }
        tpoll += clock64() - tp;
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (tid == 0) { out[3 * g] = clock64() - t0; out[3 * g + 1] = tpoll; out[3 * g + 2] = nr; }
}
int main() {
    const int ngr = 4096, rounds = 4000;
    unsigned long long* pool; int *lat, *smid0, *near8, *far8, *sm_die, *taken, *tbl; long long* out;
    cudaMalloc(&pool, (size_t)ngr * 2048); cudaMalloc(&lat, ngr * 4); cudaMalloc(&smid0, 4); cudaMalloc(&near8, 32); cudaMalloc(&far8, 32);
    cudaMalloc(&sm_die, 256 * 4); cudaMalloc(&taken, 256 * 4); cudaMalloc(&tbl, 2 * 4 * GPS * 4); cudaMalloc(&out, (132 * 3 + 1) * 8);
    cudaMemset(taken, 0, 256 * 4); cudaMemset(sm_die, 0xFF, 256 * 4);
    calib_grains<<<1, 32>>>(pool, ngr, lat, smid0);
    std::vector<int> hl(ngr); int s0; cudaMemcpy(hl.data(), lat, ngr * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&s0, smid0, 4, cudaMemcpyDeviceToHost);
    std::vector<int> nearg, farg;
    int lo = 1 << 30, hi = 0; for (int v : hl) { lo = std::min(lo, v); hi = std::max(hi, v); }
    const int thr = (lo + hi) / 2;
    for (int i = 0; i < ngr; ++i) (hl[i] < thr ? nearg : farg).push_back(i);
    printf("calibration SM %d: latency %d..%d cycles, threshold %d: %zu near, %zu far grains\n", s0, lo, hi, thr, nearg.size(), farg.size());
    cudaMemcpy(near8, nearg.data(), 32, cudaMemcpyHostToDevice); cudaMemcpy(far8, farg.data(), 32, cudaMemcpyHostToDevice);
    calib_sms<<<148 * 8, 32>>>(pool, near8, far8, sm_die, taken);
    int hd[256]; cudaMemcpy(hd, sm_die, sizeof(hd), cudaMemcpyDeviceToHost);
    int n0 = 0, n1 = 0; for (int i = 0; i < 148; ++i) { n0 += hd[i] == 0; n1 += hd[i] == 1; }
    printf("SMs on the die of SM %d: %d, on the other die: %d (unclassified %d)\n", s0, n0, n1, 148 - n0 - n1);
    for (int cfg = 0; cfg < 3; ++cfg) {
        // 0: one copy, grains as they come (mixed dies); 1: one copy, all on die 0; 2: two copies, die-local reads
        std::vector<int> ht(2 * 4 * GPS);
        for (int rep = 0; rep < 2; ++rep)
            for (int i = 0; i < 4 * GPS; ++i)
                ht[rep * 4 * GPS + i] = cfg == 0 ? i : (rep == 0 ? nearg[i] : farg[i]);
        cudaMemcpy(tbl, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(pool, 0xFF, (size_t)ngr * 2048); cudaMemset(out, 0, (132 * 3 + 1) * 8);
        int nrep = cfg == 2 ? 2 : 1, rr = rounds;
        void* args[] = {&pool, &tbl, &sm_die, &nrep, &rr, &out};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)xchg, dim3(132), dim3(96), args, 0, 0);
        long long h[132 * 3 + 1]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); if (h[132 * 3]) { printf("runaway spin\n"); return 1; }
        double per = 0, pol = 0, nr = 0; for (int g = 4; g < 124; ++g) { per += h[3 * g]; pol += h[3 * g + 1]; nr += h[3 * g + 2]; }
        const char* nm[] = {"one copy, mixed dies   ", "one copy, all on die 0 ", "two copies, local reads"};
        printf("%s: period %.0f cycles, poll %.0f cycles, %.2f polling rounds -> %.0f cycles per polling round (%s)\n", nm[cfg], per / 120 / rounds,
               pol / 120 / rounds, nr / 120 / rounds, pol / nr, cudaGetErrorString(e));
    }
    return 0;
}
