// micro-benchmark: hand-over latency through L2 between SM pairs (die map), and the cost of sharing a 128-byte line
// between writers on both dies (developer tool).  One CTA per SM (cooperative launch, 148 CTAs).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ void stg(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ldg(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
// phase j: CTA 0 and CTA j play ping-pong on two lines (a: 0 -> j, b: j -> 0); everybody else waits for the phase counter
__global__ void pairs(unsigned long long* buf, int iters, long long* out, int* smids) {
    const int g = blockIdx.x, G = gridDim.x;
    if (threadIdx.x) return;
    unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); smids[g] = (int)smid;
    unsigned long long* a = buf; unsigned long long* b = buf + 64; unsigned long long* phase = buf + 128;
    for (int j = 1; j < G; ++j) {
        if (g == 0) {
            while (ldg(phase) != (unsigned long long)(2 * j - 1)) {}      // partner ready
            long long t0 = clock64();
            for (int i = 1; i <= iters; ++i) { unsigned long long v = (unsigned long long)j * 1000000 + i; stg(a, v); while (ldg(b) != v) {} }
            out[j] = clock64() - t0;
            stg(phase, 2 * j);
        } else if (g == j) {
            while (ldg(phase) != (unsigned long long)(2 * j - 2)) {}
            stg(phase, 2 * j - 1);
            for (int i = 1; i <= iters; ++i) { unsigned long long v = (unsigned long long)j * 1000000 + i; while (ldg(a) != v) {} stg(b, v); }
        }
    }
}
// one reader (CTA r) polls word 0 of a line; the producer (CTA p) writes word 0 after a "noise" writer (CTA q) wrote word 8
// (same line) or word 16 (next line... 128 B apart) just before.  Round trip as above with the reader answering on its own line.
__global__ void share(unsigned long long* buf, int iters, int p, int q, int r, int same_line, long long* out) {
    const int g = blockIdx.x;
    if (threadIdx.x) return;
    unsigned long long* x = buf + 1024;                   // word 0 of the data line
    unsigned long long* nz = x + (same_line ? 8 : 16 * 16);  // noise word: same line (64 B further) or another line
    unsigned long long* back = buf + 2048;                // reader -> producer
    unsigned long long* tok = buf + 3072;                 // producer -> noise writer -> producer
    if (g == p) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            stg(tok, i); while (ldg(tok + 16) != (unsigned long long)i) {}       // noise writer has written its word
            stg(x, i); while (ldg(back) != (unsigned long long)i) {}
        }
        out[0] = clock64() - t0;
    } else if (g == q) {
        for (int i = 1; i <= iters; ++i) { while (ldg(tok) != (unsigned long long)i) {} stg(nz, i); stg(tok + 16, i); }
    } else if (g == r) {
        for (int i = 1; i <= iters; ++i) { while (ldg(x) != (unsigned long long)i) {} stg(back, i); }
    }
}
int main() {
    unsigned long long* buf; long long* out; int* smids;
    cudaMalloc(&buf, 1 << 20); cudaMalloc(&out, 256 * 8); cudaMalloc(&smids, 256 * 4);
    cudaMemset(buf, 0, 1 << 20);
    int iters = 200;
    void* args[] = {&buf, &iters, &out, &smids};
    cudaFuncSetAttribute(pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)pairs, dim3(148), dim3(32), args, 200 * 1024, 0);
    long long h[256]; int sm[256];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(sm, smids, sizeof(sm), cudaMemcpyDeviceToHost);
    printf("pairs: %s; CTA 0 on SM %d\n", cudaGetErrorString(e), sm[0]);
    std::vector<double> v; for (int j = 1; j < 148; ++j) v.push_back((double)h[j] / iters / 2);
    std::vector<double> s = v; std::sort(s.begin(), s.end());
    printf("one hand-over (cycles): min %.0f p25 %.0f p50 %.0f p75 %.0f max %.0f\n", s[0], s[s.size() / 4], s[s.size() / 2], s[3 * s.size() / 4], s.back());
    double thr = (s[0] + s.back()) / 2; int nn = 0; for (double x : v) nn += x < thr;
    printf("threshold %.0f: %d partners near, %d far\n", thr, nn, 147 - nn);
    for (int j = 1; j < 148; ++j) printf("%d:%d:%.0f ", j, sm[j], v[j - 1]);
    printf("\n");
    // pick a near and a far partner of CTA 0
    int nearj = -1, farj = -1, near2 = -1, far2 = -1;
    for (int j = 1; j < 148; ++j) { if (v[j - 1] < thr) { if (nearj < 0) nearj = j; else if (near2 < 0) near2 = j; } else { if (farj < 0) farj = j; else if (far2 < 0) far2 = j; } }
    cudaFuncSetAttribute(share, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* nm; int p, q, r; } cs[] = {{"producer near reader, noise writer near", 0, near2, nearj}, {"producer near reader, noise writer far ", 0, farj, nearj},
                                                  {"producer far from reader, noise near producer", 0, nearj, farj}, {"producer far from reader, noise near reader  ", 0, far2, farj}};
    for (auto& c : cs)
        for (int same = 0; same < 2; ++same) {
            cudaMemset(buf, 0, 1 << 20);
            int it2 = 2000;
            void* a2[] = {&buf, &it2, &c.p, &c.q, &c.r, &same, &out};
            e = cudaLaunchCooperativeKernel((const void*)share, dim3(148), dim3(32), a2, 200 * 1024, 0);
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            printf("%s, noise on %s: %.0f cycles per round (token round trip + data round trip) (%s)\n", c.nm, same ? "the SAME line " : "another line  ", (double)h[0] / it2, cudaGetErrorString(e));
        }
    return 0;
}
