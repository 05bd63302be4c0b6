// Krylov vector kernels: the numpy calls of scipy's gmres inner loop (np.vdot, np.linalg.norm, w -= h*v,
// x += y @ V; scipy/sparse/linalg/_isolve/iterative.py as called from /root/reference/code.py:516).
//
// Reductions are single-launch and deterministic: every CTA reduces a fixed slice with warp shuffles and
// writes one partial; the CTA that draws the last ticket sums the partials in index order.  The partial buffer
// belongs to the (device, stream) pair of the call.
#include "hp_internal.cuh"

#include <stdlib.h>

#define HP_RED_THREADS 256
#define HP_RED_MAX_CTAS 1184   // 8 per SM on a 148-SM part

// Reduction scratch (partials of the CTAs + the ticket): one per (device, stream).  Two reductions in flight on the same
// stream are ordered by the stream; different streams or devices get their own buffers.
#include <map>
#include <mutex>
struct HpRedScratch { cplx* partials = nullptr; unsigned int* ticket = nullptr; };
static std::map<std::pair<int, cudaStream_t>, HpRedScratch> g_red;
static std::mutex g_red_mu;

static int hp_red_scratch(cudaStream_t st, HpRedScratch& out) {
    int dev = 0;
    HP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_red_mu);
    auto it = g_red.find({dev, st});
    if (it == g_red.end()) {
        HpRedScratch r;
        HP_CUDA(cudaMalloc(&r.partials, sizeof(cplx) * HP_RED_MAX_CTAS));
        if (cudaMalloc(&r.ticket, sizeof(unsigned int)) != cudaSuccess || cudaMemset(r.ticket, 0, sizeof(unsigned int)) != cudaSuccess) {
            cudaFree(r.partials); cudaFree(r.ticket);
            hp_set_error("reduction scratch: allocation failed on device %d", dev);
            return 2;
        }
        it = g_red.emplace(std::make_pair(dev, st), r).first;
    }
    out = it->second;
    return 0;
}

__device__ __forceinline__ cplx hp_warp_sum(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// MODE 0: out = sum conj(x) y ; MODE 1: out = (sqrt(sum |x|^2), 0)
template <int MODE>
__global__ void __launch_bounds__(HP_RED_THREADS) hp_reduce_kernel(int64_t n, const cplx* __restrict__ x,
        const cplx* __restrict__ y, cplx* __restrict__ partials, unsigned int* ticket, cplx* __restrict__ out) {
    __shared__ cplx wsum[HP_RED_THREADS / 32];
    __shared__ bool last;
    cplx acc = cmake(0.0, 0.0);
    const int64_t stride = (int64_t)gridDim.x * HP_RED_THREADS;
    for (int64_t e = (int64_t)blockIdx.x * HP_RED_THREADS + threadIdx.x; e < n; e += stride) {
        cplx a = x[e];
        if (MODE == 0) {
            cplx b = y[e];
            acc.x = fma(a.x, b.x, fma(a.y, b.y, acc.x));
            acc.y = fma(a.x, b.y, fma(-a.y, b.x, acc.y));
        } else {
            acc.x = fma(a.x, a.x, fma(a.y, a.y, acc.x));
        }
    }
    acc = hp_warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx t = wsum[0];
        for (int w = 1; w < HP_RED_THREADS / 32; ++w) t = cadd(t, wsum[w]);
        partials[blockIdx.x] = t;
        __threadfence();
        unsigned int k = atomicAdd(ticket, 1u);
        last = (k == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        cplx t = cmake(0.0, 0.0);
        const volatile double* pv = (const volatile double*)partials;   // written by other CTAs: bypass L1
        for (int c = threadIdx.x; c < (int)gridDim.x; c += 32) t = cadd(t, cmake(pv[2 * c], pv[2 * c + 1]));
        t = hp_warp_sum(t);
        if (threadIdx.x == 0) {
            if (MODE == 1) t = cmake(sqrt(t.x), 0.0);
            *out = t;
            *ticket = 0u;
        }
    }
}

// Fused Gram-Schmidt step: w -= h v (h read from device memory) and, in the same pass, the reduction the next step needs
// from the updated w: MODE 0 out = conj(vn) . w (the next coefficient), MODE 1 out = ||w|| (after the last vector),
// MODE 2 out = sum |w|^2 (distributed vectors: the square root is taken after the all-reduce).
// Same grid, same slices and the same accumulation order as hp_reduce_kernel, so the results are bit-identical to the
// separate axpy and reduction (4 passes over n-vectors instead of 5).
template <int MODE>
__global__ void __launch_bounds__(HP_RED_THREADS) hp_axpy_reduce_kernel(int64_t n, const cplx* __restrict__ hcoef,
        const cplx* __restrict__ v, cplx* __restrict__ w, const cplx* __restrict__ vn, cplx* __restrict__ partials,
        unsigned int* ticket, cplx* __restrict__ out) {
    __shared__ cplx wsum[HP_RED_THREADS / 32];
    __shared__ bool last;
    const cplx a = cscale(-1.0, *hcoef);
    cplx acc = cmake(0.0, 0.0);
    const int64_t stride = (int64_t)gridDim.x * HP_RED_THREADS;
    for (int64_t e = (int64_t)blockIdx.x * HP_RED_THREADS + threadIdx.x; e < n; e += stride) {
        const cplx b = cfma(a, v[e], w[e]);
        w[e] = b;
        if (MODE == 0) {
            const cplx x = vn[e];
            acc.x = fma(x.x, b.x, fma(x.y, b.y, acc.x));
            acc.y = fma(x.x, b.y, fma(-x.y, b.x, acc.y));
        } else {
            acc.x = fma(b.x, b.x, fma(b.y, b.y, acc.x));
        }
    }
    acc = hp_warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx t = wsum[0];
        for (int k = 1; k < HP_RED_THREADS / 32; ++k) t = cadd(t, wsum[k]);
        partials[blockIdx.x] = t;
        __threadfence();
        unsigned int k = atomicAdd(ticket, 1u);
        last = (k == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        cplx t = cmake(0.0, 0.0);
        const volatile double* pv = (const volatile double*)partials;   // written by other CTAs: bypass L1
        for (int c = threadIdx.x; c < (int)gridDim.x; c += 32) t = cadd(t, cmake(pv[2 * c], pv[2 * c + 1]));
        t = hp_warp_sum(t);
        if (threadIdx.x == 0) {
            if (MODE == 1) t = cmake(sqrt(t.x), 0.0);
            *out = t;
            *ticket = 0u;
        }
    }
}

static unsigned hp_red_grid(int64_t n) {
    int64_t g = (n + HP_RED_THREADS * 4 - 1) / (HP_RED_THREADS * 4);
    if (g < 1) g = 1;
    if (g > HP_RED_MAX_CTAS) g = HP_RED_MAX_CTAS;
    return (unsigned)g;
}

// y += alpha x, alpha read from device memory (sign = -1 subtracts)
__global__ void __launch_bounds__(256) hp_axpy_dev_kernel(int64_t n, const cplx* __restrict__ alpha, double sign,
        const cplx* __restrict__ x, cplx* __restrict__ y) {
    cplx a = cscale(sign, *alpha);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) y[e] = cfma(a, x[e], y[e]);
}

__global__ void __launch_bounds__(256) hp_axpy_kernel(int64_t n, cplx a, const cplx* __restrict__ x,
        cplx* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) y[e] = cfma(a, x[e], y[e]);
}

__global__ void __launch_bounds__(256) hp_scale_copy_kernel(int64_t n, cplx a, const cplx* __restrict__ x,
        cplx* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) y[e] = cmul(a, x[e]);
}

#define HP_COMBINE_MAX 32
struct HpCombineArgs { cplx y[HP_COMBINE_MAX]; };
__global__ void __launch_bounds__(256) hp_combine_kernel(int64_t n, int k, const cplx* __restrict__ V, int64_t ldv,
        HpCombineArgs a, cplx* __restrict__ x) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        cplx acc = x[e];
        for (int j = 0; j < k; ++j) acc = cfma(a.y[j], V[(size_t)j * ldv + e], acc);
        x[e] = acc;
    }
}

static unsigned hp_ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

extern "C" int hp_dotc(int64_t n, const double* x, const double* y, double* out, void* stream) {
    HpRedScratch r;
    if (hp_red_scratch((cudaStream_t)stream, r)) return 2;
    hp_count_launch(); hp_reduce_kernel<0><<<hp_red_grid(n), HP_RED_THREADS, 0, (cudaStream_t)stream>>>(
        n, (const cplx*)x, (const cplx*)y, r.partials, r.ticket, (cplx*)out);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_nrm2(int64_t n, const double* x, double* out, void* stream) {
    HpRedScratch r;
    if (hp_red_scratch((cudaStream_t)stream, r)) return 2;
    hp_count_launch(); hp_reduce_kernel<1><<<hp_red_grid(n), HP_RED_THREADS, 0, (cudaStream_t)stream>>>(
        n, (const cplx*)x, (const cplx*)x, r.partials, r.ticket, (cplx*)out);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_axpy(int64_t n, double a_re, double a_im, const double* x, double* y, void* stream) {
    hp_count_launch(); hp_axpy_kernel<<<hp_ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, cmake(a_re, a_im), (const cplx*)x, (cplx*)y);
    HP_CUDA(cudaGetLastError());
    return 0;
}

// y += sign * (*alpha_dev) x with the scalar read from device memory (distributed Gram-Schmidt: no host round trip)
extern "C" int hp_axpy_dev(int64_t n, const double* alpha_dev, double sign, const double* x, double* y, void* stream) {
    hp_count_launch(); hp_axpy_dev_kernel<<<hp_ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, (const cplx*)alpha_dev, sign, (const cplx*)x, (cplx*)y);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_scale_copy(int64_t n, double a_re, double a_im, const double* x, double* y, void* stream) {
    hp_count_launch(); hp_scale_copy_kernel<<<hp_ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, cmake(a_re, a_im), (const cplx*)x, (cplx*)y);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_mgs(int64_t n, int k, const double* V, int64_t ldv, double* w, double* hcol, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    HpRedScratch r;
    if (hp_red_scratch(st, r)) return 2;
    cplx* const g_partials = r.partials;
    unsigned int* const g_ticket = r.ticket;
    cplx* h = (cplx*)hcol;
    const cplx* Vc = (const cplx*)V;
    cplx* wc = (cplx*)w;
    const unsigned g = hp_red_grid(n);
    hp_count_launch(); hp_reduce_kernel<1><<<g, HP_RED_THREADS, 0, st>>>(n, wc, wc, g_partials, g_ticket, h + k + 1);       // h0
    if (k == 0 || getenv("HP_MGS_UNFUSED")) {
        for (int j = 0; j < k; ++j) {
            hp_count_launch(); hp_reduce_kernel<0><<<g, HP_RED_THREADS, 0, st>>>(n, Vc + (size_t)j * ldv, wc, g_partials, g_ticket, h + j);
            hp_count_launch(); hp_axpy_dev_kernel<<<hp_ew_grid(n), 256, 0, st>>>(n, h + j, -1.0, Vc + (size_t)j * ldv, wc);
        }
        hp_count_launch(); hp_reduce_kernel<1><<<g, HP_RED_THREADS, 0, st>>>(n, wc, wc, g_partials, g_ticket, h + k);        // h1
    } else {
        // h_0 = v_0 . w;  then every pass subtracts h_j v_j and forms what comes next from the updated w
        hp_count_launch(); hp_reduce_kernel<0><<<g, HP_RED_THREADS, 0, st>>>(n, Vc, wc, g_partials, g_ticket, h);
        for (int j = 0; j + 1 < k; ++j) {
            hp_count_launch(); hp_axpy_reduce_kernel<0><<<g, HP_RED_THREADS, 0, st>>>(n, h + j, Vc + (size_t)j * ldv, wc, Vc + (size_t)(j + 1) * ldv,
                                                                              g_partials, g_ticket, h + j + 1);
        }
        hp_count_launch(); hp_axpy_reduce_kernel<1><<<g, HP_RED_THREADS, 0, st>>>(n, h + k - 1, Vc + (size_t)(k - 1) * ldv, wc, wc, g_partials, g_ticket,
                                                                          h + k);                                     // h1
    }
    HP_CUDA(cudaGetLastError());
    return 0;
}

// one fused Gram-Schmidt step for distributed vectors (the coefficient is all-reduced between the steps by the caller):
//   w -= (*hcoef_dev) v ;  out_dev = vdot(vnext, w) over the local entries (vnext != NULL)  or  sum |w|^2 (vnext == NULL)
extern "C" int hp_mgs_step(int64_t n, const double* hcoef_dev, const double* v, double* w, const double* vnext, double* out_dev,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    HpRedScratch r;
    if (hp_red_scratch(st, r)) return 2;
    const unsigned g = hp_red_grid(n);
    hp_count_launch();
    if (vnext) hp_axpy_reduce_kernel<0><<<g, HP_RED_THREADS, 0, st>>>(n, (const cplx*)hcoef_dev, (const cplx*)v, (cplx*)w, (const cplx*)vnext,
                                                                     r.partials, r.ticket, (cplx*)out_dev);
    else hp_axpy_reduce_kernel<2><<<g, HP_RED_THREADS, 0, st>>>(n, (const cplx*)hcoef_dev, (const cplx*)v, (cplx*)w, (const cplx*)w,
                                                              r.partials, r.ticket, (cplx*)out_dev);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_combine(int64_t n, int k, const double* V, int64_t ldv, const double* y_host, double* x,
                          void* stream) {
    const cplx* Vc = (const cplx*)V;
    for (int j0 = 0; j0 < k; j0 += HP_COMBINE_MAX) {
        int kk = k - j0 < HP_COMBINE_MAX ? k - j0 : HP_COMBINE_MAX;
        HpCombineArgs a;
        for (int j = 0; j < kk; ++j) a.y[j] = cmake(y_host[2 * (j0 + j)], y_host[2 * (j0 + j) + 1]);
        hp_count_launch(); hp_combine_kernel<<<hp_ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, kk, Vc + (size_t)j0 * ldv, ldv, a,
                                                                          (cplx*)x);
    }
    HP_CUDA(cudaGetLastError());
    return 0;
}
