// Block Gram-Schmidt passes for slab-distributed Krylov vectors: w against ALL k basis vectors in one pass, for up to 8
// systems in one launch.
//
// scipy's gmres orthogonalises with modified Gram-Schmidt (scipy/sparse/linalg/_isolve/iterative.py, called from
// /root/reference/code.py:516): k dependent steps per Arnoldi column, and on distributed vectors every step ends in an
// all-reduce.  With several groups of right-hand sides in flight (slab.GroupPipeline) each of those k + 2 all-reduces waits
// for the slowest rank, whose vector kernels queue behind the slab sweep of another group (up to one slab sweep, 4 ms at
// N = 8): 12 global synchronisations per column cost more than the sweeps of the column.  Classical Gram-Schmidt applied
// twice ("twice is enough") needs THREE per column whatever k is and is at least as orthogonal as MGS:
//     pass A   d_j = v_j^H w (j < k),  |w|^2                                     all-reduce
//     pass B   w <- w - sum_j d_j v_j,  e_j = v_j^H w,                           all-reduce
//     pass C   w <- w - sum_j e_j v_j,  |w|^2                                    all-reduce
//     h_j = d_j + e_j
// One kernel: a thread keeps the k basis entries of its element in registers, applies the update (pass B, C) and
// accumulates the k dot products and the norm from the updated value: every vector is read once per pass (k + 1 reads, one
// write).  Reductions are single-launch and deterministic (fixed slices, last CTA sums the partials in index order).
#include "hp_internal.cuh"

#include <map>
#include <mutex>

#define HP_CGS_THREADS 256
#define HP_CGS_MAX_CTAS 592     // per system: 4 per SM on a 148-SM part
#define HP_CGS_RMAX 8

struct HpCgsScratch { cplx* partials = nullptr; unsigned int* tickets = nullptr; };
static std::map<std::pair<int, cudaStream_t>, HpCgsScratch> g_cgs;
static std::mutex g_cgs_mu;

static int hp_cgs_scratch(cudaStream_t st, HpCgsScratch& out) {
    int dev = 0;
    HP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_cgs_mu);
    auto it = g_cgs.find({dev, st});
    if (it == g_cgs.end()) {
        HpCgsScratch r;
        size_t pb = sizeof(cplx) * (size_t)HP_CGS_RMAX * HP_CGS_MAX_CTAS * 21;
        if (cudaMalloc(&r.partials, pb) != cudaSuccess || cudaMalloc(&r.tickets, sizeof(unsigned int) * HP_CGS_RMAX) != cudaSuccess ||
            cudaMemset(r.tickets, 0, sizeof(unsigned int) * HP_CGS_RMAX) != cudaSuccess) {
            cudaFree(r.partials); cudaFree(r.tickets);
            hp_set_error("Gram-Schmidt scratch: allocation failed on device %d", dev);
            return 2;
        }
        it = g_cgs.emplace(std::make_pair(dev, st), r).first;
    }
    out = it->second;
    return 0;
}

struct HpCgsArgs {
    const cplx* V[HP_CGS_RMAX];      // basis of system y: rows V + j * ldv
    cplx* w[HP_CGS_RMAX];
    const cplx* coef[HP_CGS_RMAX];   // k coefficients of the update (device memory)
    cplx* out[HP_CGS_RMAX];          // k dot products (DOTS) and, at out[k], sum |w|^2
};

__device__ __forceinline__ cplx cgs_warp_sum(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// KT: compile-time bound of k (the basis entries and the accumulators live in registers)
template <int KT, bool UPD, bool DOTS>
__global__ void __launch_bounds__(HP_CGS_THREADS) hp_cgs_kernel(HpCgsArgs a, int64_t n, int k, int64_t ldv, cplx* __restrict__ partials,
                                                               unsigned int* tickets) {
    constexpr int NA = DOTS ? KT + 1 : 1;          // accumulators: the dots, then the norm
    __shared__ cplx wsum[HP_CGS_THREADS / 32][KT + 1];
    __shared__ cplx cs[KT];
    __shared__ bool last;
    const int y = blockIdx.y;
    const cplx* __restrict__ V = a.V[y];
    cplx* __restrict__ w = a.w[y];
    if (UPD && threadIdx.x < KT) cs[threadIdx.x] = (int)threadIdx.x < k ? cscale(-1.0, a.coef[y][threadIdx.x]) : cmake(0.0, 0.0);
    __syncthreads();
    cplx acc[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) acc[j] = cmake(0.0, 0.0);
    const int64_t stride = (int64_t)gridDim.x * HP_CGS_THREADS;
    for (int64_t e = (int64_t)blockIdx.x * HP_CGS_THREADS + threadIdx.x; e < n; e += stride) {
        cplx vj[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) vj[j] = j < k ? V[(size_t)j * ldv + e] : cmake(0.0, 0.0);
        cplx wv = w[e];
        if (UPD) {
#pragma unroll
            for (int j = 0; j < KT; ++j) wv = cfma(cs[j], vj[j], wv);
            w[e] = wv;
        }
        if (DOTS) {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                acc[j].x = fma(vj[j].x, wv.x, fma(vj[j].y, wv.y, acc[j].x));
                acc[j].y = fma(vj[j].x, wv.y, fma(-vj[j].y, wv.x, acc[j].y));
            }
        }
        acc[NA - 1].x = fma(wv.x, wv.x, fma(wv.y, wv.y, acc[NA - 1].x));
    }
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        cplx t = cgs_warp_sum(acc[j]);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][j] = t;
    }
    __syncthreads();
    cplx* mine = partials + ((size_t)y * HP_CGS_MAX_CTAS + blockIdx.x) * 21;
    if (threadIdx.x < NA) {
        cplx t = wsum[0][threadIdx.x];
        for (int q = 1; q < HP_CGS_THREADS / 32; ++q) t = cadd(t, wsum[q][threadIdx.x]);
        mine[threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tk = atomicAdd(tickets + y, 1u);
        last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        const volatile double* pv = (const volatile double*)(partials + (size_t)y * HP_CGS_MAX_CTAS * 21);   // other CTAs wrote these
        // warp q sums accumulator j = q, q + 8, ...: fixed order over the CTAs
        const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
        for (int j = wq; j < NA; j += HP_CGS_THREADS / 32) {
            cplx t = cmake(0.0, 0.0);
            for (int c = lane; c < (int)gridDim.x; c += 32) t = cadd(t, cmake(pv[2 * ((size_t)c * 21 + j)], pv[2 * ((size_t)c * 21 + j) + 1]));
            t = cgs_warp_sum(t);
            if (lane == 0) {
                if (DOTS) { if (j < KT) { if (j < k) a.out[y][j] = t; } else a.out[y][k] = t; }
                else a.out[y][k] = t;
            }
        }
        if (threadIdx.x == 0) tickets[y] = 0u;
    }
}

template <int KT>
static void hp_cgs_launch(const HpCgsArgs& a, int R, int64_t n, int k, int64_t ldv, int update, int dots, const HpCgsScratch& sc, cudaStream_t st) {
    int64_t g = (n + HP_CGS_THREADS * 4 - 1) / (HP_CGS_THREADS * 4);
    if (g < 1) g = 1;
    if (g > HP_CGS_MAX_CTAS) g = HP_CGS_MAX_CTAS;
    dim3 grid((unsigned)g, (unsigned)R);
    if (update && dots) hp_cgs_kernel<KT, true, true><<<grid, HP_CGS_THREADS, 0, st>>>(a, n, k, ldv, sc.partials, sc.tickets);
    else if (update) hp_cgs_kernel<KT, true, false><<<grid, HP_CGS_THREADS, 0, st>>>(a, n, k, ldv, sc.partials, sc.tickets);
    else hp_cgs_kernel<KT, false, true><<<grid, HP_CGS_THREADS, 0, st>>>(a, n, k, ldv, sc.partials, sc.tickets);
}

// One pass for R <= 8 systems of n local entries, k <= 20 basis vectors each (rows of V_devs[r], leading dimension ldv):
//   update != 0:  w_r <- w_r - sum_j coef_r[j] v_j        (coef_devs[r]: k complex numbers in device memory)
//   dots   != 0:  out_r[j] = v_j^H w_r (j < k) from the updated w_r
//   always:       out_r[k] = sum |w_r|^2 from the updated w_r
extern "C" int hp_cgs_pass(int R, int64_t n, int k, const double* const* V_devs, int64_t ldv, double* const* w_devs,
                           const double* const* coef_devs, double* const* out_devs, int update, int dots, void* stream) {
    if (R < 1 || R > HP_CGS_RMAX || k < 0 || k > 20 || !V_devs || !w_devs || !out_devs || (update && !coef_devs) || (!update && !dots)) {
        hp_set_error("hp_cgs_pass: need 1 <= R <= 8, 0 <= k <= 20, update or dots; got R=%d k=%d", R, k);
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HpCgsScratch sc;
    if (hp_cgs_scratch(st, sc)) return 2;
    HpCgsArgs a = {};
    for (int r = 0; r < R; ++r) {
        a.V[r] = (const cplx*)V_devs[r]; a.w[r] = (cplx*)w_devs[r];
        a.coef[r] = update ? (const cplx*)coef_devs[r] : nullptr; a.out[r] = (cplx*)out_devs[r];
    }
    hp_count_launch();
    if (k <= 4) hp_cgs_launch<4>(a, R, n, k, ldv, update, dots, sc, st);
    else if (k <= 8) hp_cgs_launch<8>(a, R, n, k, ldv, update, dots, sc, st);
    else if (k <= 12) hp_cgs_launch<12>(a, R, n, k, ldv, update, dots, sc, st);
    else if (k <= 16) hp_cgs_launch<16>(a, R, n, k, ldv, update, dots, sc, st);
    else hp_cgs_launch<20>(a, R, n, k, ldv, update, dots, sc, st);
    HP_CUDA(cudaGetLastError());
    return 0;
}
