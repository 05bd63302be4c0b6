// Front block of the sweeping preconditioner: the first b grid rows (the fixed PML of the operator).
//
// Reference: get_A_FF_block (/root/reference/code.py:178-183) keeps only the diagonal blocks A_11..A_bb,
// so H_F is b independent complex tridiagonal systems of size n (one per grid row); the reference factors
// it with SuperLU in algo2_3 (code.py:346-347) and solves with it twice per application in algo2_4
// (code.py:364-365 and 381-384).  Here: Thomas factors computed once, one thread per grid row.
#include "hp_internal.cuh"

// factors stored [i][j] (x1 index major) so that the b threads of the solve read contiguous memory
__global__ void hp_front_factor_kernel(int n, int b, double ih2, cplx omega2, const cplx* __restrict__ s1t,
                                       const cplx* __restrict__ is1t, const cplx* __restrict__ s2t,
                                       const cplx* __restrict__ is2t, const double* __restrict__ kappa,
                                       cplx* __restrict__ low, cplx* __restrict__ invd, cplx* __restrict__ up,
                                       int* status) {
    int j0 = threadIdx.x;            // 0-based grid row
    if (j0 >= b) return;
    int j = j0 + 1;
    cplx is2c = is2t[2 * j];
    cplx g3 = cscale(ih2, s2t[2 * j - 1]), g4 = cscale(ih2, s2t[2 * j + 1]);
    cplx dprev = cmake(1.0, 0.0), cprev = cmake(0.0, 0.0);
    for (int i = 1; i <= n; ++i) {
        cplx is1c = is1t[2 * i];
        cplx c1 = cscale(ih2, cmul(s1t[2 * i - 1], is2c));
        cplx c2 = cscale(ih2, cmul(s1t[2 * i + 1], is2c));
        cplx c3 = cmul(g3, is1c), c4 = cmul(g4, is1c);
        cplx c5 = cscale(kappa[(size_t)j0 * n + (i - 1)], cmul(omega2, cmul(is1c, is2c)));
        c5 = csub(c5, cadd(cadd(c1, c2), cadd(c3, c4)));
        cplx w = cmake(0.0, 0.0), d = c5;
        if (i > 1) {
            w = cdiv(c1, dprev);
            d = cfms(w, cprev, c5);
        }
        if (d.x == 0.0 && d.y == 0.0) atomicOr(status, 8);
        size_t o = (size_t)(i - 1) * b + j0;
        low[o] = w;
        invd[o] = cinv(d);
        up[o] = c2;
        dprev = d;
        cprev = c2;
    }
}

// mode 0: out[j][.] = Tri_j^{-1} in[j][.] for the rows j = row0 .. row0 + nrows - 1  (thread per row)
// mode 1: single row j = b-1 with rhs = upc * u_next,  out[.] = base[.] - solution      (code.py:381-384)
// The recurrences are sequential in i; their operands are not, so they are fetched HP_FRONT_U steps ahead
// into registers (and a few hundred steps ahead into L2) and only the complex multiply-add chain is exposed.
#define HP_FRONT_U 16
__global__ void hp_front_solve_kernel(int n, int b, int row0, int nrows, int mode, const cplx* __restrict__ low,
                                      const cplx* __restrict__ invd, const cplx* __restrict__ up,
                                      const cplx* in, cplx* out, const cplx* base, cplx upfac,
                                      const cplx* __restrict__ is1t, cplx* work) {
    int t = threadIdx.x;
    if (t >= nrows) return;
    int j0 = row0 + t;
    const cplx* r = in + (size_t)t * n;
    cplx* y = work + (size_t)t * n;
    cplx prev = cmake(0.0, 0.0);
    for (int i0 = 0; i0 < n; i0 += HP_FRONT_U) {
        cplx rr[HP_FRONT_U], ll[HP_FRONT_U];
        if (i0 + 256 < n) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(r + i0 + 256));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(low + (size_t)(i0 + 256) * b + j0));
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k) {
            int i = i0 + k;
            if (i < n) {
                rr[k] = r[i];
                ll[k] = low[(size_t)i * b + j0];
                if (mode == 1) rr[k] = cmul(cmul(upfac, is1t[2 * (i + 1)]), rr[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k) {
            int i = i0 + k;
            if (i < n) { prev = cfms(ll[k], prev, rr[k]); rr[k] = prev; }
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k)
            if (i0 + k < n) y[i0 + k] = rr[k];
    }
    cplx xn = cmake(0.0, 0.0);
    cplx* o = out + (size_t)t * n;
    for (int i1 = n - 1; i1 >= 0; i1 -= HP_FRONT_U) {
        cplx yy[HP_FRONT_U], uu[HP_FRONT_U], dd[HP_FRONT_U], bb[HP_FRONT_U];
        if (i1 - 256 >= 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(up + (size_t)(i1 - 256) * b + j0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(invd + (size_t)(i1 - 256) * b + j0));
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k) {
            int i = i1 - k;
            if (i >= 0) {
                size_t f = (size_t)i * b + j0;
                yy[k] = y[i]; uu[k] = up[f]; dd[k] = invd[f];
                if (mode == 1) bb[k] = base[i];
            }
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k) {
            int i = i1 - k;
            if (i >= 0) { xn = cmul(cfms(uu[k], xn, yy[k]), dd[k]); yy[k] = mode == 1 ? csub(bb[k], xn) : xn; }
        }
#pragma unroll
        for (int k = 0; k < HP_FRONT_U; ++k)
            if (i1 - k >= 0) o[i1 - k] = yy[k];
    }
}

// u_row[c] -= fac * is1t[2(c+1)] * src[c]
__global__ void hp_row_couple_kernel(int n, cplx fac, const cplx* __restrict__ is1t, const cplx* __restrict__ src,
                                     cplx* __restrict__ row) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) row[c] = cfms(cmul(fac, is1t[2 * (c + 1)]), src[c], row[c]);
}

int hp_front_setup(hp_solver* s, cudaStream_t st) {
    size_t sz = sizeof(cplx) * (size_t)s->b * s->n;
    if (!s->f_low) {
        HP_CUDA(cudaMalloc(&s->f_low, sz));
        HP_CUDA(cudaMalloc(&s->f_invd, sz));
        HP_CUDA(cudaMalloc(&s->f_up, sz));
        HP_CUDA(cudaMalloc(&s->TF, sz + sizeof(cplx) * (size_t)s->b * s->n));   // TF followed by the work rows
    }
    double ih2 = 1.0 / (s->pml.h * s->pml.h);
    hp_count_launch(); hp_front_factor_kernel<<<1, 32, 0, st>>>(s->n, s->b, ih2, s->omega2, s->s1t, s->is1t, s->s2t, s->is2t, s->kappa,
                                             s->f_low, s->f_invd, s->f_up, s->status);
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_front_begin(hp_solver* s, double* u_dev, void* stream) {
    if (!s || !s->f_low) { hp_set_error("hp_front_begin: preconditioner not set up"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->n, b = s->b;
    cplx* u = (cplx*)u_dev;
    cplx* work = s->TF + (size_t)b * n;
    hp_count_launch(); hp_front_solve_kernel<<<1, 32, 0, st>>>(n, b, 0, b, 0, s->f_low, s->f_invd, s->f_up, u, s->TF, nullptr,
                                            cmake(0, 0), s->is1t, work);
    if (b < n) {
        // u_{b+1} -= A_{b+1,b} (T_F u_F)_b : A_{b+1,b} = diag(c3) of grid row b+1 (code.py:145-154, :365)
        double ih2 = 1.0 / (s->pml.h * s->pml.h);
        cplx fac = cscale(ih2, s->s2t_h[2 * (b + 1) - 1]);
        hp_count_launch(); hp_row_couple_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, fac, s->is1t, s->TF + (size_t)(b - 1) * n,
                                                              u + (size_t)b * n);
    }
    HP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int hp_front_end(hp_solver* s, double* u_dev, void* stream) {
    if (!s || !s->f_low) { hp_set_error("hp_front_end: preconditioner not set up"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->n, b = s->b;
    cplx* u = (cplx*)u_dev;
    cplx* work = s->TF + (size_t)b * n;
    if (b > 1) HP_CUDA(cudaMemcpyAsync(u, s->TF, sizeof(cplx) * (size_t)(b - 1) * n, cudaMemcpyDeviceToDevice, st));
    if (b < n) {
        // u_b = (T_F u_F)_b - Tri_b^{-1} (A_{b,b+1} u_{b+1}) : A_{b,b+1} = diag(c4) of grid row b (code.py:131-140)
        double ih2 = 1.0 / (s->pml.h * s->pml.h);
        cplx fac = cscale(ih2, s->s2t_h[2 * b + 1]);
        hp_count_launch(); hp_front_solve_kernel<<<1, 32, 0, st>>>(n, b, b - 1, 1, 1, s->f_low, s->f_invd, s->f_up, u + (size_t)b * n,
                                                u + (size_t)(b - 1) * n, s->TF + (size_t)(b - 1) * n, fac, s->is1t,
                                                work);
    } else {
        HP_CUDA(cudaMemcpyAsync(u + (size_t)(b - 1) * n, s->TF + (size_t)(b - 1) * n, sizeof(cplx) * n,
                                cudaMemcpyDeviceToDevice, st));
    }
    HP_CUDA(cudaGetLastError());
    return 0;
}

// T_F u_F lives in the solver between hp_front_begin and hp_front_end; with several right-hand sides in flight
// (slab.py pipelines them through the slabs) the caller parks it in its own buffer of b*n complex numbers.
// dir 0: solver -> buf_dev, dir 1: buf_dev -> solver.
extern "C" int hp_front_tf_copy(hp_solver* s, double* buf_dev, int dir, void* stream) {
    if (!s || !s->TF) { hp_set_error("hp_front_tf_copy: preconditioner not set up"); return 1; }
    size_t sz = sizeof(cplx) * (size_t)s->b * s->n;
    HP_CUDA(cudaMemcpyAsync(dir == 0 ? (void*)buf_dev : (void*)s->TF, dir == 0 ? (const void*)s->TF : (const void*)buf_dev, sz,
                            cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
