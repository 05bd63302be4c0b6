// Internal declarations shared by the translation units of libhelmholtz_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "hp_setup_core.h"

#define HP_QMAX 64          // widest leaf (block rows) of the x1 partition
#define HP_MAX_LEVELS 12    // up to 4096 leaves

void hp_set_error(const char* fmt, ...);

#define HP_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            hp_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

struct hp_solver {
    int n = 0, b = 0;
    HpPml pml;
    cplx omega2;
    int num_sms = 0;
    // operator tables on the half grid t = 0..2n+2 (unshifted PML)
    cplx *s1t = nullptr, *is1t = nullptr, *s2t = nullptr, *is2t = nullptr;
    double* c_mat = nullptr;      // (n+2)^2 as given
    double* kappa = nullptr;      // n*n, kappa[(j-1)*n + (i-1)] = 1 / c_mat[i-1][j-1]^2  (grid-aligned)
    // strip factorisation
    int d = 0, P = 0, QP = 0;     // tree depth, leaves, padded leaf width
    int m_lo = 0, m_hi = -1;      // strips held by this solver
    std::vector<int> leaf_start_h;
    int* leaf_start = nullptr;    // device, P+1 (0-based first block row of each leaf)
    int lvoff[HP_MAX_LEVELS + 2];
    cplx *W = nullptr, *G = nullptr, *nodes = nullptr;
    size_t W_stride = 0, G_stride = 0, N_stride = 0;   // complex entries per strip
    int64_t bytes = 0;
    // front block: Thomas factors of the b tridiagonal diagonal blocks (reference H_F)
    cplx *fw = nullptr, *finvd = nullptr, *fup = nullptr;   // [b][n]
    cplx* TF = nullptr;                                      // [b][n]   T_F u_F kept between the stages
    cplx* ztmp = nullptr;                                    // [2][n]
    // sweep scratch (tree exchange)
    cplx *seg = nullptr, *xi = nullptr, *ext = nullptr;
    int sweep_ctas = 0, sweep_lpc = 0;
    int* status = nullptr;        // device flag: non-zero when a pivot vanished during setup
};

struct HpCtxDev {   // by-value kernel argument
    HpStripCtx c;
};

// hp_assembly.cu
int hp_launch_tables(hp_solver* s, cudaStream_t st);
// hp_setup.cu
int hp_setup_strips(hp_solver* s, int qmax, int m_lo, int m_hi, cudaStream_t st);
void hp_free_strips(hp_solver* s);
// hp_front.cu
int hp_front_setup(hp_solver* s, cudaStream_t st);
// hp_sweep.cu
int hp_sweep_launch(hp_solver* s, int mode, cplx* u, const cplx* vin, cplx* yout, int m_from, int m_to, cudaStream_t st);
