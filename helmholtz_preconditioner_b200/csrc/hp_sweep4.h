// Cluster sweep kernel (csrc/hp_sweep4.cu): shared-memory plan and host entry points.
#pragma once
#include "hp_internal.cuh"

#define HP4_PL 3          // separators per lane in the gather of x: P-1 <= 32*HP4_PL
#define HP4_PW 2          // words per lane and round of the poll warp (K*b <= 64 in one round)
#define HP4_EW 3          // gathered entries per warp and batch (one batch when ceil(3b/K) <= 4*HP4_EW)

struct Hp4Plan {
    int RC, NCH, S;                 // rows per W chunk (power of two), chunks per strip, slots of the W ring
    size_t w_st, g_st, n_st, r_st;  // stage strides in bytes (multiples of 128)
    size_t total;                   // dynamic shared memory per CTA
    size_t log_off;                 // developer instantiation: byte offset of the event log (0 = none), see HP_STAMP4
    int win0;                       // first strip (iteration) of the logged window
};
#define HP4_LOG_STRIPS 32
#define HP4_LOG_BYTES (HP4_LOG_STRIPS * 16 * 4)

struct HpSweepArgs;
int hp_sweep4_plan(const HpLayout& L, int b, size_t max_smem, Hp4Plan& pl, int RT = 1);   // RT right-hand sides per launch
int hp_sweep4_max_clusters(const HpLayout& L, int b);
int hp_sweep4_launch(hp_solver* s, HpSweepArgs& a, cudaStream_t st);
bool hp_profiler_attached();
// csrc/hp_sweep4m.cu: RT right-hand sides per launch
int hp_sweep4m_supported(hp_solver* s, int RT);      // 0 = this layout can run RT right-hand sides per launch
int hp_sweep4m_launch(hp_solver* s, HpSweepArgs& a, int RT, cudaStream_t st);
// csrc/hp_sweep4d.cu: 8 right-hand sides per launch on the FP64 tensor cores
int hp_sweep4d_supported(hp_solver* s);
int hp_sweep4d_launch(hp_solver* s, HpSweepArgs& a, cudaStream_t st);
