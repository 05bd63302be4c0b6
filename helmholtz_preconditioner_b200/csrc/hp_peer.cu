// Peer mailboxes: the hand-over of field rows between the slabs of consecutive GPUs (slab.py) through peer memory over
// NVLink, with waits that occupy no SM.
//
// Why not NCCL send/recv here: a receive (or a collective) that waits for a sweep on another GPU is a kernel that spins on
// an SM.  The cluster sweep kernels need every cluster slot of the device (33 clusters of 4 CTAs at 4096^2, each CTA a
// whole SM), so a communication CTA that sits on one of those SMs keeps a sweep of another group of right-hand sides
// partially resident; with several groups in flight (slab.GroupPipeline) the partially resident sweeps and the waiting
// communication kernels of two GPUs wait for each other.  Here the sender copies the rows into a staging area of the
// receiver (one small kernel, stores over NVLink) and releases a sequence number next to them; the receiver's stream
// waits for that number with a stream memory operation (cuStreamWaitValue32), which holds no SM.
//
// A mailbox is one cudaMalloc'ed block, exported with cudaIpcGetMemHandle and mapped by the neighbours:
//     [0, 256)            flags (uint32): 0 forward rows, 1 backward rows, 2 application finished (from rank 0)
//     [256, ...)          forward staging [R][n] complex, then backward staging [R][n] complex
#include <cuda.h>
#include <string.h>

#include "hp_internal.cuh"

typedef CUresult (*hp_wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static hp_wait32_fn hp_wait32() {
    static hp_wait32_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (hp_wait32_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

extern "C" int hp_mailbox_create(int64_t bytes, void** ptr_dev, unsigned char* handle64) {
    if (!ptr_dev || !handle64 || bytes < 256) { hp_set_error("hp_mailbox_create: bad argument"); return 1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    HP_CUDA(cudaMalloc(&p, (size_t)bytes));
    HP_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); hp_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return 2; }
    memcpy(handle64, &h, 64);
    *ptr_dev = p;
    return 0;
}

extern "C" int hp_mailbox_open(const unsigned char* handle64, void** ptr_dev) {
    if (!ptr_dev || !handle64) { hp_set_error("hp_mailbox_open: bad argument"); return 1; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    HP_CUDA(cudaIpcOpenMemHandle(ptr_dev, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int hp_mailbox_close(void* ptr_dev) {
    if (ptr_dev) HP_CUDA(cudaIpcCloseMemHandle(ptr_dev));
    return 0;
}

extern "C" int hp_mailbox_free(void* ptr_dev) {
    if (ptr_dev) HP_CUDA(cudaFree(ptr_dev));
    return 0;
}

// the stream waits until *flag_dev >= value (flag in this device's memory); no SM is held while it waits
extern "C" int hp_stream_wait_geq(const void* flag_dev, unsigned int value, void* stream) {
    hp_wait32_fn fn = hp_wait32();
    if (!fn) { hp_set_error("hp_stream_wait_geq: cuStreamWaitValue32 is not available from this driver"); return 2; }
    CUresult r = fn((CUstream)stream, (CUdeviceptr)flag_dev, (cuuint32_t)value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) { hp_set_error("cuStreamWaitValue32 failed (%d)", (int)r); return 2; }
    return 0;
}

struct HpRowPtrs { const cplx* src[8]; cplx* dst[8]; };

// rows -> staging of the neighbour, then the sequence number (release at system scope after every thread's stores)
__global__ void __launch_bounds__(1024) hp_handover_kernel(HpRowPtrs p, int R, int n, cplx* __restrict__ staging, unsigned int* flag,
                                                           unsigned int value) {
    for (int r = 0; r < R; ++r) {
        const double2* s = reinterpret_cast<const double2*>(p.src[r]);
        double2* d = reinterpret_cast<double2*>(staging + (size_t)r * n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
    }
}

// staging (this device) -> the rows of the local buffers
__global__ void __launch_bounds__(1024) hp_collect_kernel(HpRowPtrs p, int R, int n, const cplx* staging) {
    const int r = blockIdx.x;
    const volatile double* s = reinterpret_cast<const volatile double*>(staging + (size_t)r * n);
    double* d = reinterpret_cast<double*>(p.dst[r]);
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) d[i] = s[i];
}

__global__ void hp_signal_kernel(HpRowPtrs p, int count, unsigned int value) {
    if ((int)threadIdx.x < count) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"((unsigned int*)p.dst[threadIdx.x]), "r"(value) : "memory");
    }
}

// R <= 8 rows of n complex numbers each, src_rows[r] on this device -> staging_peer + r * n, then *flag_peer = value
extern "C" int hp_handover_rows(int R, const double* const* src_rows, double* staging_peer, int64_t n, void* flag_peer,
                                unsigned int value, void* stream) {
    if (R < 1 || R > 8 || !src_rows || !staging_peer || !flag_peer) { hp_set_error("hp_handover_rows: bad argument"); return 1; }
    HpRowPtrs p = {};
    for (int r = 0; r < R; ++r) p.src[r] = (const cplx*)src_rows[r];
    hp_count_launch();
    hp_handover_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p, R, (int)n, (cplx*)staging_peer, (unsigned int*)flag_peer, value);
    HP_CUDA(cudaGetLastError());
    return 0;
}

// staging + r * n (this device) -> dst_rows[r]; call after hp_stream_wait_geq on the flag that guards the staging area
extern "C" int hp_collect_rows(int R, const double* staging, double* const* dst_rows, int64_t n, void* stream) {
    if (R < 1 || R > 8 || !dst_rows || !staging) { hp_set_error("hp_collect_rows: bad argument"); return 1; }
    HpRowPtrs p = {};
    for (int r = 0; r < R; ++r) p.dst[r] = (cplx*)dst_rows[r];
    hp_count_launch();
    hp_collect_kernel<<<R, 1024, 0, (cudaStream_t)stream>>>(p, R, (int)n, (const cplx*)staging);
    HP_CUDA(cudaGetLastError());
    return 0;
}

// *flags_peer[i] = value for count <= 8 flags (rank 0 tells the other ranks that an application of algo2_4 is complete)
extern "C" int hp_signal_flags(int count, void* const* flags_peer, unsigned int value, void* stream) {
    if (count < 1 || count > 8 || !flags_peer) { hp_set_error("hp_signal_flags: bad argument"); return 1; }
    HpRowPtrs p = {};
    for (int i = 0; i < count; ++i) p.dst[i] = (cplx*)flags_peer[i];
    hp_count_launch();
    hp_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, count, value);
    HP_CUDA(cudaGetLastError());
    return 0;
}
