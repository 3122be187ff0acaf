// common.cuh -- shared state, error handling and small device helpers for libcenn.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../include/cenn.h"

// One-shot all-reduce over NVLink peer memory (dist.cu / nhwc.cuh xr_sum_inplace): every rank owns a mailbox
// [2 parities][world source rows][XR_MAXF] 8-byte {value, exchange tag} words; peers' mailboxes are mapped with CUDA IPC.
static const int XR_MAXF = 16384;       // floats per exchange (2 x 8192 statistics columns)
static const int XR_MAX_WORLD = 16;
struct XrCtx {
    float *data[XR_MAX_WORLD];                  // data[r]: rank r's mailbox payload (data[rank] is local memory)
    unsigned long long *flags[XR_MAX_WORLD];    // flags[r][parity]: last epoch rank r has published
    unsigned long long *epoch;                  // this rank's exchange counter (device memory)
    int world, rank;
    int push;                                   // 1: store into the peers' mailboxes and poll local memory (default); 0: publish locally, poll the peers
    long long timeout_cycles;                   // a peer that has not published after this many SM cycles traps the kernel (CENN_XR_TIMEOUT_S, default 120 s:
                                                // long enough for a rank that writes a checkpoint or stalls in its loader, short enough not to hang a box forever)
};

struct cenn_state {
    int device = 0;
    int precision = CENN_FP32;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
    // scratch: small reduction accumulators (doubles) + a growable workspace
    double *red = nullptr;            // [RED_SLOTS] device doubles
    double *red_host = nullptr;       // pinned mirror
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *ws2 = nullptr;
    size_t ws2_bytes = 0;
    // data parallelism (dist.cu): NCCL communicator of this process's GPU
    void *comm = nullptr;                 // latency-critical small reductions, on the compute stream
    void *comm2 = nullptr;                // bulk gradient buckets, on comm_stream
    int world = 1, rank = 0;
    cudaStream_t comm_stream = nullptr;   // gradient all-reduces overlapped with the backward sweep
    bool xr_enabled = false;              // peer mailboxes mapped: BN statistics are exchanged inside the finalize kernels
    XrCtx xr = {};                        // exchanges issued from the compute stream
    XrCtx xr2 = {};                       // a second, independent mailbox sequence for a concurrent chain (trainer side stream)
    XrCtx xr3 = {};                       // a third one for the communication stream (barriers around the sharded reduce + Adam kernel)
    void *xr_own = nullptr;
};
static const int RED_SLOTS = 64;

void cenn_set_error(const char *fmt, ...);
int cenn_check_cuda(cudaError_t e, const char *what, const char *file, int line);
void *cenn_workspace(cenn_state *s, size_t bytes);   // stream-ordered reuse; grows with cudaMalloc
void *cenn_workspace2(cenn_state *s, size_t bytes);
int cenn_dist_all_reduce_on(cenn_state *s, void *buf, int64_t count, int is_double, cudaStream_t stream);
int cenn_dist_all_reduce_bulk(cenn_state *s, float *buf, int64_t count);
int cenn_dist_all_reduce_bulk_bf16(cenn_state *s, void *buf, int64_t count);
int cenn_dist_ipc_map(cenn_state *s, void *base, void **peers /*[XR_MAX_WORLD]*/);   // collective: peers[r] = rank r's allocation (CUDA IPC)
void cenn_dist_ipc_unmap(cenn_state *s, void **peers);
int cenn_dist_barrier(cenn_state *s);
int cenn_dist_group(int begin);       // ncclGroupStart / ncclGroupEnd   // second communicator, s->comm_stream

#define CK(expr) do { if (cenn_check_cuda((expr), #expr, __FILE__, __LINE__)) return 1; } while (0)
#define CK_LAUNCH(s) do { (s)->launches++; if (cenn_check_cuda(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)) return 1; } while (0)
#define REQUIRE(cond, ...) do { if (!(cond)) { cenn_set_error(__VA_ARGS__); return 1; } } while (0)
#define API_BEGIN(s) do { if (!(s)) { cenn_set_error("null cenn_state"); return 1; } \
    if (cenn_check_cuda(cudaSetDevice((s)->device), "cudaSetDevice", __FILE__, __LINE__)) return 1; } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for bandwidth kernels: a multiple of the SM count (148 on B200), capped by the work
static inline int bw_grid(const cenn_state *s, int64_t work_items, int threads, int per_sm = 8) {
    int64_t need = ceil_div64(work_items, threads);
    int64_t cap = (int64_t)s->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream is still draining -- its prologue (barrier init, TMEM allocation, descriptor prefetch, launch
// latency) overlaps the predecessor's tail.  pdl_wait() blocks until the predecessor grid has COMPLETED and its memory is visible:
// every kernel launched through LK() executes it before its first dependent global access.  pdl_trigger() lets the successor be
// scheduled as soon as every CTA of this grid has started.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// `LK(kernel, grid, block, smem, stream)(args...)` = `kernel<<<grid, block, smem, stream>>>(args...)`, plus the PDL attribute when
// enabled (CENN_PDL=0 disables).  Only kernels that call pdl_wait() may be launched this way.
inline bool cenn_pdl_enabled() { static const bool on = !(getenv("CENN_PDL") && atoi(getenv("CENN_PDL")) == 0); return on; }
template <typename... KArgs>
struct CennLauncher {
    void (*kern)(KArgs...); dim3 grid, block; size_t smem; cudaStream_t stream;
    template <typename... A> void operator()(A &&...a) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = cenn_pdl_enabled() ? 1 : 0;
        cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(a)...);          // errors surface through cudaGetLastError (KLAUNCH / CK_LAUNCH)
    }
};
template <typename... KArgs>
inline CennLauncher<KArgs...> LK(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream) { return CennLauncher<KArgs...>{kern, grid, block, smem, stream}; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum; result valid in thread 0.  `sh` must hold >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : T(0);
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}
#endif
