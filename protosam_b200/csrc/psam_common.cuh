// Shared host/device helpers of libpsam_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/psam_b200.h"

namespace psam {

// ---- tracing: every entry point of the C ABI that enqueues work is an NVTX range (header-only NVTX3: a no-op unless
// a tool such as nsys / ncu --nvtx is attached), so timelines and `ncu --nvtx-include` can be cut per stage ----
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define PSAM_TRACE(name) psam::NvtxRange nvtx_range__(name)

// ---- host-side error plumbing (thread-local message, see psam_last_error) ----
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// per-kernel device timing (psam_profile_enable / psam_profile_collect): events around every launch
extern bool g_profiling;
void prof_begin(cudaStream_t stream);
void prof_end(const char* what, cudaStream_t stream);
#define PSAM_PROF_BEGIN(stream)                      \
    do {                                             \
        if (psam::g_profiling) psam::prof_begin(stream); \
    } while (0)

// Every kernel of the library asks for the maximum shared-memory carveout (once per kernel and device): CTAs of two
// kernels whose L1/shared splits differ cannot be resident on one SM at the same time, and the pipeline relies on the
// ALU-bound prompt kernels of one volume running beside the tensor-bound GEMM CTA of the next.
void max_carveout_once(const void* kernel, bool* done_per_device);
#define PSAM_MAX_CARVEOUT(kernel)                                                     \
    do {                                                                              \
        static bool done__[64] = {};                                                  \
        psam::max_carveout_once(reinterpret_cast<const void*>(kernel), done__);       \
    } while (0)

#define PSAM_CHECK_ARG(cond, ...)                     \
    do {                                              \
        if (!(cond)) {                                \
            psam::set_error(__VA_ARGS__);             \
            return PSAM_ERR_ARG;                      \
        }                                             \
    } while (0)

#define PSAM_CHECK_LAUNCH(what)                                                        \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            psam::set_error("%s: %s", what, cudaGetErrorString(e__));                  \
            return PSAM_ERR_LAUNCH;                                                    \
        }                                                                              \
        psam::count_launch();                                                          \
        if (psam::g_profiling) psam::prof_end(what, stream);                           \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of the caller's workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n)
    {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

// ---- optional CTA-level trace (tools/trace_timeline.py): while a trace buffer is installed (psam_trace_install), the big
// kernels record, per CTA, the SM it ran on and its start / end time (globaltimer, ns).  Every translation unit has its
// own copy of the pointer; psam_trace_install sets all of them.  One predictable branch per CTA when off. ----
struct TraceRec {
    unsigned long long t0, t1;
    unsigned int smid, kernel, cta, pad;
};
static __device__ TraceRec* g_trace_buf = nullptr;     // [0] is the header: t0 = number of records taken so far, t1 = capacity
void trace_register(void (*setter)(TraceRec*));

__device__ __forceinline__ unsigned long long trace_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// call from ONE thread at the start of a CTA; returns the record (or nullptr) to hand to trace_end.  Compiled in only
// with -DPSAM_TRACE_KERNELS (make TRACE=1): the production library carries no trace code in its kernels.
#ifdef PSAM_TRACE_KERNELS
__device__ __forceinline__ TraceRec* trace_begin(unsigned int kernel)
{
    TraceRec* buf = g_trace_buf;
    if (buf == nullptr) return nullptr;
    const unsigned long long slot = atomicAdd(&buf[0].t0, 1ull) + 1ull;
    if (slot >= buf[0].t1) return nullptr;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    TraceRec* r = buf + slot;
    r->smid = smid; r->kernel = kernel; r->cta = blockIdx.x + gridDim.x * blockIdx.y; r->pad = 0;
    r->t0 = trace_now(); r->t1 = 0;
    return r;
}

__device__ __forceinline__ void trace_end(TraceRec* r)
{
    if (r) r->t1 = trace_now();
}
constexpr bool kTraceCompiled = true;
#else
__device__ __forceinline__ TraceRec* trace_begin(unsigned int) { return nullptr; }
__device__ __forceinline__ void trace_end(TraceRec*) {}
constexpr bool kTraceCompiled = false;
#endif

#define PSAM_TRACE_TU()                                                                     \
    static void trace_set_this_tu(psam::TraceRec* p) { cudaMemcpyToSymbol(psam::g_trace_buf, &p, sizeof(p)); } \
    static const int trace_registered__ = (psam::trace_register(trace_set_this_tu), 0)

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_excl_scan_i(int v, int lane)
{
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    return inc - v;
}

}  // namespace psam
