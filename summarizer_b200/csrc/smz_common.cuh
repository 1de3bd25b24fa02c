// Shared host-side helpers of libsummarizer_b200.so (error convention of include/summarizer_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/summarizer_b200.h"

namespace smz {

char *last_error_buf();  // thread-local, 512 bytes (smz_api.cu)

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SMZ_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::smz::fail(SMZ_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,        \
                               cudaGetErrorString(_e));                                        \
    } while (0)

#define SMZ_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) return ::smz::fail(SMZ_ERR_ARG, __VA_ARGS__);                             \
    } while (0)

// SMZ_DEBUG_SYNC=1: synchronise after every launch and report the step that failed (debugging aid;
// never set in production — the library is otherwise fully asynchronous).
bool debug_sync_enabled();
#define SMZ_DEBUG_STEP(st, name)                                                               \
    do {                                                                                       \
        if (::smz::debug_sync_enabled()) {                                                     \
            cudaError_t _e = cudaStreamSynchronize(st);                                        \
            if (_e != cudaSuccess)                                                             \
                return ::smz::fail(SMZ_ERR_CUDA, "step '%s' failed: %s", name, cudaGetErrorString(_e)); \
        }                                                                                      \
    } while (0)

// SMZ_PROFILE=1: per-step device timing of the scorer pipelines (CUDA events around every launch, summary
// printed to stderr by smz_profile_report()).  Development aid; off by default.
bool profile_enabled();
void profile_mark(cudaStream_t st, const char *step);   // call before a step; "" closes the last one
void profile_report();

// Programmatic dependent launch (PDL).  A batch-1 training step is ~35 kernels of 3-15 us: the drain -> launch gap and each
// kernel's prologue (barrier init, TMEM allocation, descriptor prefetch) are a third of it.  Kernels on these chains call
// pdl_trigger() first thing (the NEXT kernel of the stream may start being scheduled) and pdl_wait() before their first
// access to global memory (returns when the PREVIOUS kernel has completed and its writes are visible); launched through
// launch_pdl() they carry cudaLaunchAttributeProgrammaticStreamSerialization, so their prologue overlaps the predecessor's
// tail.  Both instructions are no-ops in a kernel launched without the attribute / without a PDL successor, and a CUDA-graph
// capture records the edge as a programmatic dependency.  SMZ_PDL=0 launches everything fully serialised.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// Philox4x32-10 counter-based generator (Salmon et al., SC'11): 128 random bits per (counter, key)
#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c.x), l0 = 0xD2511F53u * c.x;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c.z), l1 = 0xCD9E8D57u * c.z;
        c = make_uint4(h1 ^ c.y ^ k.x, l1, h0 ^ c.w ^ k.y, l0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Number of SMs of the current device (cached per device id).
int sm_count();
// Host -> device upload of a small descriptor array that is safe to capture in a CUDA graph: up to 3584 bytes travel
// as a by-value kernel parameter (copied into the launch / graph node), so no host buffer has to outlive the call;
// larger arrays go through cudaMemcpyAsync as before.
int upload_small(void *dst, const void *src, size_t bytes, cudaStream_t st);
// Max opt-in dynamic shared memory per block of the current device.
int max_smem_optin();


}  // namespace smz
