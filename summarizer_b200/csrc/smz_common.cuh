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

// Number of SMs of the current device (cached per device id).
int sm_count();
// Host -> device upload of a small descriptor array that is safe to capture in a CUDA graph: up to 3584 bytes travel
// as a by-value kernel parameter (copied into the launch / graph node), so no host buffer has to outlive the call;
// larger arrays go through cudaMemcpyAsync as before.
int upload_small(void *dst, const void *src, size_t bytes, cudaStream_t st);
// Max opt-in dynamic shared memory per block of the current device.
int max_smem_optin();


}  // namespace smz
