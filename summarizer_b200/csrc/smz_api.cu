// Library-level entry points: version, error string, device check.
#include "smz_common.cuh"

#include <stdlib.h>

namespace smz {

char *last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

bool debug_sync_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SMZ_DEBUG_SYNC"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int max_smem_optin() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 227 * 1024;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0)
        return 227 * 1024;
    return v;
}

}  // namespace smz

extern "C" const char *smz_version(void) { return "summarizer_b200 0.1 (sm_100a)"; }

extern "C" const char *smz_last_error(void) { return smz::last_error_buf(); }

extern "C" int smz_device_check(void) {
    int dev = 0, major = 0, minor = 0;
    SMZ_CUDA_CHECK(cudaGetDevice(&dev));
    SMZ_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SMZ_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10)
        return smz::fail(SMZ_ERR_DEVICE, "device %d is sm_%d%d; libsummarizer_b200 only contains sm_100a code "
                         "(no other backend, no CPU fallback)", dev, major, minor);
    return SMZ_OK;
}
