// Library-level entry points: version, error string, device check.
#include "smz_common.cuh"

#include <string.h>

#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

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

bool profile_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SMZ_PROFILE"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

namespace {
struct Mark { cudaEvent_t ev; const char *step; cudaStream_t st; };
std::vector<Mark> &marks() { static std::vector<Mark> m; return m; }
}  // namespace

void profile_mark(cudaStream_t st, const char *step) {
    if (!profile_enabled()) return;
    Mark m;
    m.step = step; m.st = st;
    if (cudaEventCreate(&m.ev) != cudaSuccess) return;
    cudaEventRecord(m.ev, st);
    marks().push_back(m);
}

void profile_report() {
    if (!profile_enabled()) return;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<Mark> &m = marks();
    // consecutive marks on the same stream delimit a step
    std::map<cudaStream_t, size_t> last;
    for (size_t i = 0; i < m.size(); i++) {
        auto it = last.find(m[i].st);
        if (it != last.end() && m[it->second].step[0] != 0) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, m[it->second].ev, m[i].ev) == cudaSuccess) {
                auto &a = acc[m[it->second].step];
                a.first += ms; a.second += 1;
            }
        }
        last[m[i].st] = i;
    }
    double tot = 0.;
    for (auto &kv : acc) tot += kv.second.first;
    fprintf(stderr, "[smz profile] %zu marks, %.3f ms of stream time\n", m.size(), tot);
    for (auto &kv : acc)
        fprintf(stderr, "[smz profile] %-14s %9.3f ms %6.1f%%  %5d launches  %8.1f us/launch\n", kv.first.c_str(), kv.second.first,
                100. * kv.second.first / tot, kv.second.second, 1e3 * kv.second.first / kv.second.second);
    for (auto &x : m) cudaEventDestroy(x.ev);
    m.clear();
}

namespace {
struct SmallBlob { unsigned char b[3584]; };
__global__ void store_blob_kernel(const __grid_constant__ SmallBlob blob, unsigned char *__restrict__ dst, int n) {
    smz::pdl_trigger();
    smz::pdl_wait();          // dst may still be read by the previous kernel of the stream
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = blob.b[i];
}
}  // namespace

int upload_small(void *dst, const void *src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return SMZ_OK;
    if (bytes <= sizeof(SmallBlob)) {
        SmallBlob blob;
        memcpy(blob.b, src, bytes);
        SMZ_CUDA_CHECK(launch_pdl(store_blob_kernel, dim3(1), dim3(256), 0, st, blob, reinterpret_cast<unsigned char *>(dst), (int)bytes));
        return SMZ_OK;
    }
    SMZ_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return SMZ_OK;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SMZ_PDL"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
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

extern "C" void smz_profile_report(void) { smz::profile_report(); }
