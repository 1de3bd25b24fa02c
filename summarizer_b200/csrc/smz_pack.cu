// Annotator summaries as 1 bit per frame: the host-to-device staging form of utils/eval.py:125-165's
// `user_summary` input.  evaluate_summary binarises it first (`user_summary[user_summary > 0] = 1`, eval.py:148-149),
// so (x > 0) is all the F-score ever reads: packing it on the host before the PCIe copy moves 32x fewer bytes than
// the float32 rows (2.4 MB -> 75 KB per sweep video), and the device side becomes a popcount over words.
//   smz_host_pack_user_summary  host threads, SSE2 compare + movemask (runs on the CPU by design: it is the copy's
//                               staging step, not a compute fallback — no result is produced here)
//   smz_fscore_packed           device: overlap / annotator counts from the packed rows, then the same
//                               fscore_final arithmetic as smz_fscore (bit-identical F).
#include <emmintrin.h>

#include <thread>
#include <vector>

#include "smz_common.cuh"

namespace smz {
int launch_fscore_final(const smz_video_desc *desc, int n_videos, const int32_t *msum, const int32_t *overlap,
                        const int32_t *gsum, float *f, double *avg_f, double *max_f, cudaStream_t st);
}

namespace {

void pack_row(const float *src, int n, uint32_t *dst) {
    const __m128 zero = _mm_setzero_ps();
    const int full = n / 32;
    for (int w = 0; w < full; w++) {
        const float *p = src + 32 * w;
        uint32_t bits = 0;
        for (int q = 0; q < 8; q++)
            bits |= (uint32_t)_mm_movemask_ps(_mm_cmpgt_ps(_mm_loadu_ps(p + 4 * q), zero)) << (4 * q);
        dst[w] = bits;
    }
    if (n % 32) {
        uint32_t bits = 0;
        for (int j = 32 * full; j < n; j++) bits |= (src[j] > 0.f ? 1u : 0u) << (j & 31);
        dst[full] = bits;
    }
}

__global__ void __launch_bounds__(256)
fscore_bits_kernel(const smz_video_desc *__restrict__ desc, int v0, const uint32_t *__restrict__ ubits,
                   const int64_t *__restrict__ bits_off, const uint32_t *__restrict__ mask,
                   int32_t *__restrict__ overlap, int32_t *__restrict__ gsum) {
    const int v = v0 + blockIdx.x;
    const smz_video_desc d = desc[v];
    const int W = (d.n_frames + 31) >> 5;
    const uint32_t *vm = mask + d.mask_off;
    const uint32_t *ub = ubits + bits_off[v];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int u = warp; u < d.n_users; u += 8) {
        const uint32_t *row = ub + (int64_t)u * W;
        int ov = 0, gs = 0;
        for (int w = lane; w < W; w += 32) {
            const uint32_t x = __ldg(row + w);
            ov += __popc(x & __ldg(vm + w));
            gs += __popc(x);
        }
        ov = __reduce_add_sync(0xffffffffu, ov);
        gs = __reduce_add_sync(0xffffffffu, gs);
        if (lane == 0) { overlap[d.ucount_off + u] = ov; gsum[d.ucount_off + u] = gs; }
    }
}

}  // namespace

namespace smz {
int launch_fscore_bits(const smz_video_desc *desc, int n_videos, const uint32_t *user_bits, const int64_t *bits_off,
                       const uint32_t *mask, int32_t *overlap, int32_t *gsum, cudaStream_t st) {
    fscore_bits_kernel<<<n_videos, 256, 0, st>>>(desc, 0, user_bits, bits_off, mask, overlap, gsum);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
}  // namespace smz

// h_desc / h_user / h_bits_off / h_bits are HOST pointers.  Row u of video v: h_user + user_off + u*user_ld
// (n_frames floats) -> h_bits + h_bits_off[v] + u*ceil(n_frames/32) words, bit j of word w = frame 32w+j > 0.
extern "C" int smz_host_pack_user_summary(const smz_video_desc *h_desc, int n_videos, const float *h_user,
                                          const int64_t *h_bits_off, uint32_t *h_bits, int n_threads) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && h_desc && h_user && h_bits_off && h_bits, "host_pack_user_summary: bad argument");
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_videos) n_threads = n_videos;
    auto work = [=](int t) {
        for (int v = t; v < n_videos; v += n_threads) {
            const smz_video_desc &d = h_desc[v];
            const int W = (d.n_frames + 31) >> 5;
            for (int u = 0; u < d.n_users; u++)
                pack_row(h_user + d.user_off + (int64_t)u * d.user_ld, d.n_frames, h_bits + h_bits_off[v] + (int64_t)u * W);
        }
    };
    if (n_threads == 1) { work(0); return SMZ_OK; }
    std::vector<std::thread> pool;
    pool.reserve(n_threads);
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
    return SMZ_OK;
}

extern "C" int smz_fscore_packed(const smz_video_desc *desc, int n_videos, const uint32_t *user_bits, const int64_t *bits_off,
                                 const uint32_t *mask, const int32_t *msum, int32_t *overlap, int32_t *gsum, float *f,
                                 double *avg_f, double *max_f, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0, "negative size");
    SMZ_REQUIRE(desc && user_bits && bits_off && mask && msum && overlap && gsum && f, "NULL pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = smz::launch_fscore_bits(desc, n_videos, user_bits, bits_off, mask, overlap, gsum, st);
    if (rc != SMZ_OK) return rc;
    return smz::launch_fscore_final(desc, n_videos, msum, overlap, gsum, f, avg_f, max_f, st);
}
