// Shot selection and F-score evaluation kernels (sm_100a) + their C ABI.
//
// Replaces, on the device, the reference's host-side numpy/Python/OR-tools path
//   utils/eval.py:15-35   upsample
//   utils/eval.py:74-123  generate_summary   (segment mean -> knapsack/rank -> summary vector)
//   utils/knapsack.py:5-23 knapsack_ortools  (OR-tools 7.5 KnapsackDynamicProgrammingSolver)
//   utils/eval.py:125-165 evaluate_summary   (per-user overlap / precision / recall / F)
// with results that are bit-identical for every integer (values, selections, masks, counts)
// and for the float32 quantities the reference computes in float32 (segment means, F).
//
// Kernels
//   select_kernel   one persistent CTA per video (grid = #SMs x occupancy): segment pooling in
//                   numpy's pairwise summation order, value quantisation in fp64, 0/1 knapsack DP
//                   with the DP row in shared memory (double buffered, one __syncthreads per item),
//                   a take-bit matrix (shared memory when it fits, else the work buffer) and the
//                   OR-tools backtrack, then summary vector + bit mask + popcount.
//                   Shared-memory/latency bound.
//   fscore_kernel   grid (frame chunk, video); streams user_summary once with 128-bit loads,
//                   counts with integer ops, one packed REDUX per user per warp.  HBM bound.
//   fscore_final    one thread per video: float32 F per user, numpy-order mean, max.
#include "smz_common.cuh"
#include "smz_fscore_dev.cuh"

#include <limits.h>

namespace {

using namespace smzdev;

// ------------------------------------------------------------------------------------------
// fscore kernels
// ------------------------------------------------------------------------------------------
// grid (frame range, video): a CTA streams `span` frames (a multiple of SMZ_FSCORE_CHUNK) of all annotator rows.  Large
// batches give every CTA a whole video (one warp reduction per row and video), small ones split videos for parallelism.
__global__ void __launch_bounds__(FSCORE_THREADS, 6)
fscore_kernel(const smz_video_desc *__restrict__ desc, int v0, int span, const float *__restrict__ user,
              const uint32_t *__restrict__ mask, int32_t *__restrict__ overlap, int32_t *__restrict__ gsum) {
    __shared__ uint32_t s_ov[FSCORE_MAX_USERS], s_gs[FSCORE_MAX_USERS];
    const smz_video_desc d = desc[v0 + blockIdx.y];
    const int f_base = blockIdx.x * span;
    if (f_base >= d.n_frames) return;
    const int n_users = min(d.n_users, FSCORE_MAX_USERS);      // the C ABI callers reject larger counts up front
    for (int u = threadIdx.x; u < n_users; u += FSCORE_THREADS) { s_ov[u] = 0u; s_gs[u] = 0u; }
    __syncthreads();
    fscore_rows_acc<FSCORE_UNROLL>(d, f_base, f_base + span, user, mask + d.mask_off, s_ov, s_gs);
    __syncthreads();
    for (int u = threadIdx.x; u < n_users; u += FSCORE_THREADS) {
        if (s_ov[u]) atomicAdd(overlap + d.ucount_off + u, (int)s_ov[u]);
        if (s_gs[u]) atomicAdd(gsum + d.ucount_off + u, (int)s_gs[u]);
    }
}

__global__ void fscore_final_kernel(const smz_video_desc *__restrict__ desc, int n_videos,
                                    const int32_t *__restrict__ msum, const int32_t *__restrict__ overlap,
                                    const int32_t *__restrict__ gsum, float *__restrict__ f,
                                    double *__restrict__ avg_f, double *__restrict__ max_f) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_videos) return;
    const smz_video_desc d = desc[v];
    fscore_final_video(d.n_users, msum[v], overlap + d.ucount_off, gsum + d.ucount_off, f + d.ucount_off,
                       avg_f ? avg_f + v : nullptr, max_f ? max_f + v : nullptr, d.summ_len < d.n_frames);
}

// utils/eval.py:136-145: binarise, truncate / zero-pad to n_frames, pack 32 frames per word
__global__ void pack_summary_kernel(const smz_video_desc *__restrict__ desc, const float *__restrict__ machine,
                                    uint32_t *__restrict__ mask, int32_t *__restrict__ msum) {
    const int v = blockIdx.y;
    const smz_video_desc d = desc[v];
    const int mwords = (d.n_frames + 31) >> 5;
    const int lim = min(d.n_frames, d.summ_len);
    const float *m = machine + d.summ_off;
    int cnt = 0;
    for (int wd = blockIdx.x * blockDim.x / 32 + (threadIdx.x >> 5); wd < mwords; wd += gridDim.x * blockDim.x / 32) {
        const int fidx = (wd << 5) + (threadIdx.x & 31);
        const bool on = fidx < lim && m[fidx] > 0.f;
        const uint32_t b = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0) { mask[d.mask_off + wd] = b; cnt += __popc(b); }
    }
    if (cnt) atomicAdd(msum + v, cnt);
}

__global__ void upsample_kernel(const smz_video_desc *__restrict__ desc, const float *__restrict__ scores,
                                const int32_t *__restrict__ picks, float *__restrict__ out,
                                int32_t *__restrict__ status) {
    const int v = blockIdx.y;
    const smz_video_desc d = desc[v];
    FrameCursor cur;
    cur.init(scores + d.score_off, picks + d.picks_off, d.n_scores, d.n_picks, d.n_frames);
    if (status && blockIdx.x == 0 && threadIdx.x == 0)
        status[v] = (cur.n_bound - 1 > d.n_scores + 1) ? SMZ_STATUS_INTERVALS : 0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < d.n_frames; f += gridDim.x * blockDim.x) {
        cur.seek(f);
        out[d.frame_off + f] = cur.at(f);
    }
}

}  // namespace

namespace smz {
// overlap / annotator counts of a range of videos; the caller has zeroed overlap[] and gsum[]
int launch_fscore_counts(const smz_video_desc *desc, int n_videos, int max_n_frames, const float *user_summary,
                         const uint32_t *mask, int32_t *overlap, int32_t *gsum, cudaStream_t st) {
    const int chunks = (max_n_frames + SMZ_FSCORE_CHUNK - 1) / SMZ_FSCORE_CHUNK;
    if (chunks > 0) {
        // chunks per CTA: as many as still leave ~4 waves of CTAs (6 resident per SM)
        int64_t per = ((int64_t)n_videos * chunks) / ((int64_t)smz::sm_count() * 6 * 4);
        if (per < 1) per = 1;
        if (per > chunks) per = chunks;
        const int ranges = (int)((chunks + per - 1) / per);
        for (int v0 = 0; v0 < n_videos; v0 += 65535) {
            const int nv = n_videos - v0 < 65535 ? n_videos - v0 : 65535;
            fscore_kernel<<<dim3(ranges, nv), FSCORE_THREADS, 0, st>>>(desc, v0, (int)per * SMZ_FSCORE_CHUNK, user_summary, mask,
                                                                     overlap, gsum);
        }
        SMZ_CUDA_CHECK(cudaGetLastError());
    }
    return SMZ_OK;
}

int launch_fscore_final(const smz_video_desc *desc, int n_videos, const int32_t *msum, const int32_t *overlap,
                        const int32_t *gsum, float *f, double *avg_f, double *max_f, cudaStream_t st) {
    fscore_final_kernel<<<(n_videos + 127) / 128, 128, 0, st>>>(desc, n_videos, msum, overlap, gsum, f, avg_f, max_f);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
}  // namespace smz

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int smz_fscore(const smz_video_desc *desc, int n_videos, int max_n_frames, int total_users,
                          const float *user_summary, const uint32_t *mask, const int32_t *msum,
                          int32_t *overlap, int32_t *gsum, float *f, double *avg_f, double *max_f, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && max_n_frames >= 0 && total_users >= 0, "negative size");
    SMZ_REQUIRE(desc && user_summary && mask && msum && overlap && gsum && f, "NULL pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SMZ_CUDA_CHECK(cudaMemsetAsync(overlap, 0, sizeof(int32_t) * (size_t)total_users, st));
    SMZ_CUDA_CHECK(cudaMemsetAsync(gsum, 0, sizeof(int32_t) * (size_t)total_users, st));
    rc = smz::launch_fscore_counts(desc, n_videos, max_n_frames, user_summary, mask, overlap, gsum, st);
    if (rc != SMZ_OK) return rc;
    return smz::launch_fscore_final(desc, n_videos, msum, overlap, gsum, f, avg_f, max_f, st);
}

extern "C" int smz_pack_summary(const smz_video_desc *desc, int n_videos, int max_n_frames, const float *machine,
                                uint32_t *mask, int32_t *msum, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && n_videos <= 65535, "pack_summary: 1..65535 videos per call");
    SMZ_REQUIRE(desc && machine && mask && msum, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SMZ_CUDA_CHECK(cudaMemsetAsync(msum, 0, sizeof(int32_t) * (size_t)n_videos, st));
    const int mwords = (max_n_frames + 31) / 32;
    int gx = (mwords + 7) / 8;  // 8 warps per CTA, one word per warp per step
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    pack_summary_kernel<<<dim3(gx, n_videos), 256, 0, st>>>(desc, machine, mask, msum);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_upsample(const smz_video_desc *desc, int n_videos, int max_n_frames, const float *scores,
                            const int32_t *picks, float *frame_scores, int32_t *status, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && n_videos <= 65535, "upsample: 1..65535 videos per call");
    SMZ_REQUIRE(desc && scores && picks && frame_scores, "NULL pointer");
    int gx = (max_n_frames + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 256) gx = 256;
    upsample_kernel<<<dim3(gx, n_videos), 256, 0, (cudaStream_t)stream>>>(desc, scores, picks, frame_scores, status);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
