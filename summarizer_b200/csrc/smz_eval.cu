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
#include "smz_eval_dev.cuh"

#include <limits.h>

namespace {

constexpr int FSCORE_THREADS = 256;
constexpr int FSCORE_MAX_USERS = 1024;
constexpr int FSCORE_UNROLL = 4;
static_assert(SMZ_FSCORE_CHUNK == FSCORE_THREADS * 8, "each thread owns two float4 columns");

using namespace smzdev;

// ------------------------------------------------------------------------------------------
// fscore kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// 4-bit mask of (g > 0) for the frames f..f+3 of one annotator row, frames >= n_frames masked off
__device__ __forceinline__ uint32_t pos_bits(const float4 x, int f, int n_frames) {
    uint32_t b = (x.x > 0.f ? 1u : 0u) | (x.y > 0.f ? 2u : 0u) | (x.z > 0.f ? 4u : 0u) | (x.w > 0.f ? 8u : 0u);
    const int rem = n_frames - f;  // > 0 here
    if (rem < 4) b &= (1u << rem) - 1u;
    return b;
}

__global__ void __launch_bounds__(FSCORE_THREADS)
fscore_kernel(const smz_video_desc *__restrict__ desc, int v0, const float *__restrict__ user,
              const uint32_t *__restrict__ mask, int32_t *__restrict__ overlap, int32_t *__restrict__ gsum) {
    __shared__ uint32_t s_cnt[FSCORE_MAX_USERS];
    const int v = v0 + blockIdx.y;
    const smz_video_desc d = desc[v];
    const int n_frames = d.n_frames;
    const int f_base = blockIdx.x * SMZ_FSCORE_CHUNK;
    if (f_base >= n_frames) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_users = d.n_users;
    for (int u = tid; u < n_users; u += FSCORE_THREADS) s_cnt[u] = 0u;
    __syncthreads();

    const int fa = f_base + 4 * tid;
    const int fb = fa + 4 * FSCORE_THREADS;
    const bool in_a = fa < n_frames, in_b = fb < n_frames;
    const uint32_t *vm = mask + d.mask_off;
    const uint32_t ma = in_a ? ((__ldg(vm + (fa >> 5)) >> (fa & 31)) & 0xfu) : 0u;
    const uint32_t mb = in_b ? ((__ldg(vm + (fb >> 5)) >> (fb & 31)) & 0xfu) : 0u;
    const float *base = user + d.user_off;
    const int64_t ld = d.user_ld;
    const bool vec = (((d.user_off | ld) & 3) == 0) && ((reinterpret_cast<uintptr_t>(user) & 15) == 0);

    for (int u0 = 0; u0 < n_users; u0 += FSCORE_UNROLL) {
        float4 xa[FSCORE_UNROLL], xb[FSCORE_UNROLL];
#pragma unroll
        for (int k = 0; k < FSCORE_UNROLL; k++) {
            xa[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            xb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u0 + k < n_users) {
                const float *row = base + (int64_t)(u0 + k) * ld;
                if (vec) {
                    if (in_a) xa[k] = ld_stream_f4(row + fa);
                    if (in_b) xb[k] = ld_stream_f4(row + fb);
                } else {
                    if (in_a) {
                        xa[k].x = ld_stream_f1(row + fa);
                        if (fa + 1 < n_frames) xa[k].y = ld_stream_f1(row + fa + 1);
                        if (fa + 2 < n_frames) xa[k].z = ld_stream_f1(row + fa + 2);
                        if (fa + 3 < n_frames) xa[k].w = ld_stream_f1(row + fa + 3);
                    }
                    if (in_b) {
                        xb[k].x = ld_stream_f1(row + fb);
                        if (fb + 1 < n_frames) xb[k].y = ld_stream_f1(row + fb + 1);
                        if (fb + 2 < n_frames) xb[k].z = ld_stream_f1(row + fb + 2);
                        if (fb + 3 < n_frames) xb[k].w = ld_stream_f1(row + fb + 3);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < FSCORE_UNROLL; k++) {
            if (u0 + k < n_users) {  // uniform across the CTA
                const uint32_t ga = in_a ? pos_bits(xa[k], fa, n_frames) : 0u;
                const uint32_t gb = in_b ? pos_bits(xb[k], fb, n_frames) : 0u;
                // low half: overlap, high half: annotator ones (per-CTA totals <= 2048 < 2^16)
                uint32_t packed = (uint32_t)(__popc(ga & ma) + __popc(gb & mb)) |
                                  ((uint32_t)(__popc(ga) + __popc(gb)) << 16);
                packed = __reduce_add_sync(0xffffffffu, packed);
                if (lane == 0 && packed) atomicAdd(&s_cnt[u0 + k], packed);
            }
        }
    }
    __syncthreads();
    for (int u = tid; u < n_users; u += FSCORE_THREADS) {
        const uint32_t c = s_cnt[u];
        if (c & 0xffffu) atomicAdd(overlap + d.ucount_off + u, (int)(c & 0xffffu));
        if (c >> 16) atomicAdd(gsum + d.ucount_off + u, (int)(c >> 16));
    }
}

// utils/eval.py:151-164 in float32 (numpy 2 / NEP 50 semantics when no zero padding happened)
__global__ void fscore_final_kernel(const smz_video_desc *__restrict__ desc, int n_videos,
                                    const int32_t *__restrict__ msum, const int32_t *__restrict__ overlap,
                                    const int32_t *__restrict__ gsum, float *__restrict__ f,
                                    double *__restrict__ avg_f, double *__restrict__ max_f) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_videos) return;
    const smz_video_desc d = desc[v];
    const float ms = __fadd_rn((float)msum[v], 1e-8f);
    float mx = 0.f;
    bool any_zero = false;  // a Python-float 0. entry promotes the reference's list to float64
    float *fv = f + d.ucount_off;
    for (int u = 0; u < d.n_users; u++) {
        const float ov = (float)overlap[d.ucount_off + u];
        const float gs = __fadd_rn((float)gsum[d.ucount_off + u], 1e-8f);
        const float precision = __fdiv_rn(ov, ms);
        const float recall = __fdiv_rn(ov, gs);
        float fs = 0.f;
        if (!(precision == 0.f && recall == 0.f))
            fs = __fdiv_rn(__fmul_rn(__fmul_rn(2.f, precision), recall), __fadd_rn(precision, recall));
        else
            any_zero = true;
        fv[u] = fs;
        mx = (u == 0) ? fs : fmaxf(mx, fs);
    }
    if (avg_f) {
        if (d.n_users <= 0) avg_f[v] = 0.;
        else if (any_zero) {
            ArrayCursorF64 c64{fv};
            avg_f[v] = __ddiv_rn(pw_sum<double>(c64, 0, d.n_users), (double)d.n_users);
        }
        else {
            ArrayCursor cur{fv};
            avg_f[v] = (double)__fdiv_rn(pw_sum<float>(cur, 0, d.n_users), (float)d.n_users);
        }
    }
    if (max_f) max_f[v] = (double)mx;
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
    const int chunks = (max_n_frames + SMZ_FSCORE_CHUNK - 1) / SMZ_FSCORE_CHUNK;
    if (chunks > 0) {
        for (int v0 = 0; v0 < n_videos; v0 += 65535) {
            const int nv = n_videos - v0 < 65535 ? n_videos - v0 : 65535;
            fscore_kernel<<<dim3(chunks, nv), FSCORE_THREADS, 0, st>>>(desc, v0, user_summary, mask, overlap, gsum);
        }
        SMZ_CUDA_CHECK(cudaGetLastError());
    }
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
