// Device code shared by the stand-alone F-score kernels (smz_eval.cu) and the evaluation tail fused into the
// knapsack kernels (smz_select.cu): utils/eval.py:125-165 evaluate_summary.
#pragma once
#include "smz_eval_dev.cuh"
#include "../../include/summarizer_b200.h"

namespace smzdev {

constexpr int FSCORE_THREADS = 256;
constexpr int FSCORE_MAX_USERS = 1024;
constexpr int FSCORE_UNROLL = 4;
static_assert(SMZ_FSCORE_CHUNK == FSCORE_THREADS * 8, "each thread owns two float4 columns");

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

// One chunk of SMZ_FSCORE_CHUNK frames of one video, all annotator rows against the summary mask `vm` (the video's
// mask words, in global OR shared memory): adds the overlap / annotator-ones counts of the chunk to s_ov[u] / s_gs[u]
// (shared memory, zeroed by the caller).  Called by all FSCORE_THREADS threads of the CTA.
__device__ __forceinline__ void fscore_chunk_acc(const smz_video_desc &d, int f_base, const float *__restrict__ user,
                                                 const uint32_t *vm, uint32_t *s_ov, uint32_t *s_gs) {
    const int n_frames = d.n_frames;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_users = min(d.n_users, FSCORE_MAX_USERS);
    const int fa = f_base + 4 * tid;
    const int fb = fa + 4 * FSCORE_THREADS;
    const bool in_a = fa < n_frames, in_b = fb < n_frames;
    const uint32_t ma = in_a ? ((vm[fa >> 5] >> (fa & 31)) & 0xfu) : 0u;
    const uint32_t mb = in_b ? ((vm[fb >> 5] >> (fb & 31)) & 0xfu) : 0u;
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
                // low half: overlap, high half: annotator ones (per-warp totals <= 256 < 2^16): one REDUX for both
                uint32_t packed = (uint32_t)(__popc(ga & ma) + __popc(gb & mb)) |
                                  ((uint32_t)(__popc(ga) + __popc(gb)) << 16);
                packed = __reduce_add_sync(0xffffffffu, packed);
                if (lane == 0 && packed) {
                    if (packed & 0xffffu) atomicAdd(&s_ov[u0 + k], packed & 0xffffu);
                    atomicAdd(&s_gs[u0 + k], packed >> 16);
                }
            }
        }
    }
}

// utils/eval.py:151-164 for ONE video in float32 (numpy 2 / NEP 50 semantics when no zero padding happened); one thread.
// overlap / gsum: the video's per-annotator counts (any address space); f: the video's slice of the F output.
__device__ inline void fscore_final_video(int n_users, int msum, const int32_t *overlap, const int32_t *gsum, float *fv,
                                          double *avg_f, double *max_f) {
    const float ms = __fadd_rn((float)msum, 1e-8f);
    float mx = 0.f;
    bool any_zero = false;  // a Python-float 0. entry promotes the reference's list to float64
    for (int u = 0; u < n_users; u++) {
        const float ov = (float)overlap[u];
        const float gs = __fadd_rn((float)gsum[u], 1e-8f);
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
        if (n_users <= 0) *avg_f = 0.;
        else if (any_zero) {
            ArrayCursorF64 c64{fv};
            *avg_f = __ddiv_rn(pw_sum<double>(c64, 0, n_users), (double)n_users);
        } else {
            ArrayCursor cur{fv};
            *avg_f = (double)__fdiv_rn(pw_sum<float>(cur, 0, n_users), (float)n_users);
        }
    }
    if (max_f) *max_f = (double)mx;
}

}  // namespace smzdev
