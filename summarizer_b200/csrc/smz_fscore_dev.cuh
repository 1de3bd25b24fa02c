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

// 1 when x > 0, else 0, from the bit pattern: positive floats (denormals included) are positive int32, zeros and
// negatives are <= 0 — one DPX instruction, VIMNMX.RELU = max(min(bits, 1), 0).  (A NaN with a clear sign bit counts
// as positive; numpy's `x > 0` is False for it, but a NaN annotation makes the reference's F-score NaN anyway.)
__device__ __forceinline__ uint32_t gt0_flag(float x) { return (uint32_t)__vimin_s32_relu(__float_as_int(x), 1); }

// bit j (0..3) = a[j] > 0, bit 4 + j = b[j] > 0.  8 VIMNMX.RELU + 7 shift-adds (LEA).
__device__ __forceinline__ uint32_t pos_bits8(const float4 a, const float4 b) {
    const uint32_t lo = (gt0_flag(a.x) + 2u * gt0_flag(a.y)) + 4u * (gt0_flag(a.z) + 2u * gt0_flag(a.w));
    const uint32_t hi = (gt0_flag(b.x) + 2u * gt0_flag(b.y)) + 4u * (gt0_flag(b.z) + 2u * gt0_flag(b.w));
    return lo + 16u * hi;
}

// 4-bit mask of (g > 0) for the frames f..f+3 of one annotator row read element by element (unaligned rows)
__device__ __forceinline__ uint32_t pos_bits_scalar(const float *row, int f, int n_frames) {
    uint32_t b = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (f + j < n_frames && ld_stream_f1(row + f + j) > 0.f) b |= 1u << j;
    return b;
}

// Frames [f0, f1) of one video (f0 a multiple of SMZ_FSCORE_CHUNK), all annotator rows against the summary mask `vm`
// (the video's mask words, in global OR shared memory): adds the overlap / annotator-ones counts to s_ov[u] / s_gs[u]
// (shared memory, zeroed by the caller).  Called by all FSCORE_THREADS threads of the CTA.
// UNROLL annotator rows are streamed together: per thread and chunk two 128-bit loads per row are in flight, the
// positive flags of 8 frames become one byte (pos_bits8), overlap and ones are counted with two POPC into a packed
// per-thread counter, and the warp reduction + shared-memory atomics happen once per row and call, not per chunk.
template <int UNROLL, bool VEC>
__device__ __forceinline__ void fscore_rows_acc_impl(const smz_video_desc &d, int f0, int f1, const float *__restrict__ user,
                                                     const uint32_t *vm, uint32_t *s_ov, uint32_t *s_gs) {
    const int n_frames = d.n_frames;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_users = min(d.n_users, FSCORE_MAX_USERS);
    const int64_t ld = d.user_ld;
    f1 = min(f1, n_frames);
    // chunks [f0, f_full) lie completely inside the video: no bounds logic in their loop
    const int f_full = f0 + (f1 - f0) / SMZ_FSCORE_CHUNK * SMZ_FSCORE_CHUNK;
    const int fa0 = f0 + 4 * tid;                               // this thread's first frame (second group: + 1024)
    for (int u0 = 0; u0 < n_users; u0 += UNROLL) {
        uint32_t cnt[UNROLL];
        const float *p[UNROLL];                                  // row pointers, advanced chunk by chunk
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            cnt[k] = 0u;
            p[k] = user + d.user_off + (int64_t)min(u0 + k, n_users - 1) * ld + fa0;   // rows past the end re-read the last one (ignored)
        }
        int fa = fa0;
        if constexpr (VEC) {
#pragma unroll 1
            for (; fa < f_full; fa += SMZ_FSCORE_CHUNK) {
                float4 xa[UNROLL], xb[UNROLL];
#pragma unroll
                for (int k = 0; k < UNROLL; k++) {
                    xa[k] = ld_stream_f4(p[k]);
                    xb[k] = ld_stream_f4(p[k] + 4 * FSCORE_THREADS);
                    p[k] += SMZ_FSCORE_CHUNK;
                }
                // summary bits of this thread's 8 frames: fa % 32 and (fa + 1024) % 32 are the same nibble position
                const uint32_t sh = fa & 31;
                const uint32_t m8 = ((vm[fa >> 5] >> sh) & 0xfu) | (((vm[(fa >> 5) + 4 * FSCORE_THREADS / 32] >> sh) & 0xfu) << 4);
#pragma unroll
                for (int k = 0; k < UNROLL; k++) {
                    const uint32_t g = pos_bits8(xa[k], xb[k]);
                    cnt[k] += (uint32_t)__popc(g & m8) + ((uint32_t)__popc(g) << 16);
                }
            }
        }
        // the last, partial chunk (and every chunk of unaligned rows): bounds-checked
        for (; fa - 4 * tid < f1; fa += SMZ_FSCORE_CHUNK) {
            const int fb = fa + 4 * FSCORE_THREADS;
            const bool in_a = fa < n_frames, in_b = fb < n_frames;
            // summary bits and valid-frame bits of this thread's 8 frames (low nibble: fa.., high nibble: fb..)
            uint32_t m8 = 0u, v8 = 0u;
            if (in_a) { m8 = (vm[fa >> 5] >> (fa & 31)) & 0xfu; v8 = n_frames - fa >= 4 ? 0xfu : (1u << (n_frames - fa)) - 1u; }
            if (in_b) { m8 |= ((vm[fb >> 5] >> (fb & 31)) & 0xfu) << 4; v8 |= (n_frames - fb >= 4 ? 0xfu : (1u << (n_frames - fb)) - 1u) << 4; }
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                uint32_t g;
                if constexpr (VEC) {
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);               // rows are padded to ld % 4 == 0
                    g = pos_bits8(in_a ? ld_stream_f4(p[k]) : z, in_b ? ld_stream_f4(p[k] + 4 * FSCORE_THREADS) : z) & v8;
                } else {
                    g = in_a ? pos_bits_scalar(p[k] - fa, fa, n_frames) : 0u;
                    if (in_b) g |= pos_bits_scalar(p[k] - fa, fb, n_frames) << 4;
                }
                cnt[k] += (uint32_t)__popc(g & m8) + ((uint32_t)__popc(g) << 16);
                p[k] += SMZ_FSCORE_CHUNK;
            }
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (u0 + k < n_users) {                                          // uniform across the CTA
                // per-thread halves stay below 2^16 (8 frames per chunk); the warp sums are taken per half
                const uint32_t ov = __reduce_add_sync(0xffffffffu, cnt[k] & 0xffffu);
                const uint32_t gs = __reduce_add_sync(0xffffffffu, cnt[k] >> 16);
                if (lane == 0) {
                    if (ov) atomicAdd(&s_ov[u0 + k], ov);
                    if (gs) atomicAdd(&s_gs[u0 + k], gs);
                }
            }
        }
    }
}

template <int UNROLL>
__device__ __forceinline__ void fscore_rows_acc(const smz_video_desc &d, int f0, int f1, const float *__restrict__ user,
                                                const uint32_t *vm, uint32_t *s_ov, uint32_t *s_gs) {
    // 128-bit loads need 16-byte aligned rows: user_off % 4 == 0 && user_ld % 4 == 0 (VideoBatch pads rows so)
    const bool vec = (((d.user_off | d.user_ld) & 3) == 0) && ((reinterpret_cast<uintptr_t>(user) & 15) == 0);
    if (vec) fscore_rows_acc_impl<UNROLL, true>(d, f0, f1, user, vm, s_ov, s_gs);
    else fscore_rows_acc_impl<1, false>(d, f0, f1, user, vm, s_ov, s_gs);
}

// utils/eval.py:151-164 for ONE video: float32 (numpy 2 / NEP 50 semantics) when no zero padding happened, float64 when
// the summary was shorter than n_frames (`padded`); one thread.
// overlap / gsum: the video's per-annotator counts (any address space); f: the video's slice of the F output.
// The float64 case of utils/eval.py:143-145: a machine summary SHORTER than n_frames is zero-padded with np.zeros (float64),
// which promotes it — overlap, machine_summary.sum(), precision, recall and F are float64 then, while gt_summary.sum() + 1e-8
// is still formed in float32.  F of annotator u, recomputed on demand (one thread, rare path).
struct PaddedF64Cursor {
    const int32_t *overlap, *gsum;
    double ms;
    __device__ __forceinline__ double at(int u) const {
        const double ov = (double)overlap[u];
        const double precision = __ddiv_rn(ov, ms);
        const double recall = __ddiv_rn(ov, (double)__fadd_rn((float)gsum[u], 1e-8f));
        if (precision == 0. && recall == 0.) return 0.;
        return __ddiv_rn(__dmul_rn(__dmul_rn(2., precision), recall), __dadd_rn(precision, recall));
    }
};

__device__ inline void fscore_final_video(int n_users, int msum, const int32_t *overlap, const int32_t *gsum, float *fv,
                                          double *avg_f, double *max_f, bool padded = false) {
    if (padded) {
        PaddedF64Cursor c{overlap, gsum, __dadd_rn((double)msum, 1e-8)};
        double mx = 0.;
        for (int u = 0; u < n_users; u++) {
            const double fs = c.at(u);
            fv[u] = (float)fs;
            mx = (u == 0) ? fs : fmax(mx, fs);
        }
        if (avg_f) *avg_f = n_users <= 0 ? 0. : __ddiv_rn(pw_sum<double>(c, 0, n_users), (double)n_users);
        if (max_f) *max_f = mx;
        return;
    }
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
