// Row-wise kernels of the scorers: softmax with VASNet's masks (vasnet.py:121-130), LayerNorm
// (vasnet.py:137,143), regressor head k2 + sigmoid (vasnet.py:144-145), fp32 -> bf16 conversion.
// One warp per row, 128-bit accesses, fp32 arithmetic.  All of them are L2/HBM-bandwidth bound.
#include "smz_rows.cuh"

#include <math.h>

namespace {

using smz::GemmProblem;
using smz::kFeat;

constexpr int ROW_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void cvt_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = *reinterpret_cast<const float4 *>(x + i);
        const float4 b = *reinterpret_cast<const float4 *>(x + i + 4);
        *reinterpret_cast<uint4 *>(y + i) = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
    } else {
        for (int64_t j = i; j < n; j++) y[j] = __float2bfloat16_rn(x[j]);
    }
}

// vasnet.py:121-127 — masks are applied AFTER scaling: diagonal (ignore_self), then the aperture
// band; inside the band an entry whose square is 0 is masked too (tril(e)*triu(e) == 0 quirk).
__device__ __forceinline__ float mask_logit(float e, int i, int j, int aperture, int ignore_self) {
    if (ignore_self && j == i) e = -INFINITY;
    if (aperture >= 0) {
        int d = i - j;
        d = d < 0 ? -d : d;
        if (d > aperture || e * e == 0.f) e = -INFINITY;
    }
    return e;
}

// One warp per query row.  Fast path (row fits 16 float4 per lane, no leading pad, no dropout): the
// logits are read ONCE into registers (16 independent 128-bit loads in flight per lane), max / exp /
// sum / normalise happen there, P is written once.  Other rows take the three-pass path.
__global__ void __launch_bounds__(ROW_WARPS * 32)
softmax_kernel(const GemmProblem *__restrict__ probs, int n_probs, int total_rows, const float *__restrict__ S,
               __nv_bfloat16 *__restrict__ alpha, __nv_bfloat16 *__restrict__ P, const uint8_t *__restrict__ drop,
               const int64_t *__restrict__ drop_off, int aperture, int ignore_self) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (r >= total_rows) return;
    int lo = 0, hi = n_probs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (probs[mid].c_off / probs[mid].ldc <= r) lo = mid; else hi = mid - 1;
    }
    const GemmProblem g = probs[lo];
    const int T = g.M, ld = g.ldc;
    const int i = r - (int)(g.c_off / ld);
    const int64_t off = g.c_off + (int64_t)i * ld;
    const float *s = S + off;
    const bool plain = aperture < 0 && !ignore_self;
    const int lead = g.pad;
    const int W64 = (lead + T + 63) & ~63;

    if (T <= 2048 && lead == 0 && drop == nullptr) {
        constexpr int NV = 16;
        float v[NV][4];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int j = lane * 4 + 128 * k;
            if (j < T) {
                const float4 e = *reinterpret_cast<const float4 *>(s + j);
                v[k][0] = e.x; v[k][1] = e.y; v[k][2] = e.z; v[k][3] = e.w;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (j + t >= T) v[k][t] = -INFINITY;
                    else if (!plain) v[k][t] = mask_logit(v[k][t], i, j + t, aperture, ignore_self);
                    m = fmaxf(m, v[k][t]);
                }
            }
        }
        m = warp_max(m);
        float l = 0.f;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int j = lane * 4 + 128 * k;
            if (j < T) {
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    v[k][t] = (j + t < T) ? __expf(v[k][t] - m) : 0.f;   // a fully masked row gives NaN, as torch does
                    l += v[k][t];
                }
            }
        }
        l = warp_sum(l);
        const float inv = 1.f / l;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int j = lane * 4 + 128 * k;
            if (j < W64) {
                uint2 o = make_uint2(0u, 0u);
                if (j < T) o = make_uint2(pack2(v[k][0] * inv, v[k][1] * inv), pack2(v[k][2] * inv, v[k][3] * inv));
                *reinterpret_cast<uint2 *>(P + off + j) = o;
            }
        }
        return;
    }

    float m = -INFINITY;
    for (int j = lane * 4; j < T; j += 128) {
        const float4 e = *reinterpret_cast<const float4 *>(s + j);
        const float v[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (j + t < T) m = fmaxf(m, plain ? v[t] : mask_logit(v[t], i, j + t, aperture, ignore_self));
    }
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane * 4; j < T; j += 128) {
        const float4 e = *reinterpret_cast<const float4 *>(s + j);
        const float v[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (j + t < T) l += __expf((plain ? v[t] : mask_logit(v[t], i, j + t, aperture, ignore_self)) - m);
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    const uint8_t *keep = drop != nullptr ? drop + drop_off[lo] + (int64_t)i * T : nullptr;
    // P / alpha rows: `lead` zero columns, the T probabilities, zeros up to the next multiple of 64
    for (int jo = lane * 4; jo < W64; jo += 128) {
        float a[4], p[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int j = jo + t - lead;
            a[t] = 0.f;
            if (j >= 0 && j < T) a[t] = __expf((plain ? s[j] : mask_logit(s[j], i, j, aperture, ignore_self)) - m) * inv;
            p[t] = a[t];
            if (keep != nullptr && j >= 0 && j < T) p[t] = keep[j] ? 2.f * a[t] : 0.f;   // nn.Dropout(0.5), vasnet.py:130
        }
        if (alpha != nullptr && alpha != P)
            *reinterpret_cast<uint2 *>(alpha + off + jo) = make_uint2(pack2(a[0], a[1]), pack2(a[2], a[3]));
        *reinterpret_cast<uint2 *>(P + off + jo) = make_uint2(pack2(p[0], p[1]), pack2(p[2], p[3]));
    }
}

// torch.nn.LayerNorm(1024, eps): biased variance, eps inside the square root.
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_kernel(const float *__restrict__ y, const uint8_t *__restrict__ keep, const float *__restrict__ g,
                 const float *__restrict__ b, float eps, int rows, __nv_bfloat16 *__restrict__ yn,
                 float *__restrict__ mean, float *__restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float *row = y + (int64_t)r * kFeat;
    float x[32];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = lane * 4 + 128 * k;
        const float4 v = *reinterpret_cast<const float4 *>(row + c);
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        if (keep != nullptr) {
            const uchar4 kp = *reinterpret_cast<const uchar4 *>(keep + (int64_t)r * kFeat + c);
            x[4 * k] = kp.x ? 2.f * x[4 * k] : 0.f; x[4 * k + 1] = kp.y ? 2.f * x[4 * k + 1] : 0.f;
            x[4 * k + 2] = kp.z ? 2.f * x[4 * k + 2] : 0.f; x[4 * k + 3] = kp.w ? 2.f * x[4 * k + 3] : 0.f;
        }
        sum += (x[4 * k] + x[4 * k + 1]) + (x[4 * k + 2] + x[4 * k + 3]);
    }
    const float mu = warp_sum(sum) * (1.f / kFeat);
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++) { const float d = x[k] - mu; var += d * d; }
    const float rs = rsqrtf(warp_sum(var) * (1.f / kFeat) + eps);
    if (lane == 0) {
        if (mean != nullptr) mean[r] = mu;
        if (rstd != nullptr) rstd[r] = rs;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = lane * 4 + 128 * k;
        const float4 gg = *reinterpret_cast<const float4 *>(g + c);
        const float4 bb = *reinterpret_cast<const float4 *>(b + c);
        const float o0 = (x[4 * k] - mu) * rs * gg.x + bb.x, o1 = (x[4 * k + 1] - mu) * rs * gg.y + bb.y;
        const float o2 = (x[4 * k + 2] - mu) * rs * gg.z + bb.z, o3 = (x[4 * k + 3] - mu) * rs * gg.w + bb.w;
        *reinterpret_cast<uint2 *>(yn + (int64_t)r * kFeat + c) = make_uint2(pack2(o0, o1), pack2(o2, o3));
    }
}

__global__ void __launch_bounds__(ROW_WARPS * 32)
head_kernel(const float *__restrict__ h, const uint8_t *__restrict__ keep, const float *__restrict__ g,
            const float *__restrict__ b, float eps, const float *__restrict__ w2, const float *__restrict__ b2,
            int rows, float *__restrict__ scores, float *__restrict__ mean, float *__restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float *row = h + (int64_t)r * kFeat;
    float x[32];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = lane * 4 + 128 * k;
        const float4 v = *reinterpret_cast<const float4 *>(row + c);
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        if (keep != nullptr) {
            const uchar4 kp = *reinterpret_cast<const uchar4 *>(keep + (int64_t)r * kFeat + c);
            x[4 * k] = kp.x ? 2.f * x[4 * k] : 0.f; x[4 * k + 1] = kp.y ? 2.f * x[4 * k + 1] : 0.f;
            x[4 * k + 2] = kp.z ? 2.f * x[4 * k + 2] : 0.f; x[4 * k + 3] = kp.w ? 2.f * x[4 * k + 3] : 0.f;
        }
        sum += (x[4 * k] + x[4 * k + 1]) + (x[4 * k + 2] + x[4 * k + 3]);
    }
    const float mu = warp_sum(sum) * (1.f / kFeat);
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++) { const float d = x[k] - mu; var += d * d; }
    const float rs = rsqrtf(warp_sum(var) * (1.f / kFeat) + eps);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = lane * 4 + 128 * k;
        const float4 gg = *reinterpret_cast<const float4 *>(g + c);
        const float4 bb = *reinterpret_cast<const float4 *>(b + c);
        const float4 ww = *reinterpret_cast<const float4 *>(w2 + c);
        dot += ((x[4 * k] - mu) * rs * gg.x + bb.x) * ww.x + ((x[4 * k + 1] - mu) * rs * gg.y + bb.y) * ww.y +
               ((x[4 * k + 2] - mu) * rs * gg.z + bb.z) * ww.z + ((x[4 * k + 3] - mu) * rs * gg.w + bb.w) * ww.w;
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        const float z = dot + __ldg(b2);
        scores[r] = 1.f / (1.f + __expf(-z));
        if (mean != nullptr) mean[r] = mu;
        if (rstd != nullptr) rstd[r] = rs;
    }
}

}  // namespace

namespace smz {

int launch_cvt_bf16(const float *x, __nv_bfloat16 *y, int64_t n, cudaStream_t st) {
    if (n <= 0) return SMZ_OK;
    const int64_t blocks = (n + 8 * 256 - 1) / (8 * 256);
    cvt_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, n);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

int launch_softmax(const GemmProblem *d_probs, int n_probs, int total_rows, const float *S, __nv_bfloat16 *alpha,
                   __nv_bfloat16 *P, const uint8_t *drop, const int64_t *d_drop_off, int aperture, int ignore_self,
                   cudaStream_t st) {
    if (total_rows <= 0) return SMZ_OK;
    softmax_kernel<<<(total_rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, st>>>(
        d_probs, n_probs, total_rows, S, alpha, P, drop, d_drop_off, aperture, ignore_self);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

int launch_layernorm(const float *y, const uint8_t *keep, const float *g, const float *b, float eps, int rows,
                     __nv_bfloat16 *yn, float *mean, float *rstd, cudaStream_t st) {
    if (rows <= 0) return SMZ_OK;
    layernorm_kernel<<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, st>>>(y, keep, g, b, eps, rows, yn, mean, rstd);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

int launch_head(const float *h, const uint8_t *keep, const float *g, const float *b, float eps,
                const float *w2, const float *b2, int rows, float *scores, float *mean, float *rstd,
                cudaStream_t st) {
    if (rows <= 0) return SMZ_OK;
    head_kernel<<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, st>>>(h, keep, g, b, eps, w2, b2, rows, scores,
                                                                              mean, rstd);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

}  // namespace smz
