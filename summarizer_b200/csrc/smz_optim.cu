// Optimizer step of the training loops (models/vasnet.py:160-161,211-212, models/dsn.py:100,147-149,
// models/sumgan.py:268-275,433-436): torch.optim.Adam with the L2 term in the gradient, and
// torch.nn.utils.clip_grad_norm_(parameters, 5.0), over ALL tensors of a parameter list in a few launches.
//
//   smz_grad_sqnorm   sum of squares of every gradient -> one float32 (two deterministic stages: per-CTA partial sums in a
//                     fixed order, then one CTA adds them in a fixed order — the clip coefficient is bit-stable run to run)
//   smz_clip_grads    g *= min(1, max_norm / (sqrt(sqnorm) + 1e-6))          (clip_grad_norm_'s coefficient, in place: the
//                     reference's SumGAN trainer relies on the clipped values staying in .grad, sumgan.py:433-436)
//   smz_adam_step     one Adam update of every tensor; the step counter lives on the device (CUDA-graph replays advance it)
//
// HBM-bound element-wise work: 16 bytes read + 12 written per parameter for Adam.  Tensors are passed BY VALUE as a table in
// the kernel parameters (<= SMZ_OPTIM_MAX_TENSORS per launch), so nothing captured in a CUDA graph points at host memory.
#include <math.h>

#include "smz_common.cuh"

namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = OPT_THREADS * 4 * 4;       // elements per CTA trip: 4 float4 per thread

struct TensorTable {
    float *param[SMZ_OPTIM_MAX_TENSORS];
    float *grad[SMZ_OPTIM_MAX_TENSORS];
    float *exp_avg[SMZ_OPTIM_MAX_TENSORS];
    float *exp_avg_sq[SMZ_OPTIM_MAX_TENSORS];
    float *step[SMZ_OPTIM_MAX_TENSORS];
    long long n[SMZ_OPTIM_MAX_TENSORS];
    int chunk0[SMZ_OPTIM_MAX_TENSORS + 1];           // prefix sums of ceil(n / OPT_CHUNK): CTA -> (tensor, chunk)
    int count;
};

__device__ __forceinline__ int find_tensor(const TensorTable &t, int cta) {
    int lo = 0, hi = t.count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.chunk0[mid] <= cta) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// sum over the CTA in a fixed order: warp shuffles (fixed tree), then warp 0 adds the warp sums in index order
__device__ __forceinline__ float cta_sum_ordered(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int k = 0; k < OPT_THREADS / 32; k++) s += red[k];
    return s;          // valid in thread 0
}

__global__ void __launch_bounds__(OPT_THREADS) sqnorm_partial_kernel(const TensorTable t, float *__restrict__ partial) {
    smz::pdl_trigger();
    smz::pdl_wait();
    __shared__ float red[OPT_THREADS / 32];
    const int ti = find_tensor(t, blockIdx.x);
    const long long base = (long long)(blockIdx.x - t.chunk0[ti]) * OPT_CHUNK;
    const long long n = t.n[ti];
    const float *g = t.grad[ti];
    float acc = 0.f;
    if (g != nullptr) {
        const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long i = base + ((long long)k * OPT_THREADS + threadIdx.x) * 4;
            if (vec && i + 4 <= n) {
                const float4 x = *reinterpret_cast<const float4 *>(g + i);
                acc += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
            } else {
                for (long long j = i; j < n && j < i + 4; j++) acc += g[j] * g[j];
            }
        }
    }
    const float s = cta_sum_ordered(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] (+)= sum of partial[0..n) in a fixed order (one CTA); accumulate != 0 adds to the value already there
__global__ void __launch_bounds__(OPT_THREADS) sqnorm_final_kernel(const float *__restrict__ partial, int n, float *__restrict__ out,
                                                                    int accumulate) {
    smz::pdl_trigger();
    smz::pdl_wait();
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += OPT_THREADS) acc += partial[i];
    const float s = cta_sum_ordered(acc, red);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s;
}

__global__ void __launch_bounds__(OPT_THREADS) clip_kernel(const TensorTable t, const float *__restrict__ sqnorm, float max_norm) {
    smz::pdl_trigger();
    smz::pdl_wait();
    const float coef = fminf(max_norm / (sqrtf(__ldg(sqnorm)) + 1e-6f), 1.f);        // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
    const int ti = find_tensor(t, blockIdx.x);
    const long long base = (long long)(blockIdx.x - t.chunk0[ti]) * OPT_CHUNK;
    const long long n = t.n[ti];
    float *g = t.grad[ti];
    if (g == nullptr) return;
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long i = base + ((long long)k * OPT_THREADS + threadIdx.x) * 4;
        if (vec && i + 4 <= n) {
            float4 x = *reinterpret_cast<float4 *>(g + i);
            x.x *= coef; x.y *= coef; x.z *= coef; x.w *= coef;
            *reinterpret_cast<float4 *>(g + i) = x;
        } else {
            for (long long j = i; j < n && j < i + 4; j++) g[j] *= coef;
        }
    }
}

// torch.optim.Adam (amsgrad = False, maximize = False), the arithmetic of its fused CUDA implementation:
//   g += wd * p;  m = lerp(m, g, 1 - b1);  v = b2 * v + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Every tensor has its own step counter (device float, number of COMPLETED updates of that tensor — torch counts per
// parameter too: one that had no gradient in some step lags behind); the update reads it, bump_step_kernel behind the
// update increments it (stream order), so CUDA-graph replays advance t without host involvement.
__global__ void __launch_bounds__(OPT_THREADS) adam_kernel(const TensorTable t, float lr, float beta1, float beta2, float omb1,
                                                            float omb2, float eps, float weight_decay) {
    smz::pdl_trigger();
    smz::pdl_wait();
    const int ti = find_tensor(t, blockIdx.x);
    const float tstep = __ldg(t.step[ti]) + 1.f;
    const float bc1 = 1.f - powf(beta1, tstep);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, tstep));
    const float step_size = lr / bc1;
    const long long base = (long long)(blockIdx.x - t.chunk0[ti]) * OPT_CHUNK;
    const long long n = t.n[ti];
    float *p = t.param[ti], *m = t.exp_avg[ti], *v = t.exp_avg_sq[ti];
    const float *g = t.grad[ti];
    if (g == nullptr) return;          // a parameter without gradient is skipped, as torch does
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {
        gg = fmaf(weight_decay, pp, gg);
        mm = mm + (gg - mm) * omb1;                 // omb = 1 - beta, formed in double on the host (1 - 0.999f is 1.3e-5 off 0.001)
        vv = beta2 * vv + omb2 * gg * gg;
        pp -= step_size * mm / (sqrtf(vv) / bc2_sqrt + eps);
    };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long i = base + ((long long)k * OPT_THREADS + threadIdx.x) * 4;
        if (vec && i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4 *>(p + i), mm = *reinterpret_cast<float4 *>(m + i), vv = *reinterpret_cast<float4 *>(v + i);
            const float4 gg = *reinterpret_cast<const float4 *>(g + i);
            upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
            *reinterpret_cast<float4 *>(p + i) = pp; *reinterpret_cast<float4 *>(m + i) = mm; *reinterpret_cast<float4 *>(v + i) = vv;
        } else {
            for (long long j = i; j < n && j < i + 4; j++) upd(p[j], g[j], m[j], v[j]);
        }
    }
}

__global__ void bump_step_kernel(const TensorTable t) {
    smz::pdl_trigger();
    smz::pdl_wait();
    const int i = threadIdx.x;
    if (i < t.count && t.grad[i] != nullptr) t.step[i][0] += 1.f;
}

// host: tensors [first, first + count) of the caller's array -> kernel-parameter table; returns the number of CTAs
int fill_table(const smz_optim_tensor *tensors, int first, int count, bool need_state, TensorTable *t) {
    int chunks = 0;
    t->count = count;
    for (int i = 0; i < count; i++) {
        const smz_optim_tensor &s = tensors[first + i];
        t->param[i] = s.param; t->grad[i] = s.grad; t->exp_avg[i] = s.exp_avg; t->exp_avg_sq[i] = s.exp_avg_sq; t->step[i] = s.step;
        t->n[i] = s.n;
        t->chunk0[i] = chunks;
        chunks += (int)((s.n + OPT_CHUNK - 1) / OPT_CHUNK);
        (void)need_state;
    }
    t->chunk0[count] = chunks;
    return chunks;
}

int check_tensors(const smz_optim_tensor *tensors, int n_tensors, bool need_param, bool need_state) {
    SMZ_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors != nullptr), "optim: bad tensor list");
    for (int i = 0; i < n_tensors; i++) {
        SMZ_REQUIRE(tensors[i].n >= 0 && tensors[i].n < (1ll << 40), "optim: tensor %d has a bad size", i);
        SMZ_REQUIRE(!need_param || tensors[i].param != nullptr, "optim: tensor %d has no parameter pointer", i);
        SMZ_REQUIRE(!need_state || tensors[i].grad == nullptr || (tensors[i].exp_avg && tensors[i].exp_avg_sq && tensors[i].step),
                    "optim: tensor %d has a gradient but no Adam state", i);
    }
    return SMZ_OK;
}

}  // namespace

extern "C" int smz_grad_sqnorm_workspace_floats(const smz_optim_tensor *tensors, int n_tensors, int64_t *floats) {
    SMZ_REQUIRE(floats != nullptr, "floats is NULL");
    int rc = check_tensors(tensors, n_tensors, false, false);
    if (rc != SMZ_OK) return rc;
    int64_t most = 1;
    for (int first = 0; first < n_tensors; first += SMZ_OPTIM_MAX_TENSORS) {
        const int count = n_tensors - first < SMZ_OPTIM_MAX_TENSORS ? n_tensors - first : SMZ_OPTIM_MAX_TENSORS;
        int64_t chunks = 0;
        for (int i = 0; i < count; i++) chunks += (tensors[first + i].n + OPT_CHUNK - 1) / OPT_CHUNK;
        if (chunks > most) most = chunks;
    }
    *floats = most;
    return SMZ_OK;
}

extern "C" int smz_grad_sqnorm(const smz_optim_tensor *tensors, int n_tensors, float *sqnorm, float *ws, int64_t ws_floats,
                               void *stream) {
    SMZ_REQUIRE(sqnorm != nullptr, "sqnorm is NULL");
    int rc = check_tensors(tensors, n_tensors, false, false);
    if (rc != SMZ_OK) return rc;
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool any = false;
    for (int first = 0; first < n_tensors; first += SMZ_OPTIM_MAX_TENSORS) {
        const int count = n_tensors - first < SMZ_OPTIM_MAX_TENSORS ? n_tensors - first : SMZ_OPTIM_MAX_TENSORS;
        TensorTable t;
        const int chunks = fill_table(tensors, first, count, false, &t);
        if (chunks == 0) continue;
        SMZ_REQUIRE(ws != nullptr && ws_floats >= chunks, "grad_sqnorm: work buffer too small (%lld < %d floats)", (long long)ws_floats, chunks);
        SMZ_CUDA_CHECK(smz::launch_pdl(sqnorm_partial_kernel, dim3(chunks), dim3(OPT_THREADS), 0, st, t, ws));
        SMZ_CUDA_CHECK(smz::launch_pdl(sqnorm_final_kernel, dim3(1), dim3(OPT_THREADS), 0, st, ws, chunks, sqnorm, any ? 1 : 0));
        any = true;
    }
    if (!any) SMZ_CUDA_CHECK(cudaMemsetAsync(sqnorm, 0, sizeof(float), st));
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_clip_grads(const smz_optim_tensor *tensors, int n_tensors, const float *sqnorm, float max_norm, void *stream) {
    SMZ_REQUIRE(sqnorm != nullptr && max_norm > 0.f, "clip_grads: bad argument");
    int rc = check_tensors(tensors, n_tensors, false, false);
    if (rc != SMZ_OK) return rc;
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int first = 0; first < n_tensors; first += SMZ_OPTIM_MAX_TENSORS) {
        const int count = n_tensors - first < SMZ_OPTIM_MAX_TENSORS ? n_tensors - first : SMZ_OPTIM_MAX_TENSORS;
        TensorTable t;
        const int chunks = fill_table(tensors, first, count, false, &t);
        if (chunks == 0) continue;
        SMZ_CUDA_CHECK(smz::launch_pdl(clip_kernel, dim3(chunks), dim3(OPT_THREADS), 0, st, t, sqnorm, max_norm));
    }
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_adam_step(const smz_optim_tensor *tensors, int n_tensors, double lr, double beta1, double beta2, double eps,
                             double weight_decay, void *stream) {
    SMZ_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0. && weight_decay >= 0.,
                "adam_step: bad hyper-parameter");
    int rc = check_tensors(tensors, n_tensors, true, true);
    if (rc != SMZ_OK) return rc;
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int first = 0; first < n_tensors; first += SMZ_OPTIM_MAX_TENSORS) {
        const int count = n_tensors - first < SMZ_OPTIM_MAX_TENSORS ? n_tensors - first : SMZ_OPTIM_MAX_TENSORS;
        TensorTable t;
        const int chunks = fill_table(tensors, first, count, true, &t);
        if (chunks == 0) continue;
        SMZ_CUDA_CHECK(smz::launch_pdl(adam_kernel, dim3(chunks), dim3(OPT_THREADS), 0, st, t, (float)lr, (float)beta1, (float)beta2,
                                       (float)(1. - beta1), (float)(1. - beta2), (float)eps, (float)weight_decay));
        SMZ_CUDA_CHECK(smz::launch_pdl(bump_step_kernel, dim3(1), dim3(SMZ_OPTIM_MAX_TENSORS), 0, st, t));      // behind the update that read the counters
    }
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

// ---------------------------------------------------------------------------------------------------
// torch.nn.MSELoss() of the supervised trainers (vasnet.py:199,208): loss = mean((s - t)^2) AND its gradient
// d loss / d s = 2 (s - t) / n in one launch — torch spends six tiny kernels on the pair per step (square, mean, two fills,
// mse_backward, a fill) in a step that is launch-bound.  One CTA, fixed summation order (bit-stable).
namespace {
__global__ void __launch_bounds__(1024) mse_loss_kernel(const float *__restrict__ s, const float *__restrict__ t, long long n,
                                                         float *__restrict__ loss, float *__restrict__ ds) {
    smz::pdl_trigger();
    smz::pdl_wait();
    __shared__ float red[32];
    const float inv = 1.f / (float)n;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = s[i] - t[i];
        acc = fmaf(d, d, acc);
        if (ds != nullptr) ds[i] = 2.f * d * inv;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) loss[0] = v * inv;
    }
}
}  // namespace

extern "C" int smz_mse_loss(const float *scores, const float *target, int64_t n, float *loss, float *dscores, void *stream) {
    SMZ_REQUIRE(scores && target && loss && n > 0, "mse_loss: bad argument");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    SMZ_CUDA_CHECK(smz::launch_pdl(mse_loss_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, scores, target, (long long)n, loss,
                                   dscores));
    return SMZ_OK;
}
