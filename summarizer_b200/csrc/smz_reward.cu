// DSN diversity-representativeness reward (models/dsn.py:185-236 compute_reward) for all episodes of one
// video at once.  The reference rebuilds two T x T matrices with two fp32 GEMMs for EVERY episode
// (dsn.py:215-216,226-228) although neither depends on the sampled actions; here the Gram matrix
// G = X.X^T is computed ONCE per video on the tensor cores and every episode is a masked reduction over it.
//
//   split_kernel    x (fp32) -> [hi | lo | hi] and [hi | hi | lo] bf16 rows (K = 3072) so that ONE bf16 GEMM
//                   returns hi.hi^T + lo.hi^T + hi.lo^T, i.e. the Gram matrix to ~2^-16 relative (the reward
//                   feeds (reward - baseline), a difference of nearly equal numbers), plus ||x_t||^2 in fp32.
//   gemm_kernel     G [T, T] fp32 (tcgen05).
//   reward_rows     one warp per frame t: for every episode e, sum_{j in picks_e} d'(t, j) (if t is picked) and
//                   min_{j in picks_e} ||x_t - x_j||^2, reading row t of G once for all episodes.
//   reward_final    one warp per episode: fixed-order float64 reduction -> 0.5 * (R_div + R_rep).
#include "smz_gemm.cuh"
#include "smz_rows.cuh"

namespace {

using smz::GemmEpilogue;
using smz::GemmProblem;
using smz::kFeat;
typedef __nv_bfloat16 bf16;
constexpr int MAX_EPISODES = 8;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmin(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256)
split_kernel(const float *__restrict__ x, int T, bf16 *__restrict__ a, bf16 *__restrict__ b, float *__restrict__ sq) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    float s = 0.f;
    for (int c = lane; c < kFeat; c += 32) {
        const float v = x[(size_t)t * kFeat + c];
        const bf16 hi = __float2bfloat16_rn(v);
        const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        bf16 *ar = a + (size_t)t * 3 * kFeat, *br = b + (size_t)t * 3 * kFeat;
        ar[c] = hi; ar[kFeat + c] = lo; ar[2 * kFeat + c] = hi;
        br[c] = hi; br[kFeat + c] = hi; br[2 * kFeat + c] = lo;
        s = fmaf(v, v, s);
    }
    s = wsum(s);
    if (lane == 0) sq[t] = s;
}

// actions: [E, T] bytes (0/1).  rowdiv / rowmin: [E, T].
__global__ void __launch_bounds__(256)
reward_rows_kernel(const float *__restrict__ G, int ldg, const float *__restrict__ sq, const uint8_t *__restrict__ actions,
                   int T, int E, int thre, int far_sim, float *__restrict__ rowdiv, float *__restrict__ rowmin) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    const float sqt = sq[t], inv_t = rsqrtf(sqt);
    float div[MAX_EPISODES], mn[MAX_EPISODES];
    uint32_t mine = 0u;
#pragma unroll
    for (int e = 0; e < MAX_EPISODES; e++) {
        div[e] = 0.f; mn[e] = INFINITY;
        if (e < E && actions[(size_t)e * T + t]) mine |= 1u << e;
    }
    for (int j = lane; j < T; j += 32) {
        uint32_t m = 0u;
#pragma unroll
        for (int e = 0; e < MAX_EPISODES; e++)
            if (e < E && actions[(size_t)e * T + j]) m |= 1u << e;
        if (m == 0u) continue;
        const float g = G[(size_t)t * ldg + j];
        const float sqj = sq[j];
        int dt = t - j; dt = dt < 0 ? -dt : dt;
        float dissim = 1.f - g * inv_t * rsqrtf(sqj);                 // 1 - cos(x_t, x_j), dsn.py:214-216
        if (!far_sim && dt > thre) dissim = 1.f;                      // dsn.py:218-222
        const float dist = sqt + sqj - 2.f * g;                      // dsn.py:226-228
#pragma unroll
        for (int e = 0; e < MAX_EPISODES; e++)
            if ((m >> e) & 1u) {
                if ((mine >> e) & 1u) div[e] += dissim;
                mn[e] = fminf(mn[e], dist);
            }
    }
#pragma unroll
    for (int e = 0; e < MAX_EPISODES; e++) {
        if (e >= E) break;
        const float d = wsum(div[e]), mm = wmin(mn[e]);
        if (lane == 0) { rowdiv[(size_t)e * T + t] = d; rowmin[(size_t)e * T + t] = mm; }
    }
}

__global__ void reward_final_kernel(const float *__restrict__ rowdiv, const float *__restrict__ rowmin,
                                    const uint8_t *__restrict__ actions, int T, int E, float *__restrict__ rewards) {
    const int e = blockIdx.x, lane = threadIdx.x;
    double sd = 0., sm = 0.;
    int picks = 0;
    for (int t = lane; t < T; t += 32) {
        sd += (double)rowdiv[(size_t)e * T + t];
        sm += (double)rowmin[(size_t)e * T + t];
        picks += actions[(size_t)e * T + t] ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sd += __shfl_xor_sync(0xffffffffu, sd, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        picks += __shfl_xor_sync(0xffffffffu, picks, o);
    }
    if (lane == 0) {
        float r = 0.f;                                                // no frame selected: zero reward (dsn.py:199-203)
        if (picks > 0) {
            const float r_div = picks > 1 ? (float)(sd / ((double)picks * (double)(picks - 1))) : 0.f;   // dsn.py:207-223
            const float r_rep = expf(-(float)(sm / (double)T));       // dsn.py:229-231
            r = (r_div + r_rep) * 0.5f;
        }
        rewards[e] = r;
    }
}

int64_t up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct RewardPlan { int64_t off_a, off_b, off_g, off_sq, off_rd, off_rm, total; int ldg; };

RewardPlan reward_plan(int T, int E) {
    RewardPlan p;
    p.ldg = (int)up(T, 8);
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o += up(bytes, 1024); return at; };
    p.off_a = take((int64_t)T * 3 * kFeat * 2);
    p.off_b = take((int64_t)T * 3 * kFeat * 2);
    p.off_g = take((int64_t)T * p.ldg * 4);
    p.off_sq = take((int64_t)T * 4);
    p.off_rd = take((int64_t)E * T * 4);
    p.off_rm = take((int64_t)E * T * 4);
    p.total = o;
    return p;
}

}  // namespace

extern "C" int smz_dsn_reward_workspace_bytes(int T, int n_episodes, int64_t *bytes) {
    SMZ_REQUIRE(bytes != nullptr && T > 0 && n_episodes > 0, "reward_workspace_bytes: bad argument");
    *bytes = reward_plan(T, n_episodes).total;
    return SMZ_OK;
}

extern "C" int smz_dsn_reward(const float *x, int T, const uint8_t *actions, int n_episodes, int temp_dist_thre,
                              int far_sim, float *rewards, void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(x && actions && rewards && ws && T > 0, "dsn_reward: NULL pointer");
    SMZ_REQUIRE(n_episodes >= 1 && n_episodes <= MAX_EPISODES, "dsn_reward: 1..%d episodes per call", MAX_EPISODES);
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    const RewardPlan pl = reward_plan(T, n_episodes);
    SMZ_REQUIRE(ws_bytes >= pl.total, "dsn_reward: work buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    bf16 *a = reinterpret_cast<bf16 *>(w + pl.off_a), *b = reinterpret_cast<bf16 *>(w + pl.off_b);
    float *G = reinterpret_cast<float *>(w + pl.off_g), *sq = reinterpret_cast<float *>(w + pl.off_sq);
    float *rd = reinterpret_cast<float *>(w + pl.off_rd), *rm = reinterpret_cast<float *>(w + pl.off_rm);
    split_kernel<<<(T + 7) / 8, 256, 0, st>>>(x, T, a, b, sq);
    SMZ_CUDA_CHECK(cudaGetLastError());
    GemmProblem g = {};
    g.M = T; g.N = T; g.K = 3 * kFeat; g.ldc = pl.ldg; g.tiles_n = (T + smz::GEMM_BN - 1) / smz::GEMM_BN;
    rc = smz::gemm_bf16_tn(a, T, 3 * kFeat, 3 * kFeat, b, T, 3 * kFeat, 3 * kFeat, nullptr, 1, smz::gemm_tiles(T, T), g,
                           GemmEpilogue{G, nullptr, nullptr, 1.f, smz::GEMM_OUT_F32}, st);
    if (rc != SMZ_OK) return rc;
    reward_rows_kernel<<<(T + 7) / 8, 256, 0, st>>>(G, pl.ldg, sq, actions, T, n_episodes, temp_dist_thre, far_sim, rd, rm);
    SMZ_CUDA_CHECK(cudaGetLastError());
    reward_final_kernel<<<n_episodes, 32, 0, st>>>(rd, rm, actions, T, n_episodes, rewards);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

// ---------------------------------------------------------------------------------------------------
// Episode sampling: dsn.py:112,125-126 — dist = Bernoulli(probs); actions = dist.sample(); dist.log_prob(actions)
// for every episode of a step in ONE launch: Philox4x32-10 draws (key = seed, counter = (frame quad, episode,
// call number)), action = u < p as torch's bernoulli, log-probability as torch's Bernoulli.log_prob (probabilities
// clamped to [eps, 1 - eps]), mean over the frames per episode.  The call number lives on the device next to the seed and is bumped by
// the last CTA to finish, so a captured step draws fresh episodes at every replay.
namespace {

using smz::philox4x32_10;

// torch.distributions.Bernoulli(probs).log_prob: probs are clamped to [eps, 1 - eps] (float32 eps = 2^-23) on their
// way to logits, and log_prob = -binary_cross_entropy_with_logits(logits, action) = log(p_c) resp. log(1 - p_c)
constexpr float kProbEps = 1.1920928955078125e-07f;
__device__ __forceinline__ float bern_logp(float p, bool a) {
    const float pc = fminf(fmaxf(p, kProbEps), 1.f - kProbEps);
    return a ? logf(pc) : log1pf(-pc);
}

__global__ void __launch_bounds__(256) bernoulli_logprob_kernel(const float *__restrict__ probs, int T, int E,
                                                                  unsigned long long *__restrict__ state,
                                                                  const uint8_t *__restrict__ given,
                                                                  uint8_t *__restrict__ actions, float *__restrict__ logp,
                                                                  unsigned int *__restrict__ ticket) {
    const int e = blockIdx.x;
    const unsigned long long seed = state[0], call = state[1];
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    float acc = 0.f;
    for (int q = threadIdx.x; q * 4 < T; q += blockDim.x) {
        uint4 r = make_uint4(0, 0, 0, 0);
        if (given == nullptr) r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)e, (uint32_t)call, (uint32_t)(call >> 32)), key);
        const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int t = q * 4 + i;
            if (t >= T) break;
            const float p = probs[t];
            const bool a = given != nullptr ? given[(int64_t)e * T + t] != 0 : (float)(rv[i] >> 8) * 0x1p-24f < p;
            actions[(int64_t)e * T + t] = a ? 1 : 0;
            acc += bern_logp(p, a);
        }
    }
    __shared__ float s_part[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; i++) s += s_part[i];                  // fixed order: the same draws give the same bits
        logp[e] = s / (float)T;
        __threadfence();
        if (atomicAdd(ticket, 1u) == (unsigned)E - 1u) {             // every CTA has read the call number by now
            *ticket = 0u;
            if (given == nullptr) state[1] = call + 1ull;
        }
    }
}

// d mean_t log_prob(a_e) / d p_t, chained with the episodes' upstream gradients g[e]: (a - p) / (p (1 - p)) inside
// the clamp range, 0 outside
__global__ void bernoulli_logprob_bwd_kernel(const float *__restrict__ probs, const uint8_t *__restrict__ actions,
                                             const float *__restrict__ g, int T, int E, float *__restrict__ dprobs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float p = probs[t];
    if (!(p >= kProbEps && p <= 1.f - kProbEps)) { dprobs[t] = 0.f; return; }     // clamped: no gradient, as torch
    float acc = 0.f;
    for (int e = 0; e < E; e++) acc += g[e] * ((actions[(int64_t)e * T + t] ? 1.f : 0.f) - p);
    dprobs[t] = acc / (p * (1.f - p) * (float)T);
}

}  // namespace

extern "C" int smz_bernoulli_logprob(const float *probs, int T, int n_episodes, uint64_t *state, const uint8_t *given,
                                     uint8_t *actions, float *logp_mean, void *stream) {
    SMZ_REQUIRE(probs && state && actions && logp_mean, "bernoulli_logprob: NULL pointer");
    SMZ_REQUIRE(T > 0 && n_episodes >= 1 && n_episodes <= 1024, "bernoulli_logprob: T > 0 and 1..1024 episodes");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "state words");
    unsigned long long *s = reinterpret_cast<unsigned long long *>(state);
    bernoulli_logprob_kernel<<<n_episodes, 256, 0, (cudaStream_t)stream>>>(probs, T, n_episodes, s, given, actions, logp_mean,
                                                                            reinterpret_cast<unsigned int *>(s + 2));
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_bernoulli_logprob_backward(const float *probs, const uint8_t *actions, const float *dlogp, int T,
                                              int n_episodes, float *dprobs, void *stream) {
    SMZ_REQUIRE(probs && actions && dlogp && dprobs, "bernoulli_logprob_backward: NULL pointer");
    SMZ_REQUIRE(T > 0 && n_episodes >= 1, "bernoulli_logprob_backward: T > 0 and >= 1 episode");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    bernoulli_logprob_bwd_kernel<<<(T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(probs, actions, dlogp, T, n_episodes, dprobs);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
