// Rank correlation between machine and annotator frame scores: utils/eval.py:49-72 evaluate_scores
// (scipy.stats.rankdata(-x) with average ties, then Spearman = Pearson of the ranks, or Kendall tau-b).
// This is what Trainer.test uses to pick the best weights (models/__init__.py:60-86) and the largest
// CPU cost of the reference's evaluation.
//
//   rank_kernel      one CTA per (row, video): the row (n_frames float32) is bitonic-sorted in shared
//                    memory (<= 32768 keys = 128 KB), then every element finds its tie group by two binary
//                    searches:  rank = #greater + (#equal + 1) / 2   (descending, average ties).
//   spearman_kernel  one CTA per (annotator, video): float64 centred sums of the two rank rows.
//   kendall_kernel   one CTA per (annotator, video): O(n^2) concordant / discordant / tie pair counts (tau-b).
//   corr_mean_kernel np.mean over annotators in numpy's pairwise order.
#include "smz_common.cuh"
#include "smz_eval_dev.cuh"

#include <math.h>

namespace {

constexpr int RANK_THREADS = 1024;
constexpr int RANK_MAX_N = 32768;
constexpr int CORR_THREADS = 256;

__global__ void __launch_bounds__(RANK_THREADS)
rank_kernel(const smz_corr_desc *__restrict__ desc, const float *__restrict__ machine, const float *__restrict__ user,
            float *__restrict__ ranks, int npow2_max) {
    extern __shared__ float s_key[];
    const smz_corr_desc d = desc[blockIdx.y];
    const int row = blockIdx.x;                 // 0 = machine scores, 1.. = annotators
    if (row > d.n_users) return;
    const int n = d.n_frames;
    const float *src = row == 0 ? machine + d.m_off : user + d.u_off + (int64_t)(row - 1) * d.u_ld;
    float *dst = ranks + d.rank_off + (int64_t)row * n;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += RANK_THREADS) s_key[i] = i < n ? src[i] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (np2 >> 1); t += RANK_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // lower index of the pair
                const int p = i | j;
                const bool up = (i & k) == 0;
                const float a = s_key[i], b = s_key[p];
                if ((a > b) == up) { s_key[i] = b; s_key[p] = a; }
            }
            __syncthreads();
        }
    }
    // ascending sorted keys in s_key[0..n): rank of -x with average ties
    for (int i = threadIdx.x; i < n; i += RANK_THREADS) {
        const float v = src[i];
        int lo = 0, hi = n;                      // first index with key >= v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_key[mid] < v) lo = mid + 1; else hi = mid; }
        const int lb = lo;
        hi = n;                                  // first index with key > v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_key[mid] <= v) lo = mid + 1; else hi = mid; }
        const int greater = n - lo, equal = lo - lb;
        dst[i] = (float)greater + 0.5f * (float)(equal + 1);
    }
}

__device__ __forceinline__ double block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int k = 0; k < CORR_THREADS / 32; k++) t += s_red[k];
    return t;
}

__global__ void __launch_bounds__(CORR_THREADS)
spearman_kernel(const smz_corr_desc *__restrict__ desc, const float *__restrict__ ranks, double *__restrict__ corr) {
    __shared__ double s_red[CORR_THREADS / 32];
    const smz_corr_desc d = desc[blockIdx.y];
    const int u = blockIdx.x;
    if (u >= d.n_users) return;
    const int n = d.n_frames;
    const float *ra = ranks + d.rank_off;
    const float *rb = ranks + d.rank_off + (int64_t)(u + 1) * n;
    const double mu = 0.5 * (double)(n + 1);     // mean of any rank vector
    double sxy = 0., sxx = 0., syy = 0.;
    for (int i = threadIdx.x; i < n; i += CORR_THREADS) {
        const double a = (double)ra[i] - mu, b = (double)rb[i] - mu;
        sxy += a * b; sxx += a * a; syy += b * b;
    }
    sxy = block_sum(sxy, s_red); sxx = block_sum(sxx, s_red); syy = block_sum(syy, s_red);
    if (threadIdx.x == 0) corr[d.row0 + u] = sxy / sqrt(sxx * syy);   // constant input -> 0/0 = NaN, as scipy
}

// scipy.stats.kendalltau (tau-b): (P - Q) / sqrt((P + Q + T) * (P + Q + U)), T / U = pairs tied only in x / y
__global__ void __launch_bounds__(CORR_THREADS)
kendall_kernel(const smz_corr_desc *__restrict__ desc, const float *__restrict__ ranks, double *__restrict__ corr) {
    __shared__ double s_red[CORR_THREADS / 32];
    const smz_corr_desc d = desc[blockIdx.y];
    const int u = blockIdx.x;
    if (u >= d.n_users) return;
    const int n = d.n_frames;
    const float *ra = ranks + d.rank_off;
    const float *rb = ranks + d.rank_off + (int64_t)(u + 1) * n;
    long long con = 0, dis = 0, tx = 0, ty = 0;
    for (int i = threadIdx.x; i < n; i += CORR_THREADS) {
        const float ai = ra[i], bi = rb[i];
        for (int j = i + 1; j < n; j++) {
            const float da = ra[j] - ai, db = rb[j] - bi;
            const bool ea = da == 0.f, eb = db == 0.f;
            if (ea && eb) continue;
            if (ea) ++tx; else if (eb) ++ty; else if ((da > 0.f) == (db > 0.f)) ++con; else ++dis;
        }
    }
    const double P = block_sum((double)con, s_red), Q = block_sum((double)dis, s_red);
    const double T = block_sum((double)tx, s_red), U = block_sum((double)ty, s_red);
    if (threadIdx.x == 0) corr[d.row0 + u] = (P - Q) / sqrt((P + Q + T) * (P + Q + U));
}

__global__ void corr_mean_kernel(const smz_corr_desc *__restrict__ desc, int n_videos, const double *__restrict__ corr,
                                 double *__restrict__ avg) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_videos) return;
    const smz_corr_desc d = desc[v];
    struct Cur { const double *a; __device__ double at(int i) const { return a[i]; } } cur{corr + d.row0};
    avg[v] = d.n_users > 0 ? __ddiv_rn(smzdev::pw_sum<double>(cur, 0, d.n_users), (double)d.n_users) : NAN;
}

}  // namespace

extern "C" int smz_rank_correlation(const smz_corr_desc *desc, int n_videos, int max_n_frames, int max_n_users,
                                    const float *machine, const float *user, int metric, float *rank_ws,
                                    double *corr, double *corr_avg, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && n_videos <= 65535, "rank_correlation: 1..65535 videos per call");
    SMZ_REQUIRE(desc && machine && user && rank_ws && corr && corr_avg, "rank_correlation: NULL pointer");
    SMZ_REQUIRE(metric == SMZ_METRIC_SPEARMAN || metric == SMZ_METRIC_KENDALL, "unknown metric %d", metric);
    if (max_n_frames > RANK_MAX_N)
        return smz::fail(SMZ_ERR_UNSUPPORTED, "rank_correlation: %d frames per video exceed the %d-key shared-memory sort",
                         max_n_frames, RANK_MAX_N);
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int np2 = 32;
    while (np2 < max_n_frames) np2 <<= 1;
    const int smem = np2 * (int)sizeof(float);
    SMZ_CUDA_CHECK(cudaFuncSetAttribute((const void *)rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    rank_kernel<<<dim3(max_n_users + 1, n_videos), RANK_THREADS, smem, st>>>(desc, machine, user, rank_ws, np2);
    SMZ_CUDA_CHECK(cudaGetLastError());
    if (max_n_users > 0) {
        if (metric == SMZ_METRIC_SPEARMAN)
            spearman_kernel<<<dim3(max_n_users, n_videos), CORR_THREADS, 0, st>>>(desc, rank_ws, corr);
        else
            kendall_kernel<<<dim3(max_n_users, n_videos), CORR_THREADS, 0, st>>>(desc, rank_ws, corr);
        SMZ_CUDA_CHECK(cudaGetLastError());
    }
    corr_mean_kernel<<<(n_videos + 127) / 128, 128, 0, st>>>(desc, n_videos, corr, corr_avg);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
