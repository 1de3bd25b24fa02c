// Kernel temporal segmentation (KTS; Potapov, Douze, Harchaoui, Schmid, ECCV 2014): the change-point detector whose
// output the datasets carry as /change_points (datasets/README.md:24-27 "typically the result of KTS on the original
// video").  The reference only CONSUMES those change points — it ships no KTS code (SURVEY.md §8f NEXT-4) — so this
// follows the published algorithm (cpd_nonlin / cpd_auto of the authors' release); see oracle/kts_np.py.
//
//   gram_kernel      K = X X^T in float32 (64x64 tiles, 4x4 micro-tiles; exactly symmetric: both triangles sum d in the
//                    same order)
//   prefix kernels   K1 = prefix of diag(K), K2 = 2-D inclusive prefix of K, float64 (one thread per row / column)
//   kts_dp_kernel    row k of the DP  I[k,l] = min_t I[k-1,t] + J(t, l-1): one warp per l, lanes over t, first-minimum
//                    argmin; the scatter J(i,j) = K1[j+1]-K1[i] - (K2[j+1,j+1]+K2[i,i]-K2[j+1,i]-K2[i,j+1])/(j-i+1) is
//                    formed on the fly from the prefixes (no n x n cost matrix)
//   kts_pick_kernel  cpd_auto's penalised choice of the number of change points + the backtrack, on the device
#include "smz_common.cuh"

#include <math.h>

namespace {

constexpr double kBig = 1e101;

__global__ void __launch_bounds__(256) gram_kernel(const float *__restrict__ x, int n, int d, int ld, float *__restrict__ K) {
    __shared__ float sa[16][64 + 1], sb[16][64 + 1];
    const int bi = blockIdx.y * 64, bj = blockIdx.x * 64;
    if (bj < bi) return;                                     // upper triangle only; mirrored below
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, c = e & 15;
            sa[c][r] = (bi + r < n && k0 + c < d) ? x[(size_t)(bi + r) * ld + k0 + c] : 0.f;
            sb[c][r] = (bj + r < n && k0 + c < d) ? x[(size_t)(bj + r) * ld + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; c++) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { a[u] = sa[c][ty * 4 + u]; b[u] = sb[c][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int i = bi + ty * 4 + u, j = bj + tx * 4 + v;
            if (i < n && j < n) {
                // a(i).b(j) and a(j).b(i) accumulate the same products in the same order: mirror instead of recomputing
                K[(size_t)i * n + j] = acc[u][v];
                K[(size_t)j * n + i] = acc[u][v];
            }
        }
}

// K2[(a,b)] = sum_{i<a, j<b} K[i,j], (n+1) x (n+1), row / column 0 are zero.  Pass 1: down the columns (axis 0), pass 2:
// along the rows (axis 1) — numpy's np.cumsum(np.cumsum(K, 0), 1) in float64.
__global__ void prefix_cols_kernel(const float *__restrict__ K, int n, double *__restrict__ K2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;     // column of K
    if (j > n) return;
    const int N1 = n + 1;
    if (j == n) { for (int a = 0; a <= n; a++) K2[(size_t)a * N1] = 0.0; return; }
    double run = 0.0;
    K2[j + 1] = 0.0;
    for (int i = 0; i < n; i++) { run += (double)K[(size_t)i * n + j]; K2[(size_t)(i + 1) * N1 + j + 1] = run; }
}
__global__ void prefix_rows_kernel(int n, double *__restrict__ K2, double *__restrict__ D2) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;     // row of K2
    if (a > n) return;
    const int N1 = n + 1;
    double run = 0.0;
    double *row = K2 + (size_t)a * N1;
    for (int b = 1; b <= n; b++) { run += row[b]; row[b] = run; }
    D2[a] = row[a];
}
__global__ void prefix_diag_kernel(const float *__restrict__ K, int n, double *__restrict__ K1) {
    if (blockIdx.x || threadIdx.x) return;
    double run = 0.0;
    K1[0] = 0.0;
    for (int i = 0; i < n; i++) { run += (double)K[(size_t)i * n + i]; K1[i + 1] = run; }
}

__device__ __forceinline__ double scatter(const double *K1, const double *D2, const double *K2, int N1, int t, int l) {
    // J(i = t, j = l - 1)
    return (K1[l] - K1[t]) - (D2[l] + D2[t] - K2[(size_t)l * N1 + t] - K2[(size_t)t * N1 + l]) / (double)(l - t);
}

__global__ void kts_init_kernel(int n, int lmin, int lmax, const double *K1, const double *D2, const double *K2, double *I0) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n) return;
    I0[l] = (l >= lmin && l < lmax) ? scatter(K1, D2, K2, n + 1, 0, l) : kBig;   // I[0, lmin:lmax] = J[0, lmin-1:lmax-1]
}

__global__ void __launch_bounds__(256)
kts_dp_kernel(int k, int n, int lmin, int lmax, const double *__restrict__ K1, const double *__restrict__ D2,
              const double *__restrict__ K2, const double *__restrict__ Iprev, double *__restrict__ Icur,
              int32_t *__restrict__ prow) {
    const int l = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (l > n) return;
    if (l < (k + 1) * lmin) { if (lane == 0) { Icur[l] = kBig; prow[l] = 0; } return; }
    const int tmin = max(k * lmin, l - lmax), tmax = l - lmin + 1;
    double best = INFINITY;
    int arg = INT_MAX;
    const double *k2row = K2 + (size_t)l * (n + 1);
    const double k1l = K1[l], d2l = D2[l];
    for (int t = tmin + lane; t < tmax; t += 32) {
        // K2 is symmetric up to the rounding of the two summation orders; the row access K2[l,t] is the coalesced one
        const double c = (k1l - K1[t]) - (d2l + D2[t] - k2row[t] - K2[(size_t)t * (n + 1) + l]) / (double)(l - t) + Iprev[t];
        if (c < best) { best = c; arg = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }     // np.argmin: first minimum
    }
    if (lane == 0) { Icur[l] = tmin < tmax ? best : kBig; prow[l] = tmin < tmax ? arg : 0; }
}

__global__ void kts_score_kernel(int k, int n, const double *Icur, double *scores) {
    if (blockIdx.x || threadIdx.x) return;
    const double v = Icur[n];
    scores[k] = v > 1e99 ? INFINITY : v;
}

// cpd_auto: costs[k] = scores[k] / n + (vmax k / (2 N2)) (log(N2 / k) + 1), m_best = first argmin; then the backtrack.
__global__ void kts_pick_kernel(int n, int m, int auto_select, double vmax, int desc_rate, const double *scores,
                                const int32_t *p, int32_t *cps, int32_t *n_cps) {
    if (blockIdx.x || threadIdx.x) return;
    int best = m;
    if (auto_select) {
        const double N2 = (double)n * desc_rate;
        double bc = INFINITY;
        best = 0;
        for (int k = 0; k <= m; k++) {
            const double pen = k == 0 ? 0.0 : (vmax * k / (2.0 * N2)) * (log(N2 / k) + 1.0);
            const double c = scores[k] / (double)n + pen;
            if (c < bc) { bc = c; best = k; }
        }
    }
    int cur = n;
    for (int k = best; k >= 1; k--) { cps[k - 1] = p[(size_t)k * (n + 1) + cur]; cur = cps[k - 1]; }
    *n_cps = best;
}

int64_t up(int64_t x) { return (x + 255) / 256 * 256; }
struct KtsPlan { int64_t off_K, off_K2, off_K1, off_D2, off_I, off_p, total; };
KtsPlan kts_plan(int n, int m, bool own_gram) {
    KtsPlan p; int64_t o = 0;
    p.off_K = o; o += up(own_gram ? (int64_t)n * n * 4 : 0);
    p.off_K2 = o; o += up((int64_t)(n + 1) * (n + 1) * 8);
    p.off_K1 = o; o += up((int64_t)(n + 1) * 8);
    p.off_D2 = o; o += up((int64_t)(n + 1) * 8);
    p.off_I = o; o += up((int64_t)2 * (n + 1) * 8);
    p.off_p = o; o += up((int64_t)(m + 1) * (n + 1) * 4);
    p.total = o;
    return p;
}

}  // namespace

extern "C" int smz_kts_workspace_bytes(int n, int max_ncp, int from_features, int64_t *bytes) {
    SMZ_REQUIRE(bytes && n > 0 && max_ncp >= 0, "kts_workspace_bytes: bad argument");
    *bytes = kts_plan(n, max_ncp, from_features != 0).total;
    return SMZ_OK;
}

// K = X X^T (float32): features [n, d] (row stride ld) -> K [n, n]
extern "C" int smz_kts_gram(const float *features, int n, int d, int ld, float *K, void *stream) {
    SMZ_REQUIRE(features && K && n > 0 && d > 0 && ld >= d, "kts_gram: bad argument");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    const int tiles = (n + 63) / 64;
    gram_kernel<<<dim3(tiles, tiles), 256, 0, (cudaStream_t)stream>>>(features, n, d, ld, K);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

// Change points of the n x n kernel matrix K (device float32; NULL = compute it from `features`).  max_ncp = m of
// cpd_nonlin; auto_select != 0 = cpd_auto (penalty vmax, desc_rate).  Outputs (device): cps [max_ncp] ascending,
// n_cps [1], scores [max_ncp + 1] (the within-segment scatter for 0..max_ncp change points, +inf where infeasible).
extern "C" int smz_kts(const float *K, const float *features, int n, int d, int ld, int max_ncp, int lmin, int lmax,
                       int auto_select, double vmax, int desc_rate, int32_t *cps, int32_t *n_cps, double *scores,
                       void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE((K || features) && cps && n_cps && scores && ws, "kts: NULL pointer");
    SMZ_REQUIRE(n > 0 && max_ncp >= 0 && lmin >= 1 && lmax >= lmin, "kts: bad sizes");
    SMZ_REQUIRE((int64_t)(max_ncp + 1) * lmin <= n && n <= (int64_t)(max_ncp + 1) * lmax, "kts: need (m+1)*lmin <= n <= (m+1)*lmax");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    const KtsPlan pl = kts_plan(n, max_ncp, K == nullptr);
    SMZ_REQUIRE(ws_bytes >= pl.total, "kts: work buffer too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)pl.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    if (K == nullptr) {
        SMZ_REQUIRE(d > 0 && ld >= d, "kts: bad feature shape");
        float *Kw = reinterpret_cast<float *>(w + pl.off_K);
        rc = smz_kts_gram(features, n, d, ld, Kw, stream);
        if (rc != SMZ_OK) return rc;
        K = Kw;
    }
    double *K2 = reinterpret_cast<double *>(w + pl.off_K2), *K1 = reinterpret_cast<double *>(w + pl.off_K1);
    double *D2 = reinterpret_cast<double *>(w + pl.off_D2), *I = reinterpret_cast<double *>(w + pl.off_I);
    int32_t *p = reinterpret_cast<int32_t *>(w + pl.off_p);
    const int nb = (n + 1 + 127) / 128;
    prefix_cols_kernel<<<nb, 128, 0, st>>>(K, n, K2);
    prefix_rows_kernel<<<nb, 128, 0, st>>>(n, K2, D2);
    prefix_diag_kernel<<<1, 32, 0, st>>>(K, n, K1);
    kts_init_kernel<<<nb, 128, 0, st>>>(n, lmin, lmax, K1, D2, K2, I);
    kts_score_kernel<<<1, 32, 0, st>>>(0, n, I, scores);
    for (int k = 1; k <= max_ncp; k++) {
        double *prev = I + (size_t)((k - 1) & 1) * (n + 1), *cur = I + (size_t)(k & 1) * (n + 1);
        kts_dp_kernel<<<(n + 1 + 7) / 8, 256, 0, st>>>(k, n, lmin, lmax, K1, D2, K2, prev, cur, p + (size_t)k * (n + 1));
        kts_score_kernel<<<1, 32, 0, st>>>(k, n, cur, scores);
    }
    SMZ_CUDA_CHECK(cudaGetLastError());
    kts_pick_kernel<<<1, 32, 0, st>>>(n, max_ncp, auto_select, vmax, desc_rate < 1 ? 1 : desc_rate, scores, p, cps, n_cps);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
