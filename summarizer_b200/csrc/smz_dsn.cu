// DSN scorer (models/dsn.py:17-47): BiLSTM(1024 -> 256 per direction) + Linear(512,1) + sigmoid, and its
// backward (BPTT), on sm_100a.
//
//   pre   = xb . [W_ih ; W_ih_reverse]^T + (b_ih + b_hh)      [R, 2048] fp32   one tcgen05 GEMM for both directions
//   lstm_fwd_kernel    the recurrence — latency bound, so NOT a tensor-core kernel: one thread-block CLUSTER
//                      of 8 CTAs per (video, direction); every CTA keeps its 128 x 256 slice of W_hh (32 hidden
//                      units x 4 gates) RESIDENT IN REGISTERS as packed bf16 (64 registers per thread, loaded
//                      once), multiplies it with h_{t-1} (fp32, shared memory) using fp32 FMAs, reduces the 16
//                      partial rows of a warp with a 16-shuffle transpose-reduction, applies the gate
//                      non-linearities in registers (cell state never leaves registers) and broadcasts its 32 new
//                      h values to the 8 CTAs through distributed shared memory; one cluster barrier per step.
//   head               probs = sigmoid(y . w_out + b_out), one warp per frame.
//   lstm_bwd_kernel    same structure with W_hh^T slices (32 units x 1024 gate columns per CTA): dh -> gate
//                      gradients, dc carried in registers, gate gradients broadcast through DSMEM.
//   weight gradients   dW_ih = dG^T . xb, dW_hh = dG^T . h_{t-1} as MN-major tcgen05 GEMMs after the time loop.
#include <cooperative_groups.h>

#include <vector>

#include "smz_gemm.cuh"
#include "smz_rows.cuh"

namespace cg = cooperative_groups;

namespace {

using smz::GemmEpilogue;
using smz::GemmProblem;
typedef __nv_bfloat16 bf16;

constexpr int H = 256;            // hidden size per direction (dsn.py:19)
constexpr int G4 = 4 * H;         // gate rows per direction
constexpr int CL = 8;             // CTAs per cluster
constexpr int UNITS = H / CL;     // hidden units owned by one CTA
constexpr int LSTM_THREADS = 256; // 8 warps, 4 units each
constexpr int WREGS = 64;         // packed bf16x2 weight registers per thread
constexpr int SAVE = 5;           // saved per unit and step: i, f, g, o, c

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }

// 16 values per lane, 32 lanes: afterwards lane L holds the warp-wide sum of value (L >> 1).
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const bool hi = lane & 16;
        const float send = hi ? v[j] : v[j + 8], keep = hi ? v[j + 8] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool hi = lane & 8;
        const float send = hi ? v[j] : v[j + 4], keep = hi ? v[j + 4] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const bool hi = lane & 4;
        const float send = hi ? v[j] : v[j + 2], keep = hi ? v[j + 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---------------------------------------------------------------------------------------------------
// forward recurrence.  whh: packed by smz_dsn_pack_whh: [dir][cta][k < 64][thread < 256] bf16x2.
// Thread (warp w, lane l): rows (unit 4w+u, gate g), u,g < 4, columns [8l, 8l+8); register (u*4+g)*4+p holds
// columns 8l+2p (low half) and 8l+2p+1 (high half).
// ---------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
lstm_fwd_kernel(const float *__restrict__ pre, const uint32_t *__restrict__ whh, const int32_t *__restrict__ cu,
                int n_videos, float *__restrict__ y, float *__restrict__ save) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float hbuf[2][H];
    const int cta = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ul = warp * 4 + (lane >> 3);             // local unit this lane's 8-lane group finalises
    const int unit = cta * UNITS + ul;                 // hidden unit index (0..255)
    float *remote_h = cluster.map_shared_rank(&hbuf[0][0], lane & 7);   // destination CTA of this lane's broadcast

    int loaded_dir = -1;
    uint32_t W[WREGS];
    for (int job = cluster_id; job < 2 * n_videos; job += n_clusters) {
        const int v = job >> 1, dir = job & 1;
        const int row0 = cu[v], T = cu[v + 1] - row0;
        if (dir != loaded_dir) {
            const uint32_t *src = whh + ((size_t)(dir * CL + cta) * WREGS) * LSTM_THREADS + tid;
#pragma unroll
            for (int k = 0; k < WREGS; k++) W[k] = __ldg(src + (size_t)k * LSTM_THREADS);
            loaded_dir = dir;
        }
        hbuf[0][tid] = 0.f;                            // h_{-1} = 0 (LSTM_THREADS == H)
        float c_state = 0.f;
        cluster.sync();
        const float *pre_col = pre + (size_t)dir * G4 + unit;       // + t*2048 + gate*256
        int t = dir ? T - 1 : 0;
        float p_i = pre_col[(size_t)(row0 + t) * (2 * G4)], p_f = pre_col[(size_t)(row0 + t) * (2 * G4) + H];
        float p_g = pre_col[(size_t)(row0 + t) * (2 * G4) + 2 * H], p_o = pre_col[(size_t)(row0 + t) * (2 * G4) + 3 * H];
        for (int s = 0; s < T; s++) {
            const int cur = s & 1;
            // prefetch the next step's input projections (hidden behind this step's compute)
            const int tn = dir ? t - 1 : t + 1;
            float n_i = 0.f, n_f = 0.f, n_g = 0.f, n_o = 0.f;
            if (s + 1 < T) {
                const float *q = pre_col + (size_t)(row0 + tn) * (2 * G4);
                n_i = q[0]; n_f = q[H]; n_g = q[2 * H]; n_o = q[3 * H];
            }
            const float4 h0 = *reinterpret_cast<const float4 *>(&hbuf[cur][8 * lane]);
            const float4 h1 = *reinterpret_cast<const float4 *>(&hbuf[cur][8 * lane + 4]);
            const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            float acc[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                float a = 0.f;
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint32_t w = W[k * 4 + p];
                    a = fmaf(bf_lo(w), hv[2 * p], a);
                    a = fmaf(bf_hi(w), hv[2 * p + 1], a);
                }
                acc[k] = a;
            }
            const float tot = transpose_reduce16(acc, lane);       // lane L: row (unit L>>3 local to the warp, gate (L>>1)&3)
            const int base = lane & 24;
            const float gi = sigmoidf_(__shfl_sync(0xffffffffu, tot, base + 0) + p_i);
            const float gf = sigmoidf_(__shfl_sync(0xffffffffu, tot, base + 2) + p_f);
            const float gg = tanhf_(__shfl_sync(0xffffffffu, tot, base + 4) + p_g);
            const float go = sigmoidf_(__shfl_sync(0xffffffffu, tot, base + 6) + p_o);
            c_state = gf * c_state + gi * gg;
            const float hn = go * tanhf_(c_state);
            // every lane of the 8-lane group holds the same (h, c): lane d sends h to CTA d
            remote_h[(cur ^ 1) * H + unit] = hn;
            const int d = lane & 7;
            const size_t row = (size_t)(row0 + t);
            if (d == 0) y[row * (2 * H) + dir * H + unit] = hn;
            if (save != nullptr && d < SAVE) {
                const float val = d == 0 ? gi : d == 1 ? gf : d == 2 ? gg : d == 3 ? go : c_state;
                save[((row * 2 + dir) * SAVE + d) * H + unit] = val;
            }
            p_i = n_i; p_f = n_f; p_g = n_g; p_o = n_o;
            t = tn;
            cluster.sync();
        }
    }
}

// probs = sigmoid(y . w + b): one warp per frame, y rows of 512
__global__ void __launch_bounds__(256)
dsn_head_kernel(const float *__restrict__ y, const float *__restrict__ w, const float *__restrict__ b, int rows,
                float *__restrict__ probs) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = lane * 4 + 128 * k;
        const float4 a = *reinterpret_cast<const float4 *>(y + (size_t)r * (2 * H) + c);
        const float4 ww = *reinterpret_cast<const float4 *>(w + c);
        dot += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) probs[r] = sigmoidf_(dot + __ldg(b));
}

int64_t up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct DsnPlan {
    int64_t rows, off_xb, off_pre, off_y, off_save, off_cu, off_dy, off_dg, off_dgb, total;
};

DsnPlan dsn_plan(int64_t rows, int n_videos, bool training, bool x_bf16) {
    DsnPlan p;
    p.rows = rows;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o += up(bytes, 1024); return at; };
    p.off_xb = take(x_bf16 ? 0 : rows * smz::kFeat * 2);
    p.off_pre = take(rows * 2 * G4 * 4);
    p.off_y = take(rows * 2 * H * 4);
    p.off_save = take(training ? rows * 2 * SAVE * H * 4 : 0);
    p.off_cu = take((int64_t)(n_videos + 1) * 4);
    p.off_dy = take(training ? rows * 2 * H * 4 : 0);
    p.off_dg = take(training ? rows * 2 * G4 * 4 : 0);
    p.off_dgb = take(training ? rows * 2 * G4 * 2 : 0);
    p.total = o;
    return p;
}

int lstm_grid(int n_videos) {
    // clusters of 8 CTAs cannot straddle GPCs: 16 co-resident clusters is a safe bound on the 148-SM part
    const int jobs = 2 * n_videos;
    return CL * (jobs < 16 ? jobs : 16);
}

}  // namespace

// Packs W_hh (float32 [1024, 256], torch layout: gate rows i,f,g,o) of both directions into the per-thread
// register order of lstm_fwd_kernel (out_fwd), and W_hh^T into the order of lstm_bwd_kernel (out_bwd).
// Device pointers; out_fwd / out_bwd hold 2 * 8 * 64 * 256 uint32 each.  Runs after every optimizer step.
namespace {
__global__ void pack_whh_kernel(const float *__restrict__ wf, const float *__restrict__ wb, uint32_t *__restrict__ out_fwd,
                                uint32_t *__restrict__ out_bwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // ((dir*CL + cta)*WREGS + k)*256 + tid
    if (idx >= 2 * CL * WREGS * LSTM_THREADS) return;
    const int tid = idx % LSTM_THREADS, k = (idx / LSTM_THREADS) % WREGS;
    const int cta = (idx / (LSTM_THREADS * WREGS)) % CL, dir = idx / (LSTM_THREADS * WREGS * CL);
    const float *Wm = dir ? wb : wf;
    const int warp = tid >> 5, lane = tid & 31;
    auto pack = [](float lo, float hi) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&t);
    };
    {   // forward: row (gate g, unit), columns 8*lane + 2p, +1
        const int u = k >> 4, g = (k >> 2) & 3, p = k & 3;
        const int unit = cta * UNITS + warp * 4 + u;
        const float *row = Wm + (size_t)(g * H + unit) * H;
        const int c = 8 * lane + 2 * p;
        out_fwd[idx] = pack(row[c], row[c + 1]);
    }
    {   // backward: (W_hh^T)[unit][gate column j]: units 4*warp+u (u<4), columns j = 32*lane + 2q, +1 (q<16); register u*16+q
        const int u = k >> 4, q = k & 15;
        const int unit = cta * UNITS + warp * 4 + u;
        const int j = 32 * lane + 2 * q;
        out_bwd[idx] = pack(Wm[(size_t)j * H + unit], Wm[(size_t)(j + 1) * H + unit]);
    }
}
}  // namespace

extern "C" int smz_dsn_pack_whh(const float *whh_fwd, const float *whh_bwd, uint32_t *out_fwd, uint32_t *out_bwd, void *stream) {
    SMZ_REQUIRE(whh_fwd && whh_bwd && out_fwd && out_bwd, "pack_whh: NULL pointer");
    const int n = 2 * CL * WREGS * LSTM_THREADS;
    pack_whh_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(whh_fwd, whh_bwd, out_fwd, out_bwd);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_dsn_workspace_bytes(int total_rows, int n_videos, int training, int x_is_bf16, int64_t *bytes) {
    SMZ_REQUIRE(bytes != nullptr && total_rows > 0 && n_videos > 0, "dsn_workspace_bytes: bad argument");
    *bytes = dsn_plan(total_rows, n_videos, training != 0, x_is_bf16 != 0).total;
    return SMZ_OK;
}

extern "C" int smz_dsn_forward(const void *x, int x_is_bf16, const int32_t *h_cu_seqlens, int n_videos,
                               const smz_dsn_params *p, int training, float *probs, void *ws, int64_t ws_bytes,
                               void *stream) {
    SMZ_REQUIRE(x && h_cu_seqlens && p && probs && ws && n_videos > 0, "dsn_forward: NULL pointer");
    SMZ_REQUIRE(p->w_ih && p->bias && p->whh_packed && p->w_out && p->b_out, "dsn_forward: NULL parameter pointer");
    SMZ_REQUIRE(h_cu_seqlens[0] == 0, "dsn_forward: cu_seqlens[0] must be 0");
    for (int v = 0; v < n_videos; v++) SMZ_REQUIRE(h_cu_seqlens[v + 1] > h_cu_seqlens[v], "dsn_forward: video %d has no frames", v);
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    const int R = h_cu_seqlens[n_videos];
    const DsnPlan pl = dsn_plan(R, n_videos, training != 0, x_is_bf16 != 0);
    SMZ_REQUIRE(ws_bytes >= pl.total, "dsn_forward: work buffer too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)pl.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    int32_t *d_cu = reinterpret_cast<int32_t *>(w + pl.off_cu);
    SMZ_CUDA_CHECK(cudaMemcpyAsync(d_cu, h_cu_seqlens, (size_t)(n_videos + 1) * 4, cudaMemcpyHostToDevice, st));
    const bf16 *xb = reinterpret_cast<const bf16 *>(x);
    if (!x_is_bf16) {
        bf16 *dst = reinterpret_cast<bf16 *>(w + pl.off_xb);
        rc = smz::launch_cvt_bf16(reinterpret_cast<const float *>(x), dst, (int64_t)R * smz::kFeat, st);
        if (rc != SMZ_OK) return rc;
        xb = dst;
    }
    float *pre = reinterpret_cast<float *>(w + pl.off_pre);
    float *y = reinterpret_cast<float *>(w + pl.off_y);
    float *save = training ? reinterpret_cast<float *>(w + pl.off_save) : nullptr;
    // input projections of both directions + (b_ih + b_hh)
    GemmProblem g = {};
    g.M = R; g.N = 2 * G4; g.K = smz::kFeat; g.ldc = 2 * G4; g.tiles_n = (2 * G4) / smz::GEMM_BN;
    rc = smz::gemm_bf16_tn(xb, R, smz::kFeat, smz::kFeat, p->w_ih, 2 * G4, smz::kFeat, smz::kFeat, nullptr, 1,
                           smz::gemm_tiles(R, 2 * G4), g, GemmEpilogue{pre, p->bias, nullptr, 1.f, smz::GEMM_OUT_F32}, st);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP(st, "dsn_input_projection");
    lstm_fwd_kernel<<<lstm_grid(n_videos), LSTM_THREADS, 0, st>>>(pre, reinterpret_cast<const uint32_t *>(p->whh_packed), d_cu,
                                                                  n_videos, y, save);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_lstm_fwd");
    dsn_head_kernel<<<(R + 7) / 8, 256, 0, st>>>(y, p->w_out, p->b_out, R, probs);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_head");
    return SMZ_OK;
}
