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
//                      h values to the 8 CTAs through distributed shared memory with st.async, which completes
//                      bytes on the DESTINATION CTA's mbarrier: a step waits for "1 KB of h has landed" on its
//                      own barrier instead of a cluster-wide barrier + fence (1.27 -> 0.97 us per step).
//   head               probs = sigmoid(y . w_out + b_out), one warp per frame.
//   lstm_bwd_kernel    same structure with W_hh^T slices (32 units x 1024 gate columns per CTA): dh -> gate
//                      gradients, dc carried in registers, gate gradients broadcast through DSMEM (same st.async /
//                      mbarrier protocol, 4 KB per step).
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

// ---- DSMEM signalling: a remote store that completes bytes on the DESTINATION CTA's mbarrier -------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)),
                   "r"(__float_as_uint(d)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arm(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must not hang the GPU
    }
}

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
                int n_videos, float *__restrict__ y, float *__restrict__ save, __nv_bfloat16 *__restrict__ hprev) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float hbuf[2][H];
    __shared__ __align__(8) uint64_t hbar[2];          // hbar[b]: all 256 floats of hbuf[b] have landed (1 KB per phase)
    const int cta = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ul = warp * 4 + (lane >> 3);             // local unit this lane's 8-lane group finalises
    const int unit = cta * UNITS + ul;                 // hidden unit index (0..255)
    // destination CTA of this lane's broadcast: its hbuf slot of `unit` and its barriers, as shared::cluster addresses
    const uint32_t r_h = map_to_cta(smem_addr(&hbuf[0][unit]), lane & 7);
    const uint32_t r_bar = map_to_cta(smem_addr(&hbar[0]), lane & 7);
    if (tid == 0) {
        bar_init(&hbar[0], 1); bar_init(&hbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ph0 = 0, ph1 = 0;                         // wait parities of hbar[0] / hbar[1]
    cluster.sync();

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
            // Step s consumes hbuf[s & 1].  No cluster barrier: every CTA's new h values arrive through st.async,
            // which completes 1 KB on OUR mbarrier of that buffer; a CTA can only run one step ahead of the slowest
            // one (it needs everybody's h), so the buffer being overwritten has been read by all of our warps.
            if (s > 0) {
                if (cur) { bar_wait(&hbar[1], ph1); ph1 ^= 1u; } else { bar_wait(&hbar[0], ph0); ph0 ^= 1u; }
            }
            if (tid == 0 && s + 1 < T) bar_arm(&hbar[cur ^ 1], H * 4);
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
            if (s + 1 < T) st_async_f32(r_h + (cur ^ 1) * H * 4, hn, r_bar + (cur ^ 1) * 8);
            const int d = lane & 7;
            const size_t row = (size_t)(row0 + t);
            if (d == 0) y[row * (2 * H) + dir * H + unit] = hn;
            if (hprev != nullptr && d == 5) {           // h_{t-1} rows for the dW_hh GEMM (zero at the sequence start)
                if (s == 0) hprev[row * (2 * H) + dir * H + unit] = __float2bfloat16_rn(0.f);
                if (s + 1 < T) hprev[(size_t)(row0 + tn) * (2 * H) + dir * H + unit] = __float2bfloat16_rn(hn);
            }
            if (save != nullptr && d < SAVE) {
                const float val = d == 0 ? gi : d == 1 ? gf : d == 2 ? gg : d == 3 ? go : c_state;
                save[((row * 2 + dir) * SAVE + d) * H + unit] = val;
            }
            p_i = n_i; p_f = n_f; p_g = n_g; p_o = n_o;
            t = tn;
        }
        cluster.sync();                                // job boundary: nothing of this sequence is in flight any more
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


// ---------------------------------------------------------------------------------------------------
// backward recurrence (BPTT).  wt: W_hh^T packed by smz_dsn_pack_whh: thread (warp w, lane l) owns units
// 4w+u (u<4) and gate columns 32l + 2q, +1 (q<16); register u*16+q.  dz[t] = dL/d(pre-sigmoid head logit).
// Writes the gate gradients of every step as bf16 (dgb [R, 2048]) for the weight-gradient GEMMs and
// accumulates the bias gradients (d_bias [2048], float atomics at the end of a sequence).
// ---------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
lstm_bwd_kernel(const uint32_t *__restrict__ wt, const int32_t *__restrict__ cu, int n_videos,
                const float *__restrict__ save, const float *__restrict__ dz, const float *__restrict__ w_out,
                __nv_bfloat16 *__restrict__ dgb, float *__restrict__ d_bias) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float dgbuf[2][G4];
    __shared__ __align__(8) uint64_t gbar[2];          // gbar[b]: all 1024 gate gradients of dgbuf[b] have landed (4 KB)
    const int cta = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int unit = cta * UNITS + warp * 4 + (lane >> 3);
    const int d = lane & 7;
    // gate gradients are exchanged in [unit][gate] order: the four of one unit are ONE 16-byte st.async per destination
    const uint32_t r_g = map_to_cta(smem_addr(&dgbuf[0][4 * unit]), d);  // lane d -> CTA d
    const uint32_t r_bar = map_to_cta(smem_addr(&gbar[0]), d);
    if (tid == 0) {
        bar_init(&gbar[0], 1); bar_init(&gbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ph0 = 0, ph1 = 0;
    cluster.sync();

    int loaded_dir = -1;
    uint32_t W[WREGS];
    for (int job = cluster_id; job < 2 * n_videos; job += n_clusters) {
        const int v = job >> 1, dir = job & 1;
        const int row0 = cu[v], T = cu[v + 1] - row0;
        if (dir != loaded_dir) {
            const uint32_t *src = wt + ((size_t)(dir * CL + cta) * WREGS) * LSTM_THREADS + tid;
#pragma unroll
            for (int k = 0; k < WREGS; k++) W[k] = __ldg(src + (size_t)k * LSTM_THREADS);
            loaded_dir = dir;
        }
        for (int j = tid; j < G4; j += LSTM_THREADS) dgbuf[0][j] = 0.f;     // no gradient flows in from beyond the last step
        const float wo = __ldg(w_out + dir * H + unit);
        float dc_carry = 0.f, bsum[4] = {0.f, 0.f, 0.f, 0.f};
        cluster.sync();
        // steps are visited in the reverse of the forward order of this direction
        int t = dir ? 0 : T - 1;
        const float *sv = save + ((size_t)(row0 + t) * 2 + dir) * SAVE * H + unit;
        float gi = sv[0], gf = sv[H], gg = sv[2 * H], go = sv[3 * H], cc = sv[4 * H];
        float dzt = dz[row0 + t];
        for (int s = 0; s < T; s++) {
            const int cur = s & 1;
            if (s > 0) {                                   // same st.async / mbarrier protocol as the forward kernel
                if (cur) { bar_wait(&gbar[1], ph1); ph1 ^= 1u; } else { bar_wait(&gbar[0], ph0); ph0 ^= 1u; }
            }
            if (tid == 0 && s + 1 < T) bar_arm(&gbar[cur ^ 1], G4 * 4);
            const int tp = dir ? t + 1 : t - 1;            // the step BEFORE t in forward order (next to visit)
            // prefetch the next visited step's saved activations; its cell state is this step's c_{prev}
            float n_i = 0.f, n_f = 0.f, n_g = 0.f, n_o = 0.f, n_c = 0.f, n_dz = 0.f;
            if (s + 1 < T) {
                const float *q = save + ((size_t)(row0 + tp) * 2 + dir) * SAVE * H + unit;
                n_i = q[0]; n_f = q[H]; n_g = q[2 * H]; n_o = q[3 * H]; n_c = q[4 * H];
                n_dz = dz[row0 + tp];                      // the head gradient of the next visited step, off the critical path
            }
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const float4 g4 = *reinterpret_cast<const float4 *>(&dgbuf[cur][32 * lane + 4 * q4]);
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t w0 = W[u * 16 + 2 * q4], w1 = W[u * 16 + 2 * q4 + 1];
                    acc[u] = fmaf(bf_lo(w0), gv[0], acc[u]); acc[u] = fmaf(bf_hi(w0), gv[1], acc[u]);
                    acc[u] = fmaf(bf_lo(w1), gv[2], acc[u]); acc[u] = fmaf(bf_hi(w1), gv[3], acc[u]);
                }
            }
            // 4 values x 32 lanes -> lanes 8u..8u+7 hold the total of unit u
            {
                const bool hi = lane & 16;
                const float s0 = hi ? acc[0] : acc[2], k0 = hi ? acc[2] : acc[0];
                const float s1 = hi ? acc[1] : acc[3], k1 = hi ? acc[3] : acc[1];
                acc[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                acc[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            }
            {
                const bool hi = lane & 8;
                const float s0 = hi ? acc[0] : acc[1], k0 = hi ? acc[1] : acc[0];
                acc[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
            }
            float rec = acc[0];
            rec += __shfl_xor_sync(0xffffffffu, rec, 4);
            rec += __shfl_xor_sync(0xffffffffu, rec, 2);
            rec += __shfl_xor_sync(0xffffffffu, rec, 1);
            const float dh = dzt * wo + rec;
            const float tc = tanhf_(cc);
            const float dO = dh * tc * go * (1.f - go);
            const float dc = dh * go * (1.f - tc * tc) + dc_carry;
            const float dI = dc * gg * gi * (1.f - gi);
            const float dG = dc * gi * (1.f - gg * gg);
            const float dF = dc * n_c * gf * (1.f - gf);      // n_c = c_{t-1} (0 at the sequence start)
            dc_carry = dc * gf;
            if (s + 1 < T) {
                const uint32_t dst = r_g + (cur ^ 1) * G4 * 4, bar = r_bar + (cur ^ 1) * 8;
                st_async_f32x4(dst, dI, dF, dG, dO, bar);
            }
            if (d < 4) {
                const float val = d == 0 ? dI : d == 1 ? dF : d == 2 ? dG : dO;
                dgb[(size_t)(row0 + t) * (2 * G4) + dir * G4 + d * H + unit] = __float2bfloat16_rn(val);
            }
            bsum[0] += dI; bsum[1] += dF; bsum[2] += dG; bsum[3] += dO;
            gi = n_i; gf = n_f; gg = n_g; go = n_o; cc = n_c; dzt = n_dz;
            t = tp;
        }
        if (d < 4) atomicAdd(d_bias + dir * G4 + d * H + unit, d == 0 ? bsum[0] : d == 1 ? bsum[1] : d == 2 ? bsum[2] : bsum[3]);
        cluster.sync();                                // job boundary
    }
}

// dz = dprobs * p * (1 - p);  d w_out += sum_t dz_t * y_t;  d b_out += sum_t dz_t
__global__ void __launch_bounds__(256)
dsn_head_bwd_kernel(const float *__restrict__ y, const float *__restrict__ probs, const float *__restrict__ dprobs, int rows,
                    float *__restrict__ dz, float *__restrict__ d_w_out, float *__restrict__ d_b_out) {
    float a0 = 0.f, a1 = 0.f, ab = 0.f;
    const int c = threadIdx.x * 2;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const float pr = probs[r];
        const float z = dprobs[r] * pr * (1.f - pr);
        if (threadIdx.x == 0) { dz[r] = z; ab += z; }
        const float2 yy = *reinterpret_cast<const float2 *>(y + (size_t)r * (2 * H) + c);
        a0 = fmaf(z, yy.x, a0); a1 = fmaf(z, yy.y, a1);
    }
    atomicAdd(d_w_out + c, a0); atomicAdd(d_w_out + c + 1, a1);
    if (threadIdx.x == 0) atomicAdd(d_b_out, ab);
}

int64_t up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct DsnPlan {
    int64_t rows, off_xb, off_pre, off_y, off_save, off_cu, off_hprev, off_dz, off_dgb, total;
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
    p.off_hprev = take(training ? rows * 2 * H * 2 : 0);
    p.off_dz = take(training ? rows * 4 : 0);
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
    {   // backward: (W_hh^T)[unit][.] against the gate-gradient buffer, which is ordered [unit'][gate]: position
        // jp = 4*unit' + gate <-> row gate*H + unit' of W_hh; units 4*warp+u (u<4), positions 32*lane + 2q, +1 (q<16); register u*16+q
        const int u = k >> 4, q = k & 15;
        const int unit = cta * UNITS + warp * 4 + u;
        const int jp = 32 * lane + 2 * q;
        const int r0 = (jp & 3) * H + (jp >> 2), r1 = ((jp + 1) & 3) * H + ((jp + 1) >> 2);
        out_bwd[idx] = pack(Wm[(size_t)r0 * H + unit], Wm[(size_t)r1 * H + unit]);
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
    rc = smz::upload_small(d_cu, h_cu_seqlens, (size_t)(n_videos + 1) * 4, st);
    if (rc != SMZ_OK) return rc;
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
    bf16 *hprev = training ? reinterpret_cast<bf16 *>(w + pl.off_hprev) : nullptr;
    lstm_fwd_kernel<<<lstm_grid(n_videos), LSTM_THREADS, 0, st>>>(pre, reinterpret_cast<const uint32_t *>(p->whh_packed), d_cu,
                                                                  n_videos, y, save, hprev);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_lstm_fwd");
    dsn_head_kernel<<<(R + 7) / 8, 256, 0, st>>>(y, p->w_out, p->b_out, R, probs);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_head");
    return SMZ_OK;
}

// Backward of smz_dsn_forward(training=1) on the same batch / work buffer.  ACCUMULATES (+=) float32 gradients:
//   d_w_ih [2048,1024] (rows 0..1023 forward direction), d_w_hh [2048,256] (rows 0..1023 = weight_hh_l0,
//   1024..2047 = weight_hh_l0_reverse), d_bias [2048] (the gradient of BOTH bias_ih and bias_hh), d_w_out [512],
//   d_b_out [1].
extern "C" int smz_dsn_backward(const void *x, int x_is_bf16, const int32_t *h_cu_seqlens, int n_videos,
                                const smz_dsn_params *p, const float *probs, const float *dprobs,
                                const smz_dsn_grads *gr, void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(x && h_cu_seqlens && p && probs && dprobs && gr && ws && n_videos > 0, "dsn_backward: NULL pointer");
    SMZ_REQUIRE(gr->w_ih && gr->w_hh && gr->bias && gr->w_out && gr->b_out, "dsn_backward: NULL gradient pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    const int R = h_cu_seqlens[n_videos];
    const DsnPlan pl = dsn_plan(R, n_videos, true, x_is_bf16 != 0);
    SMZ_REQUIRE(ws_bytes >= pl.total, "dsn_backward: work buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    const int32_t *d_cu = reinterpret_cast<const int32_t *>(w + pl.off_cu);      // uploaded by the forward call
    const bf16 *xb = x_is_bf16 ? reinterpret_cast<const bf16 *>(x) : reinterpret_cast<const bf16 *>(w + pl.off_xb);
    const float *y = reinterpret_cast<const float *>(w + pl.off_y);
    const float *save = reinterpret_cast<const float *>(w + pl.off_save);
    const bf16 *hprev = reinterpret_cast<const bf16 *>(w + pl.off_hprev);
    float *dz = reinterpret_cast<float *>(w + pl.off_dz);
    bf16 *dgb = reinterpret_cast<bf16 *>(w + pl.off_dgb);

    int hb_grid = R < 296 ? R : 296;
    dsn_head_bwd_kernel<<<hb_grid, 256, 0, st>>>(y, probs, dprobs, R, dz, gr->w_out, gr->b_out);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_head_bwd");
    lstm_bwd_kernel<<<lstm_grid(n_videos), LSTM_THREADS, 0, st>>>(reinterpret_cast<const uint32_t *>(p->whh_t_packed), d_cu, n_videos,
                                                                  save, dz, p->w_out, dgb, gr->bias);
    SMZ_CUDA_CHECK(cudaGetLastError());
    SMZ_DEBUG_STEP(st, "dsn_lstm_bwd");
    const int ACC = smz::GEMM_OUT_F32 | smz::GEMM_RES_F32;
    // dW_ih += dG^T . xb  (both directions at once: M = 2048 gate rows)
    {
        GemmProblem g = {};
        g.M = 2 * G4; g.N = smz::kFeat; g.K = R; g.ldc = smz::kFeat; g.ldr = smz::kFeat; g.tiles_n = smz::kFeat / smz::GEMM_BN;
        rc = smz::gemm_bf16(true, true, dgb, R, 2 * G4, 2 * G4, xb, R, smz::kFeat, smz::kFeat, nullptr, 1,
                            smz::gemm_tiles(g.M, g.N), g, GemmEpilogue{gr->w_ih, nullptr, gr->w_ih, 1.f, ACC}, st);
        if (rc != SMZ_OK) return rc;
    }
    // dW_hh[dir] += dG_dir^T . h_{t-1, dir}
    for (int dir = 0; dir < 2; dir++) {
        GemmProblem g = {};
        g.M = G4; g.N = H; g.K = R; g.ldc = H; g.ldr = H; g.tiles_n = 1;
        g.a_col0 = dir * G4; g.b_col0 = dir * H;
        g.c_off = (int64_t)dir * G4 * H; g.r_off = g.c_off;
        rc = smz::gemm_bf16(true, true, dgb, R, 2 * G4, 2 * G4, hprev, R, 2 * H, 2 * H, nullptr, 1, smz::gemm_tiles(g.M, g.N), g,
                            GemmEpilogue{gr->w_hh, nullptr, gr->w_hh, 1.f, ACC}, st);
        if (rc != SMZ_OK) return rc;
    }
    SMZ_DEBUG_STEP(st, "dsn_wgrad");
    return SMZ_OK;
}
