// SumGAN's LSTMs (models/sumgan.py:23-115,185-210): the recurrences torch hands to cuDNN there —
//   sLSTM  nn.LSTM(1024, 1024, 2 layers, bidirectional)      sumgan.py:27-32,43
//   eLSTM  nn.LSTM(1024, 2048, 2 layers)                     sumgan.py:52-57,69
//   cLSTM  nn.LSTM(1024, 1024, 2 layers)                     sumgan.py:189-194,207
//   dLSTM  nn.LSTM(2048, 2048, 2 layers) called with seq_len 1 in a Python loop that feeds the top layer's
//          output back as the next input (sumgan.py:98-115)
// forward and BPTT; up to four sequences of equal length share a launch (kernels templated on that count): the trainer's
// batch is 1 (sumgan.py:404), but inside a training phase several sequences go through the same network.
//
// W_hh is 8 MB (H=1024) or 32 MB (H=2048) in bf16 — far beyond one cluster's registers (the DSN kernel,
// smz_dsn.cu) — so these recurrences are spread over 128 CTAs of a cooperative launch and the weights are streamed
// every step (L1/L2-resident for a layer; HBM for the 4 x 32 MB of the decoder).  A warp owns one or two hidden
// units with their four gate rows: 128-bit weight loads, fp32 FMAs against the h vector staged in shared memory,
// shuffle reductions, cell state in registers.  ONE grid-wide barrier per step and layer (monotonic global
// counter): h_t is exchanged through the output row it has to be written to anyway.  The input projections,
// dX and all weight gradients are tcgen05 GEMMs over whole sequences issued by the caller
// (summarizer_b200/models/lstm_stack.py), which is what removes the reference's per-step wgrad accumulation.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "smz_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

constexpr int CTAS = 128;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAX_UPW = 2;            // hidden units per warp: H / (CTAs per direction * 8)
constexpr int MAX_B = 4;              // sequences of equal length sharing the weights in one launch

struct SeqArgs { smz_lstm_seq d[2]; int n_dir; };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// `group` CTAs arrive; counter is monotonic, target = arrivals expected so far.  All CTAs are co-resident
// (cooperative launch); the spin is bounded so that a protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void group_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 6000000000LL) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float dot8(const uint4 w, const float4 a, const float4 b) {
    float s = __uint_as_float(w.x << 16) * a.x;
    s = fmaf(__uint_as_float(w.x & 0xffff0000u), a.y, s);
    s = fmaf(__uint_as_float(w.y << 16), a.z, s);
    s = fmaf(__uint_as_float(w.y & 0xffff0000u), a.w, s);
    s = fmaf(__uint_as_float(w.z << 16), b.x, s);
    s = fmaf(__uint_as_float(w.z & 0xffff0000u), b.y, s);
    s = fmaf(__uint_as_float(w.w << 16), b.z, s);
    s = fmaf(__uint_as_float(w.w & 0xffff0000u), b.w, s);
    return s;
}

// acc[b][g] += W[g*H + unit, 0:K] . vec_b for the four gates (W bf16 [4H, K] row-major, vec_b = vec + b*vstride fp32 in
// shared memory): per-lane partial sums, reduced by the caller; the weight chunk is loaded once for all NB sequences
template <int NB>
__device__ __forceinline__ void dot_gates_b(const bf16 *__restrict__ W, int H, int K, int unit, const float *vec, int vstride,
                                            int lane, float acc[NB][4]) {
    const bf16 *r0 = W + (size_t)unit * K;
    const size_t gs = (size_t)H * K;
#pragma unroll(NB == 1 ? 4 : 2)
    for (int c = lane * 8; c < K; c += 256) {
        uint4 w[4];
#pragma unroll
        for (int g = 0; g < 4; g++) w[g] = __ldg(reinterpret_cast<const uint4 *>(r0 + g * gs + c));
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const float4 x0 = *reinterpret_cast<const float4 *>(vec + (size_t)b * vstride + c);
            const float4 x1 = *reinterpret_cast<const float4 *>(vec + (size_t)b * vstride + c + 4);
#pragma unroll
            for (int g = 0; g < 4; g++) acc[b][g] += dot8(w[g], x0, x1);
        }
    }
}

// acc[b] += row[0:K] . vec_b (per-lane partials)
template <int NB>
__device__ __forceinline__ void dot_row_b(const bf16 *__restrict__ row, int K, const float *vec, int vstride, int lane,
                                          float acc[NB]) {
#pragma unroll(NB == 1 ? 8 : 4)
    for (int c = lane * 8; c < K; c += 256) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(row + c));
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const float4 x0 = *reinterpret_cast<const float4 *>(vec + (size_t)b * vstride + c);
            const float4 x1 = *reinterpret_cast<const float4 *>(vec + (size_t)b * vstride + c + 4);
            acc[b] += dot8(w, x0, x1);
        }
    }
}

// n floats (multiple of 4, 16-byte aligned) written by other SMs -> shared memory; src == nullptr stages zeros
__device__ __forceinline__ void stage(float *dst, const float *src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += THREADS * 4) {
        const float4 v = src ? __ldcg(reinterpret_cast<const float4 *>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(dst + i) = v;
    }
}

struct Cell { float i, f, g, o, c, h; };

__device__ __forceinline__ Cell lstm_cell(const float z[4], float c_prev) {
    Cell r;
    r.i = sigmoidf_(z[0]); r.f = sigmoidf_(z[1]); r.g = tanhf_(z[2]); r.o = sigmoidf_(z[3]);
    r.c = r.f * c_prev + r.i * r.g;
    r.h = r.o * tanhf_(r.c);
    return r;
}

// gradient of one cell: dh, dc_carry (from t+1) -> dz[4] (pre-activation gradients i,f,g,o), returns dc for t-1
__device__ __forceinline__ float lstm_cell_bwd(float dh, float dc_carry, float gi, float gf, float gg, float go, float c,
                                               float c_prev, float dz[4]) {
    const float tc = tanhf_(c);
    const float dc = dc_carry + dh * go * (1.f - tc * tc);
    dz[0] = dc * gg * gi * (1.f - gi);
    dz[1] = dc * c_prev * gf * (1.f - gf);
    dz[2] = dc * gi * (1.f - gg * gg);
    dz[3] = dh * tc * go * (1.f - go);
    return dc * gf;
}

// ---------------------------------------------------------------------------------------------------------
// One layer, whole sequence, n_dir directions side by side (CTAS / n_dir CTAs each).
// ---------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_seq_fwd_kernel(const SeqArgs p, unsigned int *counters) {
    extern __shared__ float sm[];                           // h_{t-1} of every sequence: B x H floats
    const int gd = gridDim.x / p.n_dir;
    const int dir = blockIdx.x / gd, cta = blockIdx.x % gd;
    const smz_lstm_seq &d = p.d[dir];
    const int H = d.H, T = d.T, upc = H / gd, upw = upc / WARPS;
    constexpr int B = NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned int *counter = counters + dir * 32;
    const bf16 *whh = reinterpret_cast<const bf16 *>(d.whh);
    unsigned int arrivals = 0;
    float c_state[MAX_UPW][NB];
#pragma unroll
    for (int u = 0; u < MAX_UPW; u++)
#pragma unroll
        for (int b = 0; b < NB; b++)
            c_state[u][b] = (u < upw && d.c0) ? d.c0[(size_t)b * H + cta * upc + warp * upw + u] : 0.f;
    for (int s = 0; s < T; s++) {
        const int t = d.reverse ? T - 1 - s : s;
        const int tp = d.reverse ? t + 1 : t - 1;
        for (int b = 0; b < B; b++) {
            const float *hprev = s == 0 ? (d.h0 ? d.h0 + (size_t)b * H : nullptr) : d.y + ((size_t)b * T + tp) * d.ldy;
            stage(sm + b * H, hprev, H);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
            float acc[NB][4];
#pragma unroll
            for (int b = 0; b < NB; b++)
#pragma unroll
                for (int g = 0; g < 4; g++) acc[b][g] = 0.f;
            dot_gates_b<NB>(whh, H, H, unit, sm, H, lane, acc);
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const size_t row = (size_t)b * T + t;
                float z[4];
#pragma unroll
                for (int g = 0; g < 4; g++) z[g] = __ldg(d.pre + row * d.ldpre + g * H + unit) + warp_sum(acc[b][g]);
                const Cell r = lstm_cell(z, c_state[u][b]);
                c_state[u][b] = r.c;
                if (lane == 0) {
                    d.y[row * d.ldy + unit] = r.h;
                    if (d.gates) {
                        float *gp = d.gates + row * d.ldg + unit;
                        gp[0] = r.i; gp[H] = r.f; gp[2 * H] = r.g; gp[3 * H] = r.o;
                    }
                    if (d.cs) d.cs[row * H + unit] = r.c;
                    if (s == T - 1) {
                        if (d.h_last) d.h_last[(size_t)b * H + unit] = r.h;
                        if (d.c_last) d.c_last[(size_t)b * H + unit] = r.c;
                    }
                }
            }
        }
        arrivals += gd;
        group_barrier(counter, arrivals);
    }
}

template <int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_seq_bwd_kernel(const SeqArgs p, unsigned int *counters) {
    extern __shared__ float sm[];                           // dz_t of every sequence: B x 4H floats
    const int gd = gridDim.x / p.n_dir;
    const int dir = blockIdx.x / gd, cta = blockIdx.x % gd;
    const smz_lstm_seq &d = p.d[dir];
    const int H = d.H, T = d.T, upc = H / gd, upw = upc / WARPS;
    constexpr int B = NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned int *counter = counters + dir * 32;
    const bf16 *whh_t = reinterpret_cast<const bf16 *>(d.whh_t);
    unsigned int arrivals = 0;
    float dh_rec[MAX_UPW][NB], dc_carry[MAX_UPW][NB];
#pragma unroll
    for (int u = 0; u < MAX_UPW; u++)
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const size_t k = (size_t)b * H + cta * upc + warp * upw + u;
            dh_rec[u][b] = (u < upw && d.dh_last) ? d.dh_last[k] : 0.f;
            dc_carry[u][b] = (u < upw && d.dc_last) ? d.dc_last[k] : 0.f;
        }
    for (int s = 0; s < T; s++) {
        const int t = d.reverse ? s : T - 1 - s;            // the forward pass visited t at step T-1-s
        const bool first = s == T - 1;                      // t is where the forward pass started
        const int tp = d.reverse ? t + 1 : t - 1;
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const size_t row = (size_t)b * T + t;
                const float *gp = d.gates + row * d.ldg + unit;
                const float c_prev = first ? (d.c0 ? d.c0[(size_t)b * H + unit] : 0.f) : d.cs[((size_t)b * T + tp) * H + unit];
                const float dh = dh_rec[u][b] + (d.dy ? d.dy[row * d.lddy + unit] : 0.f);
                float dz[4];
                dc_carry[u][b] = lstm_cell_bwd(dh, dc_carry[u][b], gp[0], gp[H], gp[2 * H], gp[3 * H], d.cs[row * H + unit], c_prev, dz);
                if (lane == 0) {
                    float *o = d.dgates + row * d.ldg + unit;
                    o[0] = dz[0]; o[H] = dz[1]; o[2 * H] = dz[2]; o[3 * H] = dz[3];
                }
            }
        }
        arrivals += gd;
        group_barrier(counter, arrivals);
        for (int b = 0; b < B; b++) stage(sm + b * 4 * H, d.dgates + ((size_t)b * T + t) * d.ldg, 4 * H);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
            float acc[NB] = {};
            dot_row_b<NB>(whh_t + (size_t)unit * 4 * H, 4 * H, sm, 4 * H, lane, acc);
#pragma unroll
            for (int b = 0; b < NB; b++) dh_rec[u][b] = warp_sum(acc[b]);
        }
    }
#pragma unroll
    for (int u = 0; u < MAX_UPW; u++) {
        if (u >= upw) break;
        const int unit = cta * upc + warp * upw + u;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (lane == 0) {
                if (d.dh0) d.dh0[(size_t)b * H + unit] = dh_rec[u][b];
                if (d.dc0) d.dc0[(size_t)b * H + unit] = dc_carry[u][b];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// dLSTM decode (sumgan.py:98-115): two stacked layers stepped together, layer 0's input at step t is layer 1's
// output at step t-1 (zeros at t = 0).
// ---------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_decode_fwd_kernel(const smz_lstm_decode d, unsigned int *counter) {
    extern __shared__ float sm[];                           // vA, vB: 2 x B x H floats
    const int H = d.H, T = d.T, upc = H / gridDim.x, upw = upc / WARPS;
    constexpr int B = NB;
    const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *vA = sm, *vB = sm + B * H;
    const bf16 *wih[2] = {reinterpret_cast<const bf16 *>(d.w_ih0), reinterpret_cast<const bf16 *>(d.w_ih1)};
    const bf16 *whh[2] = {reinterpret_cast<const bf16 *>(d.w_hh0), reinterpret_cast<const bf16 *>(d.w_hh1)};
    const float *bias[2] = {d.bias0, d.bias1};
    float *hs[2] = {d.hs0, d.hs1}, *gates[2] = {d.gates0, d.gates1}, *cs[2] = {d.cs0, d.cs1};
    unsigned int arrivals = 0;
    float c_state[2][MAX_UPW][NB];
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++)
#pragma unroll
            for (int b = 0; b < NB; b++)
                c_state[l][u][b] = (u < upw) ? d.c_init[((size_t)l * B + b) * H + cta * upc + warp * upw + u] : 0.f;
    for (int t = 0; t < T; t++) {
#pragma unroll
        for (int l = 0; l < 2; l++) {
            // layer 0: input = top output of the previous step (zeros at t = 0); layer 1: input = layer 0's output of this step
            const bool has_in = l == 1 || t > 0;
            for (int b = 0; b < B; b++) {
                const float *in = l == 0 ? (t > 0 ? d.hs1 + ((size_t)b * T + t - 1) * H : nullptr) : d.hs0 + ((size_t)b * T + t) * H;
                const float *rec = t > 0 ? hs[l] + ((size_t)b * T + t - 1) * H : d.h_init + ((size_t)l * B + b) * H;
                stage(vA + b * H, in, H);
                stage(vB + b * H, rec, H);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < MAX_UPW; u++) {
                if (u >= upw) break;
                const int unit = cta * upc + warp * upw + u;
                float acc[NB][4];
#pragma unroll
                for (int b = 0; b < NB; b++)
#pragma unroll
                    for (int g = 0; g < 4; g++) acc[b][g] = 0.f;
                if (has_in) dot_gates_b<NB>(wih[l], H, H, unit, vA, H, lane, acc);
                dot_gates_b<NB>(whh[l], H, H, unit, vB, H, lane, acc);
#pragma unroll
                for (int b = 0; b < NB; b++) {
                        const size_t row = (size_t)b * T + t;
                    float z[4];
#pragma unroll
                    for (int g = 0; g < 4; g++) z[g] = __ldg(bias[l] + g * H + unit) + warp_sum(acc[b][g]);
                    const Cell r = lstm_cell(z, c_state[l][u][b]);
                    c_state[l][u][b] = r.c;
                    if (lane == 0) {
                        hs[l][row * H + unit] = r.h;
                        if (gates[l]) {
                            float *gp = gates[l] + row * 4 * H + unit;
                            gp[0] = r.i; gp[H] = r.f; gp[2 * H] = r.g; gp[3 * H] = r.o;
                        }
                        if (cs[l]) cs[l][row * H + unit] = r.c;
                    }
                }
            }
            arrivals += gridDim.x;
            group_barrier(counter, arrivals);
        }
    }
}

template <int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_decode_bwd_kernel(const smz_lstm_decode d, unsigned int *counter) {
    extern __shared__ float sm[];                           // dz of every sequence: B x 4H floats
    const int H = d.H, T = d.T, upc = H / gridDim.x, upw = upc / WARPS, G4 = 4 * H;
    constexpr int B = NB;
    const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16 *wih0_t = reinterpret_cast<const bf16 *>(d.w_ih0_t), *whh0_t = reinterpret_cast<const bf16 *>(d.w_hh0_t);
    const bf16 *wih1_t = reinterpret_cast<const bf16 *>(d.w_ih1_t), *whh1_t = reinterpret_cast<const bf16 *>(d.w_hh1_t);
    unsigned int arrivals = 0;
    float r0[MAX_UPW][NB], r1[MAX_UPW][NB], r1a[MAX_UPW][NB], dc0[MAX_UPW][NB], dc1[MAX_UPW][NB];
#pragma unroll
    for (int u = 0; u < MAX_UPW; u++)
#pragma unroll
        for (int b = 0; b < NB; b++) r0[u][b] = r1[u][b] = r1a[u][b] = dc0[u][b] = dc1[u][b] = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        // A: top layer cells
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const size_t row = (size_t)b * T + t;
                const float *gp = d.gates1 + row * G4 + unit;
                const float c_prev = t > 0 ? d.cs1[(row - 1) * H + unit] : d.c_init[((size_t)B + b) * H + unit];
                const float dh = r1[u][b] + d.dy[row * H + unit];
                float dz[4];
                dc1[u][b] = lstm_cell_bwd(dh, dc1[u][b], gp[0], gp[H], gp[2 * H], gp[3 * H], d.cs1[row * H + unit], c_prev, dz);
                if (lane == 0) {
                    float *o = d.dgates1 + row * G4 + unit;
                    o[0] = dz[0]; o[H] = dz[1]; o[2 * H] = dz[2]; o[3 * H] = dz[3];
                }
            }
        }
        arrivals += gridDim.x;
        group_barrier(counter, arrivals);
        // B: dz1_t -> gradient of layer 0's output (its input role in layer 1) and of h1_{t-1}; layer 0 cells
        for (int b = 0; b < B; b++) stage(sm + b * G4, d.dgates1 + ((size_t)b * T + t) * G4, G4);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
            float via[NB] = {}, rec[NB] = {};
            dot_row_b<NB>(wih1_t + (size_t)unit * G4, G4, sm, G4, lane, via);
            dot_row_b<NB>(whh1_t + (size_t)unit * G4, G4, sm, G4, lane, rec);
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const size_t row = (size_t)b * T + t;
                const float via_in = warp_sum(via[b]);
                r1a[u][b] = warp_sum(rec[b]);
                const float *gp = d.gates0 + row * G4 + unit;
                const float c_prev = t > 0 ? d.cs0[(row - 1) * H + unit] : d.c_init[(size_t)b * H + unit];
                float dz[4];
                dc0[u][b] = lstm_cell_bwd(via_in + r0[u][b], dc0[u][b], gp[0], gp[H], gp[2 * H], gp[3 * H], d.cs0[row * H + unit], c_prev, dz);
                if (lane == 0) {
                    float *o = d.dgates0 + row * G4 + unit;
                    o[0] = dz[0]; o[H] = dz[1]; o[2 * H] = dz[2]; o[3 * H] = dz[3];
                }
            }
        }
        arrivals += gridDim.x;
        group_barrier(counter, arrivals);
        // C: dz0_t -> gradient of h0_{t-1} and (as layer 0's input, absent at t = 0) of h1_{t-1}
        for (int b = 0; b < B; b++) stage(sm + b * G4, d.dgates0 + ((size_t)b * T + t) * G4, G4);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAX_UPW; u++) {
            if (u >= upw) break;
            const int unit = cta * upc + warp * upw + u;
            float rec[NB] = {}, via[NB] = {};
            dot_row_b<NB>(whh0_t + (size_t)unit * G4, G4, sm, G4, lane, rec);
            if (t > 0) dot_row_b<NB>(wih0_t + (size_t)unit * G4, G4, sm, G4, lane, via);
#pragma unroll
            for (int b = 0; b < NB; b++) {
                r0[u][b] = warp_sum(rec[b]);
                r1[u][b] = r1a[u][b] + (t > 0 ? warp_sum(via[b]) : 0.f);
            }
        }
        __syncthreads();                                    // sm is restaged right after the next barrier
    }
#pragma unroll
    for (int u = 0; u < MAX_UPW; u++) {
        if (u >= upw) break;
        const int unit = cta * upc + warp * upw + u;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (lane == 0) {
                d.dh_init[(size_t)b * H + unit] = r0[u][b]; d.dh_init[((size_t)B + b) * H + unit] = r1[u][b];
                d.dc_init[(size_t)b * H + unit] = dc0[u][b]; d.dc_init[((size_t)B + b) * H + unit] = dc1[u][b];
            }
        }
    }
}

int check_seq(const smz_lstm_seq *dirs, int n_dir, bool backward) {
    SMZ_REQUIRE(dirs != nullptr && (n_dir == 1 || n_dir == 2), "lstm_seq: one or two directions");
    for (int i = 0; i < n_dir; i++) {
        const smz_lstm_seq &d = dirs[i];
        const int gd = CTAS / n_dir;
        SMZ_REQUIRE(d.T > 0 && d.H > 0, "lstm_seq: empty sequence");
        SMZ_REQUIRE(d.B >= 1 && d.B <= MAX_B, "lstm_seq: 1..%d sequences per launch", MAX_B);
        SMZ_REQUIRE(d.H == dirs[0].H && d.T == dirs[0].T && d.B == dirs[0].B, "lstm_seq: directions differ in T, H or B");
        SMZ_REQUIRE(d.H % (gd * WARPS) == 0 && d.H / (gd * WARPS) <= MAX_UPW && d.H % 256 == 0,
                    "lstm_seq: hidden size %d is not supported with %d direction(s) (1024 or 2048; 1024 when bidirectional)", d.H, n_dir);
        SMZ_REQUIRE(d.ldg % 4 == 0, "lstm_seq: ldg must be a multiple of 4");
        if (!backward) {
            SMZ_REQUIRE(d.pre && d.whh && d.y, "lstm_seq_forward: NULL pointer");
            SMZ_REQUIRE(d.ldy % 4 == 0 && ((uintptr_t)d.y & 15) == 0, "lstm_seq_forward: y rows must be 16-byte aligned");
            SMZ_REQUIRE(d.h0 == nullptr || ((uintptr_t)d.h0 & 15) == 0, "lstm_seq_forward: h0 must be 16-byte aligned");
        } else {
            SMZ_REQUIRE(d.whh_t && d.gates && d.cs && d.dgates, "lstm_seq_backward: NULL pointer");
            SMZ_REQUIRE(((uintptr_t)d.dgates & 15) == 0, "lstm_seq_backward: dgates must be 16-byte aligned");
        }
    }
    return SMZ_OK;
}

#define SMZ_DISPATCH_NB(B, KERNEL, ...)                                   \
    ((B) == 1 ? launch_coop(KERNEL<1>, __VA_ARGS__) : (B) == 2 ? launch_coop(KERNEL<2>, __VA_ARGS__) \
     : (B) == 3 ? launch_coop(KERNEL<3>, __VA_ARGS__) : launch_coop(KERNEL<4>, __VA_ARGS__))

template <typename Kern, typename Arg>
int launch_coop(Kern kern, const Arg &arg, unsigned int *counters, int smem_bytes, cudaStream_t st) {
    SMZ_REQUIRE(smz::sm_count() >= CTAS, "the LSTM recurrence needs %d co-resident CTAs", CTAS);
    SMZ_CUDA_CHECK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SMZ_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256, st));
    void *args[] = {(void *)&arg, (void *)&counters};
    SMZ_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)kern, dim3(CTAS), dim3(THREADS), args, (size_t)smem_bytes, st));
    return SMZ_OK;
}

}  // namespace

extern "C" int smz_lstm_seq_forward(const smz_lstm_seq *dirs, int n_dir, void *sync_ws, void *stream) {
    int rc = check_seq(dirs, n_dir, false);
    if (rc != SMZ_OK) return rc;
    SMZ_REQUIRE(sync_ws != nullptr, "lstm_seq_forward: sync_ws (256 bytes of device memory) is required");
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    SeqArgs a;
    a.n_dir = n_dir;
    for (int i = 0; i < n_dir; i++) a.d[i] = dirs[i];
    if (n_dir == 1) a.d[1] = dirs[0];
    rc = SMZ_DISPATCH_NB(dirs[0].B, lstm_seq_fwd_kernel, a, reinterpret_cast<unsigned int *>(sync_ws), dirs[0].B * dirs[0].H * 4, (cudaStream_t)stream);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP((cudaStream_t)stream, "lstm_seq_forward");
    return SMZ_OK;
}

extern "C" int smz_lstm_seq_backward(const smz_lstm_seq *dirs, int n_dir, void *sync_ws, void *stream) {
    int rc = check_seq(dirs, n_dir, true);
    if (rc != SMZ_OK) return rc;
    SMZ_REQUIRE(sync_ws != nullptr, "lstm_seq_backward: sync_ws (256 bytes of device memory) is required");
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    SeqArgs a;
    a.n_dir = n_dir;
    for (int i = 0; i < n_dir; i++) a.d[i] = dirs[i];
    if (n_dir == 1) a.d[1] = dirs[0];
    rc = SMZ_DISPATCH_NB(dirs[0].B, lstm_seq_bwd_kernel, a, reinterpret_cast<unsigned int *>(sync_ws), dirs[0].B * dirs[0].H * 16, (cudaStream_t)stream);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP((cudaStream_t)stream, "lstm_seq_backward");
    return SMZ_OK;
}

static int check_decode(const smz_lstm_decode *d, bool backward) {
    SMZ_REQUIRE(d != nullptr && d->T > 0, "lstm_decode: empty sequence");
    SMZ_REQUIRE(d->B >= 1 && d->B <= MAX_B, "lstm_decode: 1..%d sequences per launch", MAX_B);
    SMZ_REQUIRE(d->H % (CTAS * WARPS) == 0 && d->H / (CTAS * WARPS) <= MAX_UPW, "lstm_decode: hidden size %d is not supported (1024 or 2048)", d->H);
    SMZ_REQUIRE(d->h_init && d->c_init && d->hs0 && d->hs1, "lstm_decode: NULL pointer");
    if (!backward) SMZ_REQUIRE(d->w_ih0 && d->w_hh0 && d->w_ih1 && d->w_hh1 && d->bias0 && d->bias1, "lstm_decode_forward: NULL weight");
    else SMZ_REQUIRE(d->w_ih0_t && d->w_hh0_t && d->w_ih1_t && d->w_hh1_t && d->gates0 && d->gates1 && d->cs0 && d->cs1 && d->dy &&
                     d->dgates0 && d->dgates1 && d->dh_init && d->dc_init, "lstm_decode_backward: NULL pointer");
    return SMZ_OK;
}

extern "C" int smz_lstm_decode_forward(const smz_lstm_decode *d, void *sync_ws, void *stream) {
    int rc = check_decode(d, false);
    if (rc != SMZ_OK) return rc;
    SMZ_REQUIRE(sync_ws != nullptr, "lstm_decode_forward: sync_ws (256 bytes of device memory) is required");
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    rc = SMZ_DISPATCH_NB(d->B, lstm_decode_fwd_kernel, *d, reinterpret_cast<unsigned int *>(sync_ws), d->B * d->H * 8, (cudaStream_t)stream);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP((cudaStream_t)stream, "lstm_decode_forward");
    return SMZ_OK;
}

extern "C" int smz_lstm_decode_backward(const smz_lstm_decode *d, void *sync_ws, void *stream) {
    int rc = check_decode(d, true);
    if (rc != SMZ_OK) return rc;
    SMZ_REQUIRE(sync_ws != nullptr, "lstm_decode_backward: sync_ws (256 bytes of device memory) is required");
    rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    rc = SMZ_DISPATCH_NB(d->B, lstm_decode_bwd_kernel, *d, reinterpret_cast<unsigned int *>(sync_ws), d->B * d->H * 16, (cudaStream_t)stream);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP((cudaStream_t)stream, "lstm_decode_backward");
    return SMZ_OK;
}
