// Shot selection and F-score evaluation kernels (sm_100a) + their C ABI.
//
// Replaces, on the device, the reference's host-side numpy/Python/OR-tools path
//   utils/eval.py:15-35   upsample
//   utils/eval.py:74-123  generate_summary   (segment mean -> knapsack/rank -> summary vector)
//   utils/knapsack.py:5-23 knapsack_ortools  (OR-tools 7.5 KnapsackDynamicProgrammingSolver)
//   utils/eval.py:125-165 evaluate_summary   (per-user overlap / precision / recall / F)
// with results that are bit-identical for every integer (values, selections, masks, counts)
// and for the float32 quantities the reference computes in float32 (segment means, F).
//
// Kernels
//   select_kernel   one persistent CTA per video (grid = #SMs x occupancy): segment pooling in
//                   numpy's pairwise summation order, value quantisation in fp64, 0/1 knapsack DP
//                   with the DP row in shared memory (double buffered, one __syncthreads per item),
//                   a take-bit matrix (shared memory when it fits, else the work buffer) and the
//                   OR-tools backtrack, then summary vector + bit mask + popcount.
//                   Shared-memory/latency bound.
//   fscore_kernel   grid (frame chunk, video); streams user_summary once with 128-bit loads,
//                   counts with integer ops, one packed REDUX per user per warp.  HBM bound.
//   fscore_final    one thread per video: float32 F per user, numpy-order mean, max.
#include "smz_common.cuh"

#include <limits.h>

namespace {

constexpr int SELECT_THREADS = 512;
constexpr int FSCORE_THREADS = 256;
constexpr int FSCORE_MAX_USERS = 1024;
constexpr int FSCORE_UNROLL = 4;
static_assert(SMZ_FSCORE_CHUNK == FSCORE_THREADS * 8, "each thread owns two float4 columns");

// ------------------------------------------------------------------------------------------
// upsample (utils/eval.py:15-35) as a forward cursor over frames
// ------------------------------------------------------------------------------------------
struct FrameCursor {
    const float *scores;
    const int32_t *picks;
    int n_scores, n_picks, n_frames;
    int n_bound;  // number of interval boundaries (n_picks, +1 when n_frames gets appended)
    int idx;      // interval containing the last frame asked for; -1 = before picks[0]
    int next;     // first frame of interval idx+1 (INT_MAX when there is none)

    __device__ __forceinline__ int bound(int i) const { return i < n_picks ? __ldg(picks + i) : n_frames; }

    __device__ void init(const float *s, const int32_t *p, int ns, int np, int nf) {
        scores = s; picks = p; n_scores = ns; n_picks = np; n_frames = nf;
        n_bound = np + ((np > 0 && __ldg(p + np - 1) != nf) ? 1 : 0);
        idx = -1;
        next = n_bound > 0 ? bound(0) : INT_MAX;
    }
    // largest i with bound(i) <= f, or -1
    __device__ void seek(int f) {
        int lo = 0, hi = n_bound;  // first i with bound(i) > f
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (bound(mid) <= f) lo = mid + 1; else hi = mid;
        }
        idx = lo - 1;
        next = lo < n_bound ? bound(lo) : INT_MAX;
    }
    // frames must be asked for in non-decreasing order after seek()
    __device__ __forceinline__ float at(int f) {
        while (f >= next) {
            ++idx;
            next = (idx + 1 < n_bound) ? bound(idx + 1) : INT_MAX;
        }
        // interval idx exists iff idx+1 < n_bound; interval == n_scores is zero filled
        return (idx >= 0 && idx + 1 < n_bound && idx < n_scores) ? __ldg(scores + idx) : 0.f;
    }
};

struct ArrayCursor {
    const float *a;
    __device__ __forceinline__ float at(int i) const { return a[i]; }
};

struct ArrayCursorF64 {  // float32 values widened element-wise (list promoted to float64)
    const float *a;
    __device__ __forceinline__ double at(int i) const { return (double)a[i]; }
};

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// numpy FLOAT_pairwise_sum, n <= 128 (loops_utils.h.src): 8 strided accumulators.
template <class T, class Cur>
__device__ T pw_block(Cur &cur, int start, int n) {
    if (n < 8) {
        T res = T(0);
        for (int i = 0; i < n; i++) res = add_rn(res, cur.at(start + i));
        return res;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = cur.at(start + j);
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = add_rn(r[j], cur.at(start + i + j));
    }
    T res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])),
                   add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
    for (; i < n; i++) res = add_rn(res, cur.at(start + i));
    return res;
}

// full pairwise recursion (n > 128 splits at n/2 rounded down to a multiple of 8), evaluated
// with an explicit stack so the frame cursor keeps moving forward.
template <class T, class Cur>
__device__ T pw_sum(Cur &cur, int start, int n) {
    if (n <= 128) return pw_block<T>(cur, start, n);
    int s_start[32], s_n[32];
    T s_left[32];
    unsigned char s_stage[32];
    int sp = 0;
    T ret = T(0);
    s_start[0] = start; s_n[0] = n; s_stage[0] = 0;
    while (sp >= 0) {
        const int n_ = s_n[sp];
        if (s_stage[sp] == 0) {
            if (n_ <= 128) { ret = pw_block<T>(cur, s_start[sp], n_); --sp; continue; }
            int n2 = n_ / 2; n2 -= n2 % 8;
            s_stage[sp] = 1;
            s_start[sp + 1] = s_start[sp]; s_n[sp + 1] = n2; s_stage[sp + 1] = 0;
            ++sp;
        } else if (s_stage[sp] == 1) {
            int n2 = n_ / 2; n2 -= n2 % 8;
            s_left[sp] = ret;
            s_stage[sp] = 2;
            s_start[sp + 1] = s_start[sp] + n2; s_n[sp + 1] = n_ - n2; s_stage[sp + 1] = 0;
            ++sp;
        } else {
            ret = add_rn(s_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// select_kernel
// ------------------------------------------------------------------------------------------
struct SelectSmem {  // word offsets into dynamic shared memory
    int dp0, dp1, w, p, pk, mean, pre, mask, sel, bits, total;
};

__host__ __device__ inline SelectSmem select_layout(int max_n_segs, int max_capacity, int max_n_frames,
                                                    bool bits_in_smem) {
    SelectSmem L;
    const int ncell = max_capacity + 1;
    // the dp rows double as scratch (warp totals, 'rank' order): at least 32 and max_n_segs words
    int row = ncell > max_n_segs ? ncell : max_n_segs;
    row = row > 32 ? row : 32;
    int o = 0;
    L.dp0 = o; o += row;
    L.dp1 = o; o += row;
    L.w = o; o += max_n_segs;
    L.p = o; o += max_n_segs;
    L.pk = o; o += max_n_segs;
    L.mean = o; o += max_n_segs;
    L.pre = o; o += max_n_segs + 1;
    L.mask = o; o += (max_n_frames + 31) / 32;
    L.sel = o; o += 4;
    L.bits = o;
    if (bits_in_smem) o += max_n_segs * ((ncell + 31) / 32);
    L.total = o;
    return L;
}

template <bool BITS_IN_SMEM>
__global__ void __launch_bounds__(SELECT_THREADS)
select_kernel(const smz_video_desc *__restrict__ desc, int n_videos, const float *__restrict__ scores,
              const int32_t *__restrict__ picks, const int32_t *__restrict__ cps,
              const int32_t *__restrict__ nfps, const int32_t *__restrict__ values_in, int method,
              int max_n_segs, int max_capacity, int max_n_frames, float *__restrict__ out_mean, int32_t *__restrict__ out_values,
              uint8_t *__restrict__ out_picked, float *__restrict__ out_summary,
              uint32_t *__restrict__ out_mask, int32_t *__restrict__ out_msum,
              int32_t *__restrict__ out_status, uint32_t *__restrict__ ws, int64_t ws_words_per_cta) {
    extern __shared__ uint32_t smem[];
    const SelectSmem L = select_layout(max_n_segs, max_capacity, max_n_frames, BITS_IN_SMEM);
    int *dp0 = reinterpret_cast<int *>(smem + L.dp0);
    int *dp1 = reinterpret_cast<int *>(smem + L.dp1);
    int *sw = reinterpret_cast<int *>(smem + L.w);
    int *sp = reinterpret_cast<int *>(smem + L.p);
    int *spk = reinterpret_cast<int *>(smem + L.pk);
    float *smean = reinterpret_cast<float *>(smem + L.mean);
    int *spre = reinterpret_cast<int *>(smem + L.pre);
    uint32_t *smask = smem + L.mask;
    int *ssel = reinterpret_cast<int *>(smem + L.sel);
    uint32_t *bits = BITS_IN_SMEM ? (smem + L.bits) : (ws + (int64_t)blockIdx.x * ws_words_per_cta);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = SELECT_THREADS, NW = SELECT_THREADS / 32;

    for (int v = blockIdx.x; v < n_videos; v += gridDim.x) {
        const smz_video_desc d = desc[v];
        const int n = d.n_segs;
        const int cap = d.capacity;
        const int n_frames = d.n_frames;
        int status = 0;

        // ---- A. segment pooling (utils/eval.py:87-94) + value quantisation (knapsack.py:11-15)
        if (values_in != nullptr) {
            // stand-alone knapsack (utils/knapsack.py:5-23): values already quantised by the caller
            const int vlim = INT_MAX / (n > 0 ? n : 1);
            for (int s = tid; s < n; s += NT) {
                int val = __ldg(values_in + d.seg_off + s);
                if (val > vlim || val < -vlim) { status |= SMZ_STATUS_VALUE_RANGE; val = val > 0 ? vlim : -vlim; }
                smean[s] = (float)val;
                sp[s] = val;
                sw[s] = __ldg(nfps + d.seg_off + s);
                spk[s] = 0;
            }
            if (tid == 0) { ssel[0] = -1; ssel[1] = -1; ssel[2] = -1; }
        } else {
            FrameCursor cur;
            cur.init(scores + d.score_off, picks + d.picks_off, d.n_scores, d.n_picks, n_frames);
            if (cur.n_bound - 1 > d.n_scores + 1) status |= SMZ_STATUS_INTERVALS;
            const int vlim = INT_MAX / (n > 0 ? n : 1);
            for (int s = tid; s < n; s += NT) {
                const int start = __ldg(cps + 2 * (d.seg_off + s));
                int end = __ldg(cps + 2 * (d.seg_off + s) + 1) + 1;
                end = min(end, n_frames);
                const int len = end - start;
                cur.seek(start);
                const float sum = pw_sum<float>(cur, start, len);
                const float mean = __fdiv_rn(sum, (float)len);
                long long val = __double2ll_rz(__dmul_rn((double)mean, 1000.0));
                if (val > vlim || val < -vlim) { status |= SMZ_STATUS_VALUE_RANGE; val = val > 0 ? vlim : -vlim; }
                smean[s] = mean;
                sp[s] = (int)val;
                sw[s] = __ldg(nfps + d.seg_off + s);
                spk[s] = 0;
                if (out_mean) out_mean[d.seg_off + s] = mean;
                if (out_values) out_values[d.seg_off + s] = (int)val;
            }
            if (tid == 0) { ssel[0] = -1; ssel[1] = -1; ssel[2] = -1; }
        }
        {
            const int s1 = __syncthreads_or(status & SMZ_STATUS_VALUE_RANGE);
            const int s2 = __syncthreads_or(status & SMZ_STATUS_INTERVALS);
            status = (s1 ? SMZ_STATUS_VALUE_RANGE : 0) | (s2 ? SMZ_STATUS_INTERVALS : 0);
        }

        // ---- B. exclusive prefix of nfps (positions in the summary vector, utils/eval.py:111-122)
        {
            int carry = 0;
            for (int base = 0; base < n; base += NT) {
                const int i = base + tid;
                const int x = i < n ? sw[i] : 0;
                int incl = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                if (lane == 31) dp0[warp] = incl;  // dp rows are not live yet (rows hold >= 32 words)
                __syncthreads();
                int woff = 0;
                for (int k = 0; k < warp; k++) woff += dp0[k];
                int tile_total = 0;
                for (int k = 0; k < NW; k++) tile_total += dp0[k];
                if (i < n) spre[i] = carry + woff + incl - x;
                carry += tile_total;
                __syncthreads();
            }
            if (tid == 0) spre[n] = carry;
            const int mwords = (n_frames + 31) >> 5;
            for (int j = tid; j < mwords; j += NT) smask[j] = 0u;
        }
        __syncthreads();

        // ---- C. selection
        if (method == SMZ_METHOD_KNAPSACK) {
            if (spre[n] <= cap) {
                // KnapsackSolver::ReduceCapacities: the capacity constraint is inactive, all items in
                for (int s = tid; s < n; s += NT) spk[s] = 1;
            } else if (cap > 0 && n > 0) {
                const int ncell = cap + 1;
                const int words = (ncell + 31) >> 5;
                int *cur = dp0, *nxt = dp1;
                for (int c = tid; c < ncell; c += NT) cur[c] = 0;
                __syncthreads();
                // forward DP, KnapsackDynamicProgrammingSolver::SolveSubProblem for ALL items; the
                // strict '>' of upstream is kept per cell in the take-bit row of the item
                for (int i = 0; i < n; i++) {
                    const int w = sw[i], p = sp[i];
                    uint32_t *row = bits + (int64_t)i * words;
                    if (w > cap) {
                        for (int j = tid; j < words; j += NT) row[j] = 0u;
                        continue;
                    }
                    for (int base = warp * 32; base < ncell; base += NT) {
                        const int c = base + lane;
                        bool imp = false;
                        if (c < ncell) {
                            const int old = cur[c];
                            int val = old;
                            if (c >= w) {
                                const int cand = cur[c - w] + p;
                                if (cand > old) { val = cand; imp = true; }
                            }
                            nxt[c] = val;
                        }
                        const uint32_t b = __ballot_sync(0xffffffffu, imp);
                        if (lane == 0) row[base >> 5] = b;
                    }
                    __syncthreads();
                    int *t = cur; cur = nxt; nxt = t;
                }
                __syncthreads();
                // KnapsackDynamicProgrammingSolver::Solve extraction loop.  SolveSubProblem(c, k)
                // == highest item < k whose take bit at cell c is set, else 0 (ids[] default).
                int remaining = cap, num = n, round = 0;
                while (remaining > 0 && num > 0) {
                    const int slot = round % 3;
                    if (tid == 0) ssel[(round + 1) % 3] = -1;
                    const int widx = remaining >> 5, bit = remaining & 31;
                    int local = -1;
                    for (int i = tid; i < num; i += NT)
                        if ((bits[(int64_t)i * words + widx] >> bit) & 1u) local = i;
                    local = warp_max(local);
                    if (lane == 0 && local >= 0) atomicMax(&ssel[slot], local);
                    __syncthreads();
                    const int sel = max(ssel[slot], 0);
                    remaining -= sw[sel];
                    num = sel;
                    if (remaining >= 0 && tid == 0) spk[sel] = 1;
                    ++round;
                }
            }
        } else {
            // utils/eval.py:100-107: descending score, ties -> higher index first (see oracle);
            // strict '<' against the budget, no early break.
            int *order = sp;  // values are not needed by 'rank'
            int *rank = dp1;  // n <= max_n_segs; dp1 is sized below to hold it
            for (int i = tid; i < n; i += NT) {
                const float mi = smean[i];
                int r = 0;
                for (int j = 0; j < n; j++) {
                    const float mj = smean[j];
                    r += (mj > mi) || (mj == mi && j > i);
                }
                rank[i] = r;
            }
            __syncthreads();
            for (int i = tid; i < n; i += NT) order[rank[i]] = i;
            __syncthreads();
            if (tid == 0) {
                long long total = 0;
                for (int r = 0; r < n; r++) {
                    const int i = order[r];
                    if (total + sw[i] < (long long)cap) { spk[i] = 1; total += sw[i]; }
                }
            }
        }
        __syncthreads();

        // ---- D. outputs: picked flags, float summary vector, bit mask (truncated to n_frames)
        for (int s = tid; s < n; s += NT) {
            const int pk = spk[s];
            out_picked[d.seg_off + s] = (uint8_t)pk;
            if (pk) {
                const int a = spre[s];
                const int b = min(a + sw[s], n_frames);
                if (a < b) {
                    const int wa = a >> 5, wb = (b - 1) >> 5;
                    for (int wd = wa; wd <= wb; wd++) {
                        const int lo = max(a, wd << 5) & 31;
                        const int hi = min(b, (wd + 1) << 5) - (wd << 5);  // 1..32
                        const uint32_t m = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                        atomicOr(&smask[wd], m);
                    }
                }
            }
        }
        if (out_summary) {
            float *dst = out_summary + d.summ_off;
            for (int s = warp; s < n; s += NW) {
                const int a = spre[s], nf = sw[s];
                const float val = spk[s] ? 1.f : 0.f;
                for (int j = lane; j < nf; j += 32) dst[a + j] = val;
            }
        }
        __syncthreads();
        {
            const int mwords = (n_frames + 31) >> 5;
            int cnt = 0;
            for (int j = tid; j < mwords; j += NT) {
                const uint32_t m = smask[j];
                out_mask[d.mask_off + j] = m;
                cnt += __popc(m);
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) dp1[warp] = cnt;
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int k = 0; k < NW; k++) tot += dp1[k];
                out_msum[v] = tot;
                if (out_status) out_status[v] = status;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// fscore kernels
// ------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(FSCORE_THREADS)
fscore_kernel(const smz_video_desc *__restrict__ desc, int v0, const float *__restrict__ user,
              const uint32_t *__restrict__ mask, int32_t *__restrict__ overlap, int32_t *__restrict__ gsum) {
    __shared__ uint32_t s_cnt[FSCORE_MAX_USERS];
    const int v = v0 + blockIdx.y;
    const smz_video_desc d = desc[v];
    const int n_frames = d.n_frames;
    const int f_base = blockIdx.x * SMZ_FSCORE_CHUNK;
    if (f_base >= n_frames) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_users = d.n_users;
    for (int u = tid; u < n_users; u += FSCORE_THREADS) s_cnt[u] = 0u;
    __syncthreads();

    const int fa = f_base + 4 * tid;
    const int fb = fa + 4 * FSCORE_THREADS;
    const bool in_a = fa < n_frames, in_b = fb < n_frames;
    const uint32_t *vm = mask + d.mask_off;
    const uint32_t ma = in_a ? ((__ldg(vm + (fa >> 5)) >> (fa & 31)) & 0xfu) : 0u;
    const uint32_t mb = in_b ? ((__ldg(vm + (fb >> 5)) >> (fb & 31)) & 0xfu) : 0u;
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
                // low half: overlap, high half: annotator ones (per-CTA totals <= 2048 < 2^16)
                uint32_t packed = (uint32_t)(__popc(ga & ma) + __popc(gb & mb)) |
                                  ((uint32_t)(__popc(ga) + __popc(gb)) << 16);
                packed = __reduce_add_sync(0xffffffffu, packed);
                if (lane == 0 && packed) atomicAdd(&s_cnt[u0 + k], packed);
            }
        }
    }
    __syncthreads();
    for (int u = tid; u < n_users; u += FSCORE_THREADS) {
        const uint32_t c = s_cnt[u];
        if (c & 0xffffu) atomicAdd(overlap + d.ucount_off + u, (int)(c & 0xffffu));
        if (c >> 16) atomicAdd(gsum + d.ucount_off + u, (int)(c >> 16));
    }
}

// utils/eval.py:151-164 in float32 (numpy 2 / NEP 50 semantics when no zero padding happened)
__global__ void fscore_final_kernel(const smz_video_desc *__restrict__ desc, int n_videos,
                                    const int32_t *__restrict__ msum, const int32_t *__restrict__ overlap,
                                    const int32_t *__restrict__ gsum, float *__restrict__ f,
                                    double *__restrict__ avg_f, double *__restrict__ max_f) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_videos) return;
    const smz_video_desc d = desc[v];
    const float ms = __fadd_rn((float)msum[v], 1e-8f);
    float mx = 0.f;
    bool any_zero = false;  // a Python-float 0. entry promotes the reference's list to float64
    float *fv = f + d.ucount_off;
    for (int u = 0; u < d.n_users; u++) {
        const float ov = (float)overlap[d.ucount_off + u];
        const float gs = __fadd_rn((float)gsum[d.ucount_off + u], 1e-8f);
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
        if (d.n_users <= 0) avg_f[v] = 0.;
        else if (any_zero) {
            ArrayCursorF64 c64{fv};
            avg_f[v] = __ddiv_rn(pw_sum<double>(c64, 0, d.n_users), (double)d.n_users);
        }
        else {
            ArrayCursor cur{fv};
            avg_f[v] = (double)__fdiv_rn(pw_sum<float>(cur, 0, d.n_users), (float)d.n_users);
        }
    }
    if (max_f) max_f[v] = (double)mx;
}

// utils/eval.py:136-145: binarise, truncate / zero-pad to n_frames, pack 32 frames per word
__global__ void pack_summary_kernel(const smz_video_desc *__restrict__ desc, const float *__restrict__ machine,
                                    uint32_t *__restrict__ mask, int32_t *__restrict__ msum) {
    const int v = blockIdx.y;
    const smz_video_desc d = desc[v];
    const int mwords = (d.n_frames + 31) >> 5;
    const int lim = min(d.n_frames, d.summ_len);
    const float *m = machine + d.summ_off;
    int cnt = 0;
    for (int wd = blockIdx.x * blockDim.x / 32 + (threadIdx.x >> 5); wd < mwords; wd += gridDim.x * blockDim.x / 32) {
        const int fidx = (wd << 5) + (threadIdx.x & 31);
        const bool on = fidx < lim && m[fidx] > 0.f;
        const uint32_t b = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0) { mask[d.mask_off + wd] = b; cnt += __popc(b); }
    }
    if (cnt) atomicAdd(msum + v, cnt);
}

__global__ void upsample_kernel(const smz_video_desc *__restrict__ desc, const float *__restrict__ scores,
                                const int32_t *__restrict__ picks, float *__restrict__ out,
                                int32_t *__restrict__ status) {
    const int v = blockIdx.y;
    const smz_video_desc d = desc[v];
    FrameCursor cur;
    cur.init(scores + d.score_off, picks + d.picks_off, d.n_scores, d.n_picks, d.n_frames);
    if (status && blockIdx.x == 0 && threadIdx.x == 0)
        status[v] = (cur.n_bound - 1 > d.n_scores + 1) ? SMZ_STATUS_INTERVALS : 0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < d.n_frames; f += gridDim.x * blockDim.x) {
        cur.seek(f);
        out[d.frame_off + f] = cur.at(f);
    }
}

int select_smem_bytes(int max_n_segs, int max_capacity, int max_n_frames, bool bits_in_smem) {
    return select_layout(max_n_segs, max_capacity, max_n_frames, bits_in_smem).total * 4;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static int select_plan(int max_n_segs, int max_capacity, int max_n_frames, bool *bits_in_smem, int *smem_bytes) {
    if (max_n_segs < 0 || max_capacity < 0 || max_n_frames < 0)
        return smz::fail(SMZ_ERR_ARG, "negative batch maxima");
    const int optin = smz::max_smem_optin();
    int with_bits = select_smem_bytes(max_n_segs, max_capacity, max_n_frames, true);
    int without = select_smem_bytes(max_n_segs, max_capacity, max_n_frames, false);
    if (with_bits <= optin) { *bits_in_smem = true; *smem_bytes = with_bits; return SMZ_OK; }
    if (without <= optin) { *bits_in_smem = false; *smem_bytes = without; return SMZ_OK; }
    return smz::fail(SMZ_ERR_UNSUPPORTED,
                     "select_shots: DP rows for capacity %d, %d segments and %d frames need %d B of shared "
                     "memory (> %d B per CTA)", max_capacity, max_n_segs, max_n_frames, without, optin);
}

static int select_grid(int n_videos, bool bits_in_smem, int smem_bytes, int *grid) {
    int per_sm = 0;
    if (bits_in_smem) {
        SMZ_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        SMZ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_kernel<true>, SELECT_THREADS, smem_bytes));
    } else {
        SMZ_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        SMZ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_kernel<false>, SELECT_THREADS, smem_bytes));
    }
    if (per_sm < 1) per_sm = 1;
    int g = smz::sm_count() * per_sm;
    *grid = n_videos < g ? n_videos : g;
    return SMZ_OK;
}

extern "C" int smz_select_workspace_bytes(int n_videos, int max_n_segs, int max_capacity, int max_n_frames,
                                          int64_t *bytes) {
    SMZ_REQUIRE(bytes != nullptr, "bytes is NULL");
    *bytes = 0;
    if (n_videos <= 0) return SMZ_OK;
    bool in_smem; int smem_bytes;
    int rc = select_plan(max_n_segs, max_capacity, max_n_frames, &in_smem, &smem_bytes);
    if (rc != SMZ_OK) return rc;
    if (in_smem) return SMZ_OK;
    int grid;
    rc = select_grid(n_videos, in_smem, smem_bytes, &grid);
    if (rc != SMZ_OK) return rc;
    *bytes = (int64_t)grid * max_n_segs * ((max_capacity + 1 + 31) / 32) * 4;
    return SMZ_OK;
}

static int select_launch(const smz_video_desc *desc, int n_videos, const float *scores, const int32_t *picks,
                         const int32_t *cps, const int32_t *nfps, const int32_t *values_in, int method,
                         int max_n_segs, int max_capacity, int max_n_frames, float *seg_mean, int32_t *values,
                         uint8_t *picked, float *summary, uint32_t *mask, int32_t *msum, int32_t *status,
                         void *ws, int64_t ws_bytes, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0, "n_videos < 0");
    SMZ_REQUIRE(desc && nfps, "NULL input pointer");
    SMZ_REQUIRE(picked && mask && msum, "picked, mask and msum outputs are required");
    SMZ_REQUIRE(method == SMZ_METHOD_KNAPSACK || method == SMZ_METHOD_RANK, "unknown method %d", method);
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    bool in_smem; int smem_bytes, grid;
    rc = select_plan(max_n_segs, max_capacity, max_n_frames, &in_smem, &smem_bytes);
    if (rc != SMZ_OK) return rc;
    rc = select_grid(n_videos, in_smem, smem_bytes, &grid);
    if (rc != SMZ_OK) return rc;
    const int64_t words_per_cta = (int64_t)max_n_segs * ((max_capacity + 1 + 31) / 32);
    if (!in_smem) {
        SMZ_REQUIRE(ws != nullptr && ws_bytes >= (int64_t)grid * words_per_cta * 4,
                    "work buffer too small: need %lld bytes", (long long)((int64_t)grid * words_per_cta * 4));
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (in_smem)
        select_kernel<true><<<grid, SELECT_THREADS, smem_bytes, st>>>(
            desc, n_videos, scores, picks, cps, nfps, values_in, method, max_n_segs, max_capacity, max_n_frames,
            seg_mean, values, picked, summary, mask, msum, status, (uint32_t *)ws, words_per_cta);
    else
        select_kernel<false><<<grid, SELECT_THREADS, smem_bytes, st>>>(
            desc, n_videos, scores, picks, cps, nfps, values_in, method, max_n_segs, max_capacity, max_n_frames,
            seg_mean, values, picked, summary, mask, msum, status, (uint32_t *)ws, words_per_cta);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_select_shots(const smz_video_desc *desc, int n_videos, const float *scores,
                                const int32_t *picks, const int32_t *cps, const int32_t *nfps, int method,
                                int max_n_segs, int max_capacity, int max_n_frames, float *seg_mean,
                                int32_t *values, uint8_t *picked, float *summary, uint32_t *mask, int32_t *msum,
                                int32_t *status, void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(n_videos == 0 || (scores && picks && cps), "NULL input pointer");
    return select_launch(desc, n_videos, scores, picks, cps, nfps, nullptr, method, max_n_segs, max_capacity,
                         max_n_frames, seg_mean, values, picked, summary, mask, msum, status, ws, ws_bytes, stream);
}

extern "C" int smz_knapsack(const smz_video_desc *desc, int n_videos, const int32_t *values, const int32_t *nfps,
                            int max_n_segs, int max_capacity, int max_n_frames, uint8_t *picked, uint32_t *mask,
                            int32_t *msum, int32_t *status, void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(n_videos == 0 || values, "NULL values pointer");
    return select_launch(desc, n_videos, nullptr, nullptr, nullptr, nfps, values, SMZ_METHOD_KNAPSACK, max_n_segs,
                         max_capacity, max_n_frames, nullptr, nullptr, picked, nullptr, mask, msum, status, ws,
                         ws_bytes, stream);
}

extern "C" int smz_fscore(const smz_video_desc *desc, int n_videos, int max_n_frames, int total_users,
                          const float *user_summary, const uint32_t *mask, const int32_t *msum,
                          int32_t *overlap, int32_t *gsum, float *f, double *avg_f, double *max_f, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && max_n_frames >= 0 && total_users >= 0, "negative size");
    SMZ_REQUIRE(desc && user_summary && mask && msum && overlap && gsum && f, "NULL pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SMZ_CUDA_CHECK(cudaMemsetAsync(overlap, 0, sizeof(int32_t) * (size_t)total_users, st));
    SMZ_CUDA_CHECK(cudaMemsetAsync(gsum, 0, sizeof(int32_t) * (size_t)total_users, st));
    const int chunks = (max_n_frames + SMZ_FSCORE_CHUNK - 1) / SMZ_FSCORE_CHUNK;
    if (chunks > 0) {
        for (int v0 = 0; v0 < n_videos; v0 += 65535) {
            const int nv = n_videos - v0 < 65535 ? n_videos - v0 : 65535;
            fscore_kernel<<<dim3(chunks, nv), FSCORE_THREADS, 0, st>>>(desc, v0, user_summary, mask, overlap, gsum);
        }
        SMZ_CUDA_CHECK(cudaGetLastError());
    }
    fscore_final_kernel<<<(n_videos + 127) / 128, 128, 0, st>>>(desc, n_videos, msum, overlap, gsum, f, avg_f, max_f);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_pack_summary(const smz_video_desc *desc, int n_videos, int max_n_frames, const float *machine,
                                uint32_t *mask, int32_t *msum, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && n_videos <= 65535, "pack_summary: 1..65535 videos per call");
    SMZ_REQUIRE(desc && machine && mask && msum, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SMZ_CUDA_CHECK(cudaMemsetAsync(msum, 0, sizeof(int32_t) * (size_t)n_videos, st));
    const int mwords = (max_n_frames + 31) / 32;
    int gx = (mwords + 7) / 8;  // 8 warps per CTA, one word per warp per step
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    pack_summary_kernel<<<dim3(gx, n_videos), 256, 0, st>>>(desc, machine, mask, msum);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

extern "C" int smz_upsample(const smz_video_desc *desc, int n_videos, int max_n_frames, const float *scores,
                            const int32_t *picks, float *frame_scores, int32_t *status, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && n_videos <= 65535, "upsample: 1..65535 videos per call");
    SMZ_REQUIRE(desc && scores && picks && frame_scores, "NULL pointer");
    int gx = (max_n_frames + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 256) gx = 256;
    upsample_kernel<<<dim3(gx, n_videos), 256, 0, (cudaStream_t)stream>>>(desc, scores, picks, frame_scores, status);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
