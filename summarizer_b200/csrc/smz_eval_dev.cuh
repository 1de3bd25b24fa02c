// Device helpers shared by the selection and F-score kernels: the upsample cursor
// (utils/eval.py:15-35) and numpy's pairwise summation order (loops_utils.h.src).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace smzdev {

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

struct SmemFrames {   // upsample staged in shared memory as the score index of every frame (0xffff = zero score)
    const unsigned short *idx;
    const float *scores;
    __device__ __forceinline__ float at(int f) const {
        const unsigned i = idx[f];
        return i == 0xffffu ? 0.f : __ldg(scores + i);
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

// ---- 8-lane cooperative version of the same summation tree (segment pooling) ----------------
// Lane j of a group owns accumulator r[j] of numpy's 8-way unrolled block; the combine
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is done with xor-shuffles (fp add is commutative, so every
// lane ends with the same bits); the n%8 tail and n<8 blocks are summed by lane 0 and broadcast.
// All 8 lanes of a group take the same control flow; gmask names exactly those lanes.
template <class Cur>
static __device__ float pw_block_group(Cur &cur, int start, int n, int gl, unsigned gmask, int lane0) {
    float res;
    if (n < 8) {
        res = 0.f;
        if (gl == 0)
            for (int i = 0; i < n; i++) res = __fadd_rn(res, cur.at(start + i));
        return __shfl_sync(gmask, res, lane0);
    }
    float r = cur.at(start + gl);
    const int lim = n - (n % 8);
    for (int i = 8; i < lim; i += 8) r = __fadd_rn(r, cur.at(start + i + gl));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    if (gl == 0)
        for (int i = lim; i < n; i++) r = __fadd_rn(r, cur.at(start + i));
    return __shfl_sync(gmask, r, lane0);
}

// generic depth: explicit stack (local memory), kept out of line
template <class Cur>
static __device__ __noinline__ float pw_sum_group_deep(Cur &cur, int start, int n, int gl, unsigned gmask, int lane0) {
    int s_start[32], s_n[32];
    float s_left[32];
    unsigned char s_stage[32];
    int sp = 0;
    float ret = 0.f;
    s_start[0] = start; s_n[0] = n; s_stage[0] = 0;
    while (sp >= 0) {
        const int n_ = s_n[sp];
        if (s_stage[sp] == 0) {
            if (n_ <= 128) { ret = pw_block_group(cur, s_start[sp], n_, gl, gmask, lane0); --sp; continue; }
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
            ret = __fadd_rn(s_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

// the first DEPTH levels of the recursion are unrolled in registers (segments up to ~128 * 2^DEPTH frames never
// touch the stack); left before right, so a forward-only cursor keeps moving forward
template <int DEPTH, class Cur>
static __device__ __forceinline__ float pw_sum_group_rec(Cur &cur, int start, int n, int gl, unsigned gmask, int lane0) {
    if (n <= 128) return pw_block_group(cur, start, n, gl, gmask, lane0);
    if constexpr (DEPTH == 0) {
        return pw_sum_group_deep(cur, start, n, gl, gmask, lane0);
    } else {
        int n2 = n / 2; n2 -= n2 % 8;
        const float l = pw_sum_group_rec<DEPTH - 1>(cur, start, n2, gl, gmask, lane0);
        const float r = pw_sum_group_rec<DEPTH - 1>(cur, start + n2, n - n2, gl, gmask, lane0);
        return __fadd_rn(l, r);
    }
}

template <class Cur>
static __device__ float pw_sum_group(Cur &cur, int start, int n, int gl, unsigned gmask, int lane0) {
    return pw_sum_group_rec<3>(cur, start, n, gl, gmask, lane0);
}

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


}  // namespace smzdev
