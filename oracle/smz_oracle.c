/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — plain-C restatement of the reference's
 * summary-generation / F-score path, used (a) as the parity checker at sizes where the
 * numpy/pure-Python restatement is too slow and (b) as the timed CPU baseline in bench.py.
 * The product (summarizer_b200) never links or loads this file.
 *
 * Reference lines followed (relative to /root/reference/summarizer/):
 *   upsample            utils/eval.py:15-35
 *   segment mean        utils/eval.py:87-94   (numpy float32 pairwise sum, loops_utils.h.src)
 *   capacity            utils/eval.py:96
 *   value quantisation  utils/knapsack.py:11-15
 *   knapsack            utils/knapsack.py:7-21 -> OR-tools 7.5.7466 knapsack_solver.cc
 *                       (third party, NOT under /root/reference: restated, PARITY UNPINNED)
 *   rank selection      utils/eval.py:100-107
 *   summary vector      utils/eval.py:111-122
 *   F-score             utils/eval.py:125-165
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math  (see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PW_BLOCKSIZE 128

/* numpy FLOAT_pairwise_sum over a[0..n) given through an accessor (so the upsampled
 * frame-score vector never has to be materialised when the caller does not want it). */
static float pairwise_sum(const float *a, int64_t n) {
    if (n < 8) {
        float res = 0.f;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= PW_BLOCKSIZE) {
        float r[8], res;
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

static double pairwise_sum_f64(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= PW_BLOCKSIZE) {
        double r[8], res;
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum_f64(a, n2) + pairwise_sum_f64(a + n2, n - n2);
    }
}

float smzo_mean_f32(const float *a, int64_t n) { return pairwise_sum(a, n) / (float)n; }

/* utils/eval.py:15-35.  positions sorted ascending, int32.  Returns 0, or -1 when the
 * reference would raise IndexError (more intervals than scores+1). */
int smzo_upsample(const float *scores, int64_t n_scores, const int32_t *positions, int64_t n_pos,
                  int64_t n_frames, float *out) {
    memset(out, 0, sizeof(float) * (size_t)n_frames);
    int64_t n_bound = n_pos + (positions[n_pos - 1] != n_frames ? 1 : 0);
    int64_t n_int = n_bound - 1;
    if (n_int > n_scores + 1) return -1;
    for (int64_t i = 0; i < n_int; i++) {
        int64_t lo = positions[i];
        int64_t hi = (i + 1 < n_pos) ? positions[i + 1] : n_frames;
        if (hi > n_frames) hi = n_frames;
        float v = (i == n_scores) ? 0.f : scores[i];
        for (int64_t f = lo; f < hi; f++) out[f] = v;
    }
    return 0;
}

int64_t smzo_capacity(int64_t n_frames, double proportion) {
    return (int64_t)floor((double)n_frames * proportion);
}

/* OR-tools KnapsackDynamicProgrammingSolver::SolveSubProblem */
static int solve_subproblem(const int64_t *profits, const int64_t *weights, int64_t capacity,
                            int num_items, int64_t *prof, int32_t *ids) {
    for (int64_t c = 0; c <= capacity; c++) { prof[c] = 0; ids[c] = 0; }
    for (int item = 0; item < num_items; item++) {
        const int64_t w = weights[item], p = profits[item];
        for (int64_t c = capacity; c >= w; --c) {
            if (prof[c - w] + p > prof[c]) {
                prof[c] = prof[c - w] + p;
                ids[c] = item;
            }
        }
    }
    return ids[capacity];
}

/* KnapsackSolver::Init (ReduceCapacities) + KnapsackDynamicProgrammingSolver::Solve, literal:
 * the DP is re-run once per extracted item exactly as upstream does.  picked[n] <- 0/1. */
int smzo_knapsack(const int64_t *profits, const int64_t *weights, int n, int64_t capacity,
                  uint8_t *picked) {
    int64_t sumw = 0;
    for (int i = 0; i < n; i++) { sumw += weights[i]; picked[i] = 0; }
    if (sumw <= capacity) { for (int i = 0; i < n; i++) picked[i] = 1; return 0; }
    if (capacity <= 0 || n == 0) return 0;
    int64_t *prof = (int64_t *)malloc(sizeof(int64_t) * (size_t)(capacity + 1));
    int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(capacity + 1));
    if (!prof || !ids) { free(prof); free(ids); return -2; }
    int64_t remaining = capacity;
    int num_items = n;
    while (remaining > 0 && num_items > 0) {
        int s = solve_subproblem(profits, weights, remaining, num_items, prof, ids);
        remaining -= weights[s];
        num_items = s;
        if (remaining >= 0) picked[s] = 1;
    }
    free(prof); free(ids);
    return 0;
}

/* utils/eval.py:100-107 with ties fixed as "stable ascending argsort, reversed". */
static void rank_select(const double *seg, const int32_t *nfps, int n, int64_t limits, uint8_t *picked) {
    int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) order[i] = i;
    /* insertion sort, stable ascending */
    for (int i = 1; i < n; i++) {
        int k = order[i]; int j = i - 1;
        while (j >= 0 && seg[order[j]] > seg[k]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = k;
    }
    int64_t total = 0;
    for (int i = 0; i < n; i++) picked[i] = 0;
    for (int r = n - 1; r >= 0; r--) {
        int i = order[r];
        if (total + nfps[i] < limits) { picked[i] = 1; total += nfps[i]; }
    }
    free(order);
}

/* utils/eval.py:74-123.  method: 0 = knapsack, 1 = rank.
 * Outputs (any may be NULL): seg_mean[n_segs] (float32 means), values[n_segs] (int64),
 * picked[n_segs], summary[sum(nfps)].  Returns capacity (>=0) or <0 on error. */
int64_t smzo_generate_summary(const float *scores, int64_t n_scores, const int32_t *positions,
                              int64_t n_pos, const int32_t *cps, int n_segs, const int32_t *nfps,
                              int64_t n_frames, double proportion, int method,
                              float *seg_mean, int64_t *values, uint8_t *picked, float *summary) {
    float *frame = (float *)malloc(sizeof(float) * (size_t)(n_frames > 0 ? n_frames : 1));
    double *seg = (double *)malloc(sizeof(double) * (size_t)(n_segs > 0 ? n_segs : 1));
    int64_t *vals = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_segs > 0 ? n_segs : 1));
    int64_t *w = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_segs > 0 ? n_segs : 1));
    uint8_t *pk = (uint8_t *)malloc((size_t)(n_segs > 0 ? n_segs : 1));
    int64_t rc = -2;
    if (!frame || !seg || !vals || !w || !pk) goto done;
    if (smzo_upsample(scores, n_scores, positions, n_pos, n_frames, frame) != 0) { rc = -1; goto done; }
    for (int s = 0; s < n_segs; s++) {
        int64_t start = cps[2 * s], end = (int64_t)cps[2 * s + 1] + 1;
        if (end > n_frames) end = n_frames;
        float m = smzo_mean_f32(frame + start, end - start);
        if (seg_mean) seg_mean[s] = m;
        seg[s] = (double)m;
        vals[s] = (int64_t)(seg[s] * 1000.0);       /* float64 product, truncation toward 0 */
        w[s] = nfps[s];
        if (values) values[s] = vals[s];
    }
    int64_t limits = smzo_capacity(n_frames, proportion);
    if (method == 0) {
        if (smzo_knapsack(vals, w, n_segs, limits, pk) != 0) goto done;
    } else {
        rank_select(seg, nfps, n_segs, limits, pk);
    }
    if (picked) memcpy(picked, pk, (size_t)n_segs);
    if (summary) {
        int64_t pos = 0;
        for (int s = 0; s < n_segs; s++) {
            float v = pk[s] ? 1.f : 0.f;
            for (int32_t j = 0; j < nfps[s]; j++) summary[pos + j] = v;
            pos += nfps[s];
        }
    }
    rc = limits;
done:
    free(frame); free(seg); free(vals); free(w); free(pk);
    return rc;
}

/* utils/eval.py:125-165.  machine summary of length m_len is binarised (>0), truncated or
 * zero-padded to n_frames.  counts: overlap[u], gsum[u], *msum.  f[u] in float32 (the dtype
 * numpy 2 computes in when no padding happens).  avg/max: float32 pairwise mean widened to double,
 * unless some user has overlap 0 — then the reference's list holds a Python float 0. and
 * np.mean/np.max run in float64 over the float32-valued entries (utils/eval.py:156-164).
 * A summary shorter than n_frames is padded with float64 zeros (utils/eval.py:143-145): everything but
 * gt_summary.sum() + 1e-8 is float64 then (f[u] holds the float32 rounding of the float64 F).
 * Returns 0. */
int smzo_evaluate_summary(const float *machine, int64_t m_len, const float *user, int n_users,
                          int64_t n_frames, int64_t user_ld, int32_t *overlap, int32_t *gsum,
                          int32_t *msum, float *f, double *avg_f, double *max_f) {
    int64_t lim = m_len < n_frames ? m_len : n_frames;
    int64_t ms = 0;
    for (int64_t i = 0; i < lim; i++) ms += machine[i] > 0.f;
    if (msum) *msum = (int32_t)ms;
    float fs[4096];
    int any_zero = 0;
    const int padded = m_len < n_frames;   /* utils/eval.py:143-145: np.zeros (float64) padding promotes the summary */
    double fd[4096];
    fs[0] = 0.f; fd[0] = 0.0;
    if (n_users > 4096) return -1;
    for (int u = 0; u < n_users; u++) {
        const float *g = user + (int64_t)u * user_ld;
        int64_t ov = 0, gs = 0;
        for (int64_t i = 0; i < n_frames; i++) {
            int gi = g[i] > 0.f;
            gs += gi;
            if (i < lim) ov += gi & (machine[i] > 0.f);
        }
        if (overlap) overlap[u] = (int32_t)ov;
        if (gsum) gsum[u] = (int32_t)gs;
        if (padded) {   /* float64 overlap / precision / recall / F; gt_summary.sum() + 1e-8 still float32 */
            double dov = (double)ov;
            double dp = dov / ((double)ms + 1e-8);
            double dr = dov / (double)((float)gs + 1e-8f);
            fd[u] = (dp == 0.0 && dr == 0.0) ? 0.0 : ((2.0 * dp) * dr) / (dp + dr);
            fs[u] = (float)fd[u];
            if (f) f[u] = fs[u];
            continue;
        }
        float fov = (float)ov;
        float precision = fov / ((float)ms + 1e-8f);
        float recall = fov / ((float)gs + 1e-8f);
        float fsc = 0.f;
        if (!(precision == 0.f && recall == 0.f)) fsc = ((2.f * precision) * recall) / (precision + recall);
        else any_zero = 1;
        fs[u] = fsc;
        if (f) f[u] = fsc;
    }
    if (padded) {
        if (avg_f) *avg_f = n_users <= 0 ? 0.0 : pairwise_sum_f64(fd, n_users) / (double)n_users;
        if (max_f) { double m = fd[0]; for (int u = 1; u < n_users; u++) if (fd[u] > m) m = fd[u]; *max_f = m; }
        return 0;
    }
    if (avg_f) {
        if (n_users <= 0) *avg_f = 0.0;
        else if (any_zero) {
            double ds[4096];
            for (int u = 0; u < n_users; u++) ds[u] = (double)fs[u];
            *avg_f = pairwise_sum_f64(ds, n_users) / (double)n_users;
        } else *avg_f = (double)smzo_mean_f32(fs, n_users);
    }
    if (max_f) { float m = fs[0]; for (int u = 1; u < n_users; u++) if (fs[u] > m) m = fs[u]; *max_f = (double)m; }
    return 0;
}
