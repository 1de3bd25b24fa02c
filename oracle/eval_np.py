"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy / pure-Python restatement of the reference's summary-generation and
F-score path.  Every function names the reference lines it follows
(paths relative to /root/reference/summarizer/).

The restatement is written as *explicit scalar arithmetic* (not "call numpy and
hope"), because the CUDA kernels have to reproduce the same rounding sequence:

  * segment score  = numpy float32 pairwise-sum mean            (utils/eval.py:91-94)
  * knapsack value = trunc(float64(score) * 1000.0)             (utils/knapsack.py:11-15)
  * capacity       = int(floor(float64(n_frames) * proportion)) (utils/eval.py:96)
  * selection      = OR-tools 7.5 KNAPSACK_DYNAMIC_PROGRAMMING_SOLVER
                     (third party, restated; **parity unpinned** under ties)
  * summary vector = ones/zeros(nfps[s]) concatenated           (utils/eval.py:111-122)
  * F-score        = float32 ratios of exact integer counts     (utils/eval.py:125-165)
"""
import math

import numpy as np

F32 = np.float32
PW_BLOCKSIZE = 128  # numpy/core/src/umath/loops_utils.h.src


# --------------------------------------------------------------------------------------
# numpy float32 pairwise summation (what ndarray.mean()/sum() does for a contiguous
# float32 vector).  Verified bit-for-bit against np.add.reduce in tests/test_oracle.py.
# --------------------------------------------------------------------------------------
def pairwise_sum_f32(a):
    """Scalar emulation of numpy's FLOAT_pairwise_sum on a 1-D float32 array."""
    a = np.asarray(a, dtype=F32)
    n = a.shape[0]
    if n < 8:
        res = F32(0.0)
        for i in range(n):
            res = F32(res + a[i])
        return res
    if n <= PW_BLOCKSIZE:
        r = [F32(a[j]) for j in range(8)]
        i = 8
        lim = n - (n % 8)
        while i < lim:
            for j in range(8):
                r[j] = F32(r[j] + a[i + j])
            i += 8
        res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) +
                  F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
        while i < n:
            res = F32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return F32(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def mean_f32(a):
    """ndarray.mean() of a float32 vector: pairwise sum, then ONE float32 divide
    (numpy/core/_methods.py:_mean -> umr_sum, then ret.dtype.type(ret / rcount))."""
    a = np.asarray(a, dtype=F32)
    return F32(pairwise_sum_f32(a) / F32(a.shape[0]))


# --------------------------------------------------------------------------------------
# utils/eval.py:15-35  upsample
# --------------------------------------------------------------------------------------
def upsample_bounds(n_frames, positions):
    """Interval boundaries used by upsample (utils/eval.py:25-28): positions cast to
    int32 unless already int64, ``n_frames`` appended unless already last."""
    positions = np.asarray(positions)
    if positions.dtype != np.int64:
        positions = positions.astype(np.int32)
    if positions[-1] != n_frames:
        positions = np.concatenate([positions, [n_frames]])
    return positions.astype(np.int64)


def upsample(scores, n_frames, positions):
    """utils/eval.py:15-35 — piecewise-constant expansion to n_frames."""
    scores = np.asarray(scores)
    n_frames = int(n_frames)
    b = upsample_bounds(n_frames, positions)
    out = np.zeros(n_frames, dtype=F32)
    n_int = len(b) - 1
    if n_int > len(scores) + 1:
        raise IndexError("more upsample intervals than scores (reference raises IndexError)")
    for i in range(n_int):
        lo, hi = int(b[i]), int(b[i + 1])
        out[lo:hi] = 0 if i == len(scores) else scores[i]
    return out


# --------------------------------------------------------------------------------------
# utils/eval.py:87-96  segment pooling + capacity
# --------------------------------------------------------------------------------------
def segment_scores(frame_scores, cps):
    """utils/eval.py:90-94 — float(frame_scores[start:end+1].mean()) per segment.
    Returns float64 array holding the widened float32 means (Python floats)."""
    cps = np.asarray(cps)
    out = np.empty(cps.shape[0], dtype=np.float64)
    for s in range(cps.shape[0]):
        start, end = int(cps[s, 0]), int(cps[s, 1] + 1)
        out[s] = float(mean_f32(frame_scores[start:end]))
    return out


def capacity_of(n_frames, proportion):
    """utils/eval.py:96 — int(math.floor(n_frames * proportion)) in float64."""
    return int(math.floor(float(int(n_frames)) * float(proportion)))


def knapsack_values(seg_score):
    """utils/knapsack.py:11-14 — (np.array(values) * 1000).astype(int): float64 product,
    truncation toward zero."""
    return np.trunc(np.asarray(seg_score, dtype=np.float64) * 1000.0).astype(np.int64)


# --------------------------------------------------------------------------------------
# OR-tools 7.5.7466 KnapsackSolver(KNAPSACK_DYNAMIC_PROGRAMMING_SOLVER) — THIRD PARTY.
# Source: ortools/algorithms/knapsack_solver.cc (not under /root/reference; pinned by
# summarizer/requirements.txt:11).  Restated from the published algorithm; anchored on
# the reference call site utils/knapsack.py:5-23.  PARITY UNPINNED under ties.
#
#   KnapsackSolver::Init (use_reduction_ = true by default):
#     ReduceCapacities(): if sum(weights) <= capacity the single dimension is inactive,
#       every item is fixed IN and the DP never runs.
#     ReduceProblem(): uses GetLowerAndUpperBoundWhenItem, which only the branch-and-bound
#       solver overrides (base class answers [0, +inf)), so it reduces nothing for the DP.
#   KnapsackDynamicProgrammingSolver::SolveSubProblem(capacity, num_items):
#     profits/ids zeroed on [0, capacity]; for item in [0, num_items): for c from capacity
#     down to w[item]: if profit[c-w]+p > profit[c] (STRICT) then update and ids[c] = item.
#     returns ids[capacity]   (0 when nothing ever improved that cell).
#   ::Solve(): remaining = capacity; n = num_items;
#     while remaining > 0 and n > 0:
#        s = SolveSubProblem(remaining, n); remaining -= w[s]; n = s;
#        if remaining >= 0: best[s] = true
# --------------------------------------------------------------------------------------
def _solve_subproblem(profits, weights, capacity, num_items):
    prof = [0] * (capacity + 1)
    ids = [0] * (capacity + 1)
    for item in range(num_items):
        w = weights[item]
        p = profits[item]
        c = capacity
        while c >= w:
            cand = prof[c - w] + p
            if cand > prof[c]:
                prof[c] = cand
                ids[c] = item
            c -= 1
    return ids[capacity]


def knapsack_dp_ortools(profits, weights, capacity):
    """Literal restatement (re-solves the DP once per extracted item, like upstream).
    O(k*n*W): use for small cases; `knapsack_dp_takebits` is the equivalent fast form."""
    profits = [int(p) for p in profits]
    weights = [int(w) for w in weights]
    n = len(profits)
    capacity = int(capacity)
    if sum(weights) <= capacity:               # ReduceCapacities
        return list(range(n))
    best = [False] * n
    remaining = capacity
    num_items = n
    while remaining > 0 and num_items > 0:
        s = _solve_subproblem(profits, weights, remaining, num_items)
        remaining -= weights[s]
        num_items = s
        if remaining >= 0:
            best[s] = True
    return [i for i in range(n) if best[i]]


def knapsack_dp_takebits(profits, weights, capacity):
    """Equivalent single-pass form (what the CUDA kernel implements): one forward DP that
    records, per (item, cell), whether the item strictly improved the cell; the
    sub-problem answer at (num_items, c) is the highest item < num_items whose bit at c is
    set, else 0.  Valid because a DP restricted to capacity c' <= c and to an item prefix
    is a prefix of the full table.  Cross-checked against knapsack_dp_ortools in tests."""
    profits = np.asarray(profits, dtype=np.int64)
    weights = np.asarray(weights, dtype=np.int64)
    n = len(profits)
    capacity = int(capacity)
    if int(weights.sum()) <= capacity:
        return list(range(n))
    if capacity <= 0 or n == 0:
        return []
    prof = np.zeros(capacity + 1, dtype=np.int64)
    take = np.zeros((n, capacity + 1), dtype=bool)
    for i in range(n):
        w, p = int(weights[i]), int(profits[i])
        if w > capacity:
            continue
        if w <= 0:
            # upstream loop `for c = capacity; c >= w; --c` with w == 0 updates in place
            cand = prof + p
            imp = cand > prof
        else:
            cand = np.full(capacity + 1, np.iinfo(np.int64).min, dtype=np.int64)
            cand[w:] = prof[:capacity + 1 - w] + p
            imp = cand > prof
        prof = np.where(imp, cand, prof)
        take[i] = imp
    best = [False] * n
    remaining, num_items = capacity, n
    while remaining > 0 and num_items > 0:
        col = np.nonzero(take[:num_items, remaining])[0]
        s = int(col[-1]) if len(col) else 0
        remaining -= int(weights[s])
        num_items = s
        if remaining >= 0:
            best[s] = True
    return [i for i in range(n) if best[i]]


def knapsack_ortools(values, weights, items, capacity):
    """utils/knapsack.py:5-23 with the OR-tools call restated."""
    vals = knapsack_values(values)
    w = np.asarray(weights).astype(np.int64)
    return knapsack_dp_takebits(vals, w, int(capacity))


def rank_select(seg_score, nfps, limits):
    """utils/eval.py:100-107 — greedy by descending score, strict '<' on the budget.
    np.argsort(list)[::-1]: ties come out in an order that depends on numpy's (unstable,
    SIMD-dispatched) quicksort; this restatement fixes them as 'stable ascending sort,
    reversed' = higher index first.  Ties in 'rank' are therefore unpinned."""
    order = np.argsort(np.asarray(seg_score, dtype=np.float64), kind="stable")[::-1].tolist()
    picks, total = [], 0
    for i in order:
        if total + int(nfps[i]) < limits:
            picks.append(i)
            total += int(nfps[i])
    return picks


# --------------------------------------------------------------------------------------
# utils/eval.py:74-123  generate_summary
# --------------------------------------------------------------------------------------
def summary_vector(picks, nfps):
    """utils/eval.py:111-122 — concatenation of ones/zeros(nfps[s]) in segment order."""
    nfps = [int(x) for x in nfps]
    picked = set(int(p) for p in picks)
    out = np.zeros(int(sum(nfps)), dtype=F32)
    pos = 0
    for s, nf in enumerate(nfps):
        if s in picked:
            out[pos:pos + nf] = 1.0
        pos += nf
    return out


def generate_summary(scores, cps, n_frames, nfps, positions, proportion=0.15, method="knapsack",
                     return_parts=False):
    n_frames = int(n_frames)
    frame_scores = upsample(scores, n_frames, positions)
    seg_score = segment_scores(frame_scores, cps)
    limits = capacity_of(n_frames, proportion)
    if method == "knapsack":
        picks = knapsack_ortools(seg_score, nfps, len(seg_score), limits)
    elif method == "rank":
        picks = rank_select(seg_score, nfps, limits)
    else:
        raise KeyError(f"Unknown method {method}")
    summary = summary_vector(picks, nfps)
    if return_parts:
        return summary, dict(seg_score=seg_score, values=knapsack_values(seg_score),
                             capacity=limits, picks=sorted(picks))
    return summary


# --------------------------------------------------------------------------------------
# utils/eval.py:125-165  evaluate_summary
# --------------------------------------------------------------------------------------
def overlap_counts(machine_summary, user_summary):
    """Exact integer counts behind utils/eval.py:141-155: after binarisation (>0 -> 1) and
    pad/truncate of the machine summary to n_frames.  Returns (overlap[u], msum, gsum[u])."""
    user = np.asarray(user_summary)
    n_users, n_frames = user.shape
    m = (np.asarray(machine_summary) > 0)
    if len(m) > n_frames:
        m = m[:n_frames]
    elif len(m) < n_frames:
        m = np.concatenate([m, np.zeros(n_frames - len(m), dtype=bool)])
    g = user > 0
    overlap = (g & m[None, :]).sum(axis=1).astype(np.int64)
    return overlap, int(m.sum()), g.sum(axis=1).astype(np.int64)


def fscores_from_counts(overlap, msum, gsum, padded=False):
    """utils/eval.py:151-162 in the dtypes numpy 2 (NEP 50) runs it in.  Normally all
    float32.  When the machine summary was zero-padded (np.zeros -> float64 promotion of
    machine_summary, utils/eval.py:143-145) overlap and msum are float64 while
    ``gt_summary.sum() + 1e-8`` stays float32 before the (float64) divide."""
    out = []
    for ov, gs in zip(overlap, gsum):
        if padded:
            T = np.float64
            ov, ms = T(ov), T(msum)
            precision = ov / (ms + 1e-8)
            recall = ov / T(F32(F32(gs) + F32(1e-8)))
        else:
            T = np.float32
            ov, ms, gs = T(ov), T(msum), T(gs)
            precision = T(ov / T(ms + T(1e-8)))
            recall = T(ov / T(gs + T(1e-8)))
        if precision == 0 and recall == 0:
            out.append(T(0.0))
        else:
            out.append(T(T(T(T(2) * precision) * recall) / T(precision + recall)))
    return np.asarray(out, dtype=np.float64 if padded else np.float32)


def evaluate_summary(machine_summary, user_summary):
    """utils/eval.py:125-165 — returns (avg_f, max_f).
    dtype subtlety reproduced here: a user with precision == recall == 0 contributes the PYTHON
    float ``0.`` (utils/eval.py:156-157), so ``np.mean``/``np.max`` of the mixed list run in float64
    over the float32-valued entries; otherwise the list is all-float32 and the mean is a float32
    pairwise sum."""
    n_frames = np.asarray(user_summary).shape[1]
    overlap, msum, gsum = overlap_counts(machine_summary, user_summary)
    padded = len(machine_summary) < n_frames
    f = fscores_from_counts(overlap, msum, gsum, padded=padded)
    if padded or np.any(np.asarray(overlap) == 0):
        f = f.astype(np.float64)
    return np.mean(f), np.max(f)
