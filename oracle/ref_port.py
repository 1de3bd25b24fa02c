"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — the *timed* CPU baseline.

A port of the reference's evaluation path that keeps the reference's own cost structure: numpy
calls driven by Python loops for upsample / segment means / summary concatenation / per-user
F-score (utils/eval.py:15-35, 87-94, 111-122, 133-165), and a compiled C++-class solver for the
knapsack (utils/knapsack.py:7-21 calls OR-tools' C++ DP; here oracle/smz_oracle.c's literal
restatement, which like upstream re-solves the DP once per extracted item).  Results are identical
to oracle/eval_np.py; this module exists so bench.py times something shaped like the reference
rather than the scalar-emulation oracle."""
import math

import numpy as np

from . import c_oracle


def upsample(scores, n_frames, positions):
    out = np.zeros(n_frames, dtype=np.float32)
    pos = positions if positions.dtype == np.int64 else positions.astype(np.int32)
    if pos[-1] != n_frames:
        pos = np.concatenate([pos, [n_frames]])
    for i in range(len(pos) - 1):
        out[pos[i]:pos[i + 1]] = 0 if i == len(scores) else scores[i]
    return out


def generate_summary(scores, cps, n_frames, nfps, positions, proportion=0.15):
    frame_scores = upsample(scores, n_frames, positions)
    seg_score = [float(frame_scores[int(a):int(b) + 1].mean()) for a, b in cps]
    limits = int(math.floor(n_frames * proportion))
    values = (np.array(seg_score) * 1000).astype(np.int64)
    picks = set(c_oracle.knapsack(values, np.asarray(nfps, dtype=np.int64), limits))
    summary = np.zeros(1, dtype=np.float32)
    for s, nf in enumerate(nfps):                       # the reference grows the vector segment by segment
        summary = np.concatenate((summary, (np.ones if s in picks else np.zeros)(nf, dtype=np.float32)))
    return np.delete(summary, 0)


def evaluate_summary(machine_summary, user_summary):
    machine_summary = machine_summary.astype(np.float32)
    user_summary = user_summary.astype(np.float32)
    n_users, n_frames = user_summary.shape
    machine_summary[machine_summary > 0] = 1
    user_summary[user_summary > 0] = 1
    if len(machine_summary) > n_frames:
        machine_summary = machine_summary[:n_frames]
    elif len(machine_summary) < n_frames:
        machine_summary = np.concatenate([machine_summary, np.zeros(n_frames - len(machine_summary))])
    fs = []
    for u in range(n_users):
        g = user_summary[u]
        ov = (machine_summary * g).sum()
        p = ov / (machine_summary.sum() + 1e-8)
        r = ov / (g.sum() + 1e-8)
        fs.append(0. if (p == 0 and r == 0) else (2 * p * r) / (p + r))
    return np.mean(fs), np.max(fs)


def eval_video(args):
    """One video of the sweep: (scores, cps, n_frames, nfps, picks, user_summary) -> (avg_f, max_f)."""
    scores, cps, n_frames, nfps, picks, user_summary = args
    s = generate_summary(scores, cps, n_frames, nfps, picks)
    return evaluate_summary(s, user_summary)
