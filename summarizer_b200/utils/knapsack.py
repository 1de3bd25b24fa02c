"""0/1 knapsack with the reference's signature (utils/knapsack.py:5-23), solved by the sm_100a
DP kernel (smz_knapsack) with OR-tools' KNAPSACK_DYNAMIC_PROGRAMMING_SOLVER semantics
(strict-improvement updates, last-improving-item extraction, sum(weights) <= capacity shortcut)."""
import numpy as np
import torch

from .. import _native as N
from ..batch import VideoBatch


def knapsack_ortools(values, weights, items, capacity):
    """0-1 Knapsack problem solver.  values are floats scaled by 1000 and truncated
    (utils/knapsack.py:11-14); returns the ascending list of packed item indices.
    ``items`` is unused, as in the reference."""
    scale = 1000
    vals = (np.array(values) * scale).astype(np.int64)        # float64 product, truncation toward 0
    w = np.array(weights).astype(np.int64)
    n = len(w)
    if n == 0:
        return []
    if np.any(w < 0):
        raise ValueError("negative weights")
    if np.abs(vals).max(initial=0) > (2**31 - 1) // max(n, 1) or w.sum() > 2**31 - 1:
        raise OverflowError("knapsack values/weights exceed the int32 range of the DP kernel")
    desc = np.zeros(1, dtype=N.VIDEO_DESC)
    desc["n_segs"], desc["capacity"], desc["summ_len"] = n, int(capacity), int(w.sum())
    empty = np.zeros(0, np.int32)
    b = VideoBatch.from_packed(desc, empty, empty, w.astype(np.int32), None, proportion=0.15)
    b.knapsack(torch.from_numpy(vals.astype(np.int32)))
    b.check_status()
    return np.nonzero(b.picked[:n].cpu().numpy())[0].tolist()
