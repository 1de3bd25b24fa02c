"""Generates the split fixtures for the synthetic SumMe-/TVSum-shaped datasets (same fold counts and sizes as the
reference's files: 5 folds of 20/5 and 40/10 videos, plus one-fold "overfit" files whose 10 train keys are also the
test keys).  The files shipped under summarizer_b200/splits/ are the reference's own fixtures; this script only
exists to make fixtures for other synthetic sizes."""
import json
import os
import sys

import numpy as np


def split_seeded(keys, num_splits, train_percent, seed):
    rng = np.random.default_rng(seed)
    n_train = int(np.ceil(len(keys) * train_percent))
    out = []
    for _ in range(num_splits):
        perm = rng.permutation(len(keys))
        out.append({"train_keys": [keys[i] for i in sorted(perm[:n_train])],
                    "test_keys": [keys[i] for i in sorted(perm[n_train:])]})
    return out


if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else "splits_synthetic"
    os.makedirs(dst, exist_ok=True)
    for name, n in (("summe", 25), ("tvsum", 50)):
        keys = [f"video_{i}" for i in range(1, n + 1)]
        with open(os.path.join(dst, f"{name}_splits.json"), "w") as fh:
            json.dump(split_seeded(keys, 5, 0.8, seed=n), fh, indent=1)
        with open(os.path.join(dst, f"{name}_splits_overfit.json"), "w") as fh:
            json.dump([{"train_keys": keys[:10], "test_keys": keys[:10]}], fh, indent=1)
