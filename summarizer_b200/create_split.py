"""Random train/test split generator (create_split.py:30-76): ``[{"train_keys": [...], "test_keys": [...]}, ...]``
with keys ``video_<n>``.  Also generates the split fixtures shipped under summarizer_b200/splits/ for the
synthetic SumMe-/TVSum-shaped datasets (same fold counts and sizes as the reference's files: 5 folds of 20/5 and
40/10 videos, plus one-fold "overfit" files whose 10 train keys are also the test keys)."""
import argparse
import json
import os

import numpy as np


def split_random(keys, num_splits, train_percent, seed=0):
    """``num_splits`` independent random splits (test sets may overlap, as in the reference's files)."""
    rng = np.random.default_rng(seed)
    keys = list(keys)
    n_train = int(round(len(keys) * train_percent))
    out = []
    for _ in range(num_splits):
        perm = rng.permutation(len(keys))
        out.append({"train_keys": [keys[i] for i in sorted(perm[:n_train])],
                    "test_keys": [keys[i] for i in sorted(perm[n_train:])]})
    return out


def write_fixtures(dst):
    os.makedirs(dst, exist_ok=True)
    for name, n in (("summe", 25), ("tvsum", 50)):
        keys = [f"video_{i}" for i in range(1, n + 1)]
        with open(os.path.join(dst, f"{name}_splits.json"), "w") as fh:
            json.dump(split_random(keys, 5, 0.8, seed=n), fh, indent=1)
        with open(os.path.join(dst, f"{name}_splits_overfit.json"), "w") as fh:
            json.dump([{"train_keys": keys[:10], "test_keys": keys[:10]}], fh, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser("Summarizer : Create splits")
    ap.add_argument("-n", "--n-videos", type=int, help="number of videos (keys video_1..video_n)")
    ap.add_argument("--save-path", type=str, default="splits/custom_splits.json")
    ap.add_argument("--num-splits", type=int, default=5)
    ap.add_argument("--train-percent", type=float, default=0.8)
    ap.add_argument("--fixtures", action="store_true", help="(re)generate the shipped synthetic split files")
    a = ap.parse_args()
    if a.fixtures:
        write_fixtures(os.path.join(os.path.dirname(os.path.abspath(__file__)), "splits"))
    else:
        with open(a.save_path, "w") as fh:
            json.dump(split_random([f"video_{i}" for i in range(1, a.n_videos + 1)], a.num_splits, a.train_percent), fh, indent=1)
