"""Random train/test split generator with the reference's interface (create_split.py:30-76):
``split_random(keys, num_videos, num_train) -> (train_keys, test_keys)`` and the ``-d/--dataset --save-dir --save-name
--num-splits --train-percent`` command line, writing ``[{"train_keys": [...], "test_keys": [...]}, ...]``.
The dataset's keys are read with h5py when it is installed and with this package's own HDF5 reader otherwise
(``utils/hdf5.py``); ``--n-videos N`` is an extra for datasets that are not on disk (keys ``video_1..video_N``).
The generator of the split fixtures shipped under ``splits/`` lives in ``scripts/make_split_fixtures.py``."""
import argparse
import json
import math
import os

import numpy as np


def split_random(keys, num_videos, num_train):
    """Random split: ``num_train`` of the ``num_videos`` keys (drawn without replacement) train, the rest test;
    both lists keep the order of ``keys``."""
    chosen = set(np.random.choice(range(num_videos), size=num_train, replace=False).tolist())
    train_keys = [k for i, k in enumerate(keys) if i in chosen]
    test_keys = [k for i, k in enumerate(keys) if i not in chosen]
    assert not set(train_keys) & set(test_keys), "Error: train_keys and test_keys overlap"
    return train_keys, test_keys


def dataset_keys(path):
    try:
        import h5py
        with h5py.File(path, "r") as f:
            return list(f.keys())
    except ImportError:
        from .utils import hdf5
        with hdf5.File(path, "r") as f:
            return list(f.keys())


def make_splits(keys, num_splits, train_percent):
    keys = list(keys)
    num_train = int(math.ceil(len(keys) * train_percent))              # create_split.py:55
    return [dict(zip(("train_keys", "test_keys"), split_random(keys, len(keys), num_train))) for _ in range(num_splits)]


def main(argv=None):
    parser = argparse.ArgumentParser("Code to create splits in json form")
    parser.add_argument("-d", "--dataset", type=str, help="path to h5 dataset")
    parser.add_argument("-n", "--n-videos", type=int, help="instead of --dataset: keys video_1..video_N")
    parser.add_argument("--save-dir", type=str, default="splits", help="path to save output json file (default: 'splits')")
    parser.add_argument("--save-name", type=str, default="new_split", help="name to save as, excluding extension")
    parser.add_argument("--num-splits", type=int, default=5, help="how many splits to generate (default: 5)")
    parser.add_argument("--train-percent", type=float, default=0.8, help="percentage of training data (default: 0.8)")
    args = parser.parse_args(argv)
    if (args.dataset is None) == (args.n_videos is None):
        parser.error("exactly one of -d/--dataset and -n/--n-videos is required")
    keys = dataset_keys(args.dataset) if args.dataset else [f"video_{i}" for i in range(1, args.n_videos + 1)]
    splits = make_splits(keys, args.num_splits, args.train_percent)
    n_train = len(splits[0]["train_keys"]) if splits else 0
    print(f"Split breakdown: # total videos {len(keys)}. # train videos {n_train}. # test videos {len(keys) - n_train}")
    os.makedirs(args.save_dir, exist_ok=True)
    saveto = os.path.join(args.save_dir, f"{args.save_name}.json")
    with open(saveto, "w") as f:
        json.dump(splits, f, indent=4, separators=(",", ": "))
    print(f"Splits saved to {saveto}")
    return saveto


if __name__ == "__main__":
    main()
