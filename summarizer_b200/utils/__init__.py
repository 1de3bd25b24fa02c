"""Helpers with the reference's names (utils/__init__.py:4-31)."""
import json
import os


def parse_splits_filename(splits_filename):
    """utils/__init__.py:4-17 — (dataset_name, splits) of a split JSON file; the dataset name is
    the part of the file name before the first underscore."""
    stem = os.path.splitext(os.path.basename(splits_filename))[0]
    with open(splits_filename, "r") as fh:
        splits = json.load(fh)
    return stem.split("_")[0], splits


class Proportion(object):
    """utils/__init__.py:19-31 — argparse `choices` helper accepting any float in ]0, 1]."""
    _label = "a proportion value in ]0, 1]"

    def __eq__(self, value):
        return 0 < value <= 1

    def __contains__(self, item):
        return self == item

    def __iter__(self):
        yield self._label

    def __str__(self):
        return self._label
