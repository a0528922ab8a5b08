"""Loaders for tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

VERSIONS = ["barrage", "standard", "short_standard", "short_barrage", "medium_standard", "octa_barrage",
            "standard2", "medium", "fives", "tiny", "micro"]

# more recorded games of the two benchmark variants (other seeds, files of their own: oracle/gen_golden.py DEEP_PLAN)
DEEP = {"standard_b": "standard", "barrage_b": "barrage"}
TRAJ_LABELS = VERSIONS + sorted(DEEP)


def variant_of(label):
    """game variant a trajectory label was recorded on"""
    return DEEP.get(label, label)


_cache = {}


def load(name):
    if name not in _cache:
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as d:
            _cache[name] = {k: d[k] for k in d.files}
    return _cache[name]


def traj(version):
    return load("traj_" + version)


def known():
    return load("known_answers")


def original_channels():
    return load("original_channels")


def side_channels():
    return load("side_channels")


def custom_toys():
    return load("custom_toys")


def spatial_alias():
    return load("spatial_alias")


ALIAS_VERSIONS = ["barrage", "micro", "tiny", "standard2", "octa_barrage"]


def unpack_mask(bits, n):
    return np.unpackbits(bits, axis=-1)[..., :n]


def transitions(t):
    """indices i with a recorded move states[i] -(actions[i])-> states[i+1]"""
    return np.flatnonzero(t["actions_spatial"] >= 0)
