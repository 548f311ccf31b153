"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import CASES  # noqa: E402,F401


def load_golden(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def key_set(keys):
    return set(map(tuple, np.asarray(keys).reshape(-1, 3).tolist()))


def sort_keys(keys):
    keys = np.asarray(keys, np.int32).reshape(-1, 3)
    if len(keys) == 0:
        return keys
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order]


def oracle_params(ob, scene, case, **over):
    kw = dict(vox_size=case["vox_size"], trunc_margin=case["trunc"], voxels_per_block=case["vpb"], max_depth=case["max_depth"],
              use_color=1 if case["scene"].get("color") else 0)
    kw.update(over)
    return ob.params_for_scene(scene, **kw)


def engine_params(vh, scene, case, **over):
    kw = dict(vox_size=case["vox_size"], trunc_margin=case["trunc"], max_depth=case["max_depth"],
              use_color=1 if case["scene"].get("color") else 0, num_buckets=1 << 16, entries_per_bucket=4, pool_blocks=1 << 16,
              tri_arena_bytes=64 << 20)
    kw.update(over)
    return vh.params_for_scene(scene, **kw)
