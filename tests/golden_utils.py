"""Helpers shared by the parity tests: load the committed golden traces (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not p.endswith(("unit_vectors.npz", "episode_storage_obstacles.npz", "reset_draws.npz")))


def load_case(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: d[k] for k in d.files if k != "meta"}
    g["meta"] = json.loads(str(d["meta"]))
    return g


def unit_vectors():
    d = np.load(os.path.join(GOLDEN_DIR, "unit_vectors.npz"))
    return {k: d[k] for k in d.files}


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor) with NaN==NaN and inf==inf treated as equal.  The floor only guards the division
    for values that pass through zero: north_star's 1e-9 is a RELATIVE bound, and velocities / rates are O(0.1)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    same = (np.isnan(a) & np.isnan(b)) | (a == b)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    e = np.where(same, 0.0, e)
    e = np.where(np.isnan(e), np.inf, e)
    return float(e.max()) if e.size else 0.0
