"""Loads tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np

from hicpeaks_b200.synth import _finish

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind="hiccups"):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLD, "*.npz"))):
        with np.load(f) as z:
            if str(z["kind"]) == kind:
                out.append(os.path.basename(f)[:-4])
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    z = dict(np.load(os.path.join(GOLD, name + ".npz")))
    n, num, mw = int(z["in_n"]), int(z["in_num"]), int(z["in_min_ww"])
    band = z["in_band"]
    Diags = [np.ascontiguousarray(band[d, : n - d]) for d in range(num)]
    inp = _finish(n, num, mw, Diags, z["in_weights"])
    kw = {}
    for k, v in z.items():
        if k.startswith("kw_"):
            v = v.tolist()
            kw[k[3:]] = v
    return z, inp, kw, int(z["res"])


def table_rows(table):
    rows = [list(k) + [float(v) for v in table[k]] for k in sorted(table)]
    return np.array(rows, dtype=np.float64).reshape(len(rows), 12)


def load_apa(name):
    """(z, n, raw diagonals, weights) of an APA fixture (oracle/make_golden_apa.py)."""
    z = dict(np.load(os.path.join(GOLD, name + ".npz")))
    n, num = int(z["n"]), int(z["num"])
    off = np.concatenate([[0], np.cumsum([n - d for d in range(num)])])
    Diags = [z["raw"][off[d]:off[d + 1]] for d in range(num)]
    return z, n, Diags, z["weights"]
