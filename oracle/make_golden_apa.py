"""TEST INFRASTRUCTURE ONLY -- tests/golden/apa_*.npz from the UNMODIFIED reference ``hicpeaks/apa.py``
(imported by path; run in the build container only):

    python oracle/make_golden_apa.py

Inputs are stored as raw band counts + weights (the balanced matrix is ``(w[r] * w[c]) * count`` (cooler: ``bias1[row] * bias2[col] * data``), rebuilt
identically by ``oracle.apa_oracle.balanced_diags``); outputs are the reference's valid-window flags,
per-window means, averaged window and the four summary numbers.
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hicpeaks_b200.synth import synth_chromosome  # noqa: E402
from oracle import apa_oracle, ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def csr_from_diags(diags, n):
    """Full symmetric CSR with explicit NaN entries, like cooler's fetch()."""
    rows, cols, vals = [], [], []
    for d, v in enumerate(diags):
        nz = np.nonzero((v != 0) | np.isnan(v))[0]
        rows.append(nz); cols.append(nz + d); vals.append(v[nz])
        if d:
            rows.append(nz + d); cols.append(nz); vals.append(v[nz])
    return sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def make(name, n, band, w, cw, npos, seed, scale):
    ref = ref_harness.load_reference_apa()
    inp = synth_chromosome(n, band, 1, maxww=0, seed=seed, scale=scale)
    diags = apa_oracle.balanced_diags(inp["Diags"], inp["weights"])
    M = csr_from_diags(diags, n)
    rng = np.random.default_rng(seed + 1)
    i = rng.integers(0, n, npos)
    dist = rng.integers(0, band - 2 * w, npos)          # includes anchors whose window crosses the main diagonal
    j = np.minimum(i + dist, n - 1)
    i[:3] = [0, w - 1, n - 1]                           # windows that leave the matrix
    pos = list(zip(i.tolist(), j.tolist()))
    apa = ref.apa_submatrix(M, pos, w=w)
    avg, score, z, p, maxi = ref.apa_analysis(np.r_[apa], w=w, cw=cw)
    mean_arr = np.array([np.mean(a) for a in apa])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), kind="apa", n=n, num=inp["num"], w=w, cw=cw,
                        raw=np.concatenate(inp["Diags"]).astype(np.int32), weights=inp["weights"], pos=np.array(pos),
                        n_windows=len(apa), mean_arr=mean_arr, avg=avg, stats=np.array([score, z, p, maxi]),
                        win_first=apa[0], win_last=apa[-1])
    print(name, "anchors", npos, "windows", len(apa), "score %.4f z %.3f p %.3g" % (score, z, p),
          "distinct means", np.unique(mean_arr).size)


if __name__ == "__main__":
    make("apa_w5", 900, 120, 5, 3, 700, 5, 60.0)
    make("apa_w20", 1500, 200, 20, 3, 1500, 6, 120.0)
