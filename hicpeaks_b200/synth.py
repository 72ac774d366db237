"""Synthetic diagonal-band Hi-C inputs (workload definitions of BASELINE.json configs 2-5).

Produces exactly the objects the reference worker hands to ``hiccups()``
(/root/reference/scripts/pyHICCUPS:146-166) without going through cooler:

* ``Diags[d]``  -- raw counts on diagonal ``d`` (int32, length ``n - d``), ``d in [0, num)``
* ``cDiags[i]`` -- balanced values ``(w[r] * w[c]) * count`` (cooler: ``bias1[row] * bias2[col] * data``) on diagonal ``min(ww) + i`` with
  NaN -> 0 (cooler yields NaN only where a *stored* count meets a NaN weight)
* ``IR[d]``     -- mean of the non-NaN entries of balanced diagonal ``d`` (zeros count)
* ``biases``    -- ``1 / w`` (0 where ``w`` is 0 or NaN)

The generator follows SURVEY.md section 8(d): ``counts[d] ~ Poisson(300 (d+1)^-1.08)``, planted loops,
log-normal weights with 1 % NaN bins and one contiguous gap block.
"""
from __future__ import annotations

import numpy as np

__all__ = ["BandInput", "synth_chromosome", "band_from_dense", "band_pixels", "hg38_autosome_bins"]


class BandInput(dict):
    """Plain container: keys n, num, min_ww, Diags, cDiags, IR, biases, weights (attribute access too)."""

    __getattr__ = dict.__getitem__


def band_pixels(n: int, lo: int, hi: int) -> int:
    """Number of band pixels with ``lo <= c - r <= hi`` and ``c < n`` (SURVEY 8d 'unit of work')."""
    hi = min(hi, n - 1)
    if hi < lo:
        return 0
    k = hi - lo + 1
    return k * n - (lo + hi) * k // 2


def _finish(n, num, min_ww, Diags, weights):
    w = np.asarray(weights, dtype=np.float64)
    IR = {}
    cDiags = []
    for d in range(min_ww, num):
        raw = Diags[d]
        wr = w[: n - d]
        wc = w[d:]
        with np.errstate(invalid="ignore"):
            bal = wr * wc * raw.astype(np.float64)   # cooler's order: bias1[row] * bias2[col] * data, left to right
        bal[raw == 0] = 0.0                      # unstored pixels read back as 0, never NaN
        mask = np.isnan(bal)
        notnan = bal[~mask]
        with np.errstate(invalid="ignore", divide="ignore"):
            IR[d] = notnan.mean() if notnan.size else np.float64(np.nan)
        bal[mask] = 0.0
        cDiags.append(bal)
    good = ~((w == 0) | np.isnan(w))
    biases = np.zeros_like(w)
    biases[good] = 1.0 / w[good]
    return BandInput(n=int(n), num=int(num), min_ww=int(min_ww), Diags=Diags, cDiags=cDiags, IR=IR,
                     biases=biases, weights=w)


def synth_chromosome(n: int, band: int, min_ww: int, maxww: int = 10, seed: int = 0,
                     nan_frac: float = 0.01, gap_frac: float = 0.02, loops_per_bin: float = 1.0 / 200,
                     scale: float = 300.0, decay: float = 1.08) -> BandInput:
    """One synthetic chromosome of ``n`` bins with scored band ``band`` bins (= maxapart // res)."""
    rng = np.random.default_rng(seed)
    num = band + maxww + 1
    if num > n:
        raise ValueError("band wider than the chromosome")
    Diags = []
    lam = scale * np.power(np.arange(num) + 1.0, -decay)
    for d in range(num):
        Diags.append(rng.poisson(lam[d], n - d).astype(np.int32))
    nloops = int(n * loops_per_bin)
    lo_d = min(15, num - 1)
    for _ in range(nloops):
        d = int(rng.integers(lo_d, num))
        i = int(rng.integers(0, n - d))
        Diags[d][i] += int(5 * lam[d] + 20)
    w = np.exp(rng.normal(0.0, 0.2, n))
    nbad = int(round(n * nan_frac))
    if nbad:
        w[rng.choice(n, nbad, replace=False)] = np.nan
    glen = int(round(n * gap_frac))
    if glen:
        g0 = int(rng.integers(n // 4, n // 2))
        w[g0:g0 + glen] = np.nan
    w *= 0.05
    # masked bins carry no stored signal in a balanced cooler; keep raw counts (the reference
    # still reads them through M) -- only the weights are NaN, exactly like a cooler file.
    return _finish(n, num, min_ww, Diags, w)


def band_from_dense(counts: np.ndarray, weights: np.ndarray, num: int, min_ww: int) -> BandInput:
    """Band input from a dense symmetric count matrix (small tests, the chr21 example)."""
    n = counts.shape[0]
    Diags = [np.ascontiguousarray(np.diagonal(counts, d)).astype(np.int32) for d in range(num)]
    return _finish(n, num, min_ww, Diags, weights)


def hg38_autosome_bins(res: int):
    """Bin counts of the 22 hg38 autosomes (lengths as in /root/reference/example/hg38.chromsizes)."""
    sizes = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
             138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
             83257441, 80373285, 58617616, 64444167, 46709983, 50818468]
    return [-(-s // res) for s in sizes]
