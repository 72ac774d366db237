"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference APA path on dense diagonal arrays.

* ``balanced_diags``  <- cooler's ``matrix(balance=name, sparse=True)`` as consumed at
                         /root/reference/scripts/apa-analysis:94: ``(w[r] * w[c]) * count`` (cooler: ``bias1[row] * bias2[col] * data``), NaN where a stored
                         count meets a NaN weight, 0 where nothing is stored
* ``apa_submatrix``   <- /root/reference/hicpeaks/apa.py:11-28
* ``apa_analysis``    <- apa.py:30-46

The arithmetic that matters (``ndarray.mean`` = pairwise summation over the flattened window, the sequential
axis-0 mean, ``numpy.percentile``) is numpy's own, exactly what the reference calls; ``pairwise_sum`` restates
numpy's algorithm explicitly because the CUDA kernel has to reproduce it bit for bit.

Parity status: PINNED -- ``oracle/make_golden_apa.py`` runs the unmodified ``hicpeaks/apa.py`` (imported by
path) on a scipy CSR matrix and ``tests/test_oracle_golden.py`` checks this module against the stored outputs.
Nothing in the product imports this module.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr


def balanced_diags(Diags, weights, num=None):
    n = len(Diags[0])
    num = len(Diags) if num is None else num
    w = np.asarray(weights, dtype=np.float64)
    out = []
    for d in range(num):
        raw = np.asarray(Diags[d])
        with np.errstate(invalid="ignore"):
            bal = w[: n - d] * w[d:] * raw.astype(np.float64)      # cooler: bias1[row] * bias2[col] * data
        bal[raw == 0] = 0.0
        out.append(bal)
    return out


def pairwise_sum(a):
    """numpy's float64 add.reduce over a contiguous 1-D array (umath ``pairwise_sum``)."""
    n = len(a)
    if n < 8:
        res = 0.0
        for x in a:
            res += float(x)
        return res
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def window(diags, n, i, j, w):
    """Dense (2w+1)^2 window of the symmetric matrix whose upper diagonals are ``diags``."""
    rr = np.arange(i - w, i + w + 1)[:, None]
    cc = np.arange(j - w, j + w + 1)[None, :]
    lo = np.minimum(rr, cc)
    d = np.abs(cc - rr)
    out = np.zeros((2 * w + 1, 2 * w + 1))
    for dd in np.unique(d):
        if dd < len(diags):
            m = d == dd
            out[m] = diags[dd][lo[m]]
    return out


def apa_submatrix(diags, n, pos, w=5):
    """apa.py:11-28.  Returns (list of normalised windows, valid flags over pos)."""
    apa, valid = [], []
    for i, j in pos:
        ok = False
        if (i - w >= 0) and (i + w + 1 <= n) and (j - w >= 0) and (j + w + 1 <= n):
            tmp = window(diags, n, i, j, w)
            if not np.isnan(tmp).sum() > 0 and not tmp.mean() == 0:
                apa.append(tmp / tmp.mean())
                ok = True
        valid.append(ok)
    return apa, np.array(valid, dtype=bool)


def apa_analysis(apa, w=5, cw=3):
    """apa.py:30-46."""
    apa = np.asarray(apa)
    mean_arr = np.r_[[np.mean(arr) for arr in apa]]
    p99 = np.percentile(mean_arr, 99)
    p1 = np.percentile(mean_arr, 1)
    mask = (mean_arr < p99) & (mean_arr > p1)
    avg = apa[mask].mean(axis=0)
    lowerpart = avg[-cw:, :cw]
    upperpart = avg[:cw, -cw:]
    maxi = upperpart.mean() * 5
    score = avg[w, w] / lowerpart.mean()
    z = (avg[w, w] - lowerpart.mean()) / lowerpart.std()
    p = 1 - ndtr(z)
    return avg, score, z, p, maxi, mean_arr, mask
