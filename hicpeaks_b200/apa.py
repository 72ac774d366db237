"""Drop-in replacements for ``hicpeaks.apa.apa_submatrix`` / ``apa_analysis``
(/root/reference/hicpeaks/apa.py:11-46) on the CUDA engine.

``apa_submatrix`` gathers, filters and normalises the windows on the GPU and returns an ``ApaWindows``
handle (the windows stay in HBM; it behaves like the reference's list of arrays when iterated or
converted with ``numpy.asarray`` / ``numpy.r_``).  ``apa_analysis`` accepts such handles (one, or a
list of them -- one per chromosome, as ``scripts/apa-analysis:82-126`` accumulates them) or a plain
``(N, 2w+1, 2w+1)`` array, and returns ``(avg, score, z, p, maxi)``.

The per-window means (bit-identical to numpy's pairwise summation) and the sequential axis-0 sum run
in CUDA kernels through the C ABI; the percentile cut and the five summary numbers are computed
with the same numpy / scipy calls as the reference on the 41 x 41 result.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr

from . import _capi

__all__ = ["apa_submatrix", "apa_analysis", "ApaWindows"]


class ApaWindows:
    """Normalised windows of one chromosome, resident on the GPU."""

    def __init__(self, ctx, valid, mean_arr, w):
        self.ctx, self.w = ctx, int(w)
        self.index = np.nonzero(valid)[0].astype(np.int64)     # window slots that passed apa.py:18-24
        self.mean_arr = mean_arr[self.index]

    def __len__(self):
        return int(self.index.size)

    def close(self):
        """Free the GPU memory of this chromosome (band copy, windows); the handle is unusable afterwards."""
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __getitem__(self, k):
        if isinstance(k, slice):
            return list(self.ctx.apa_get_windows(self.index[k], self.w))
        return self.ctx.apa_get_windows(self.index[[k]], self.w)[0]

    def __iter__(self):
        return iter(self.ctx.apa_get_windows(self.index, self.w))

    def __array__(self, dtype=None, copy=None):
        a = self.ctx.apa_get_windows(self.index, self.w)
        return a if dtype is None else a.astype(dtype)


def _diagonals(M, num):
    n = M.shape[0]
    return [np.ascontiguousarray(np.asarray(M.diagonal(k), dtype=np.float64).ravel()) for k in range(min(num, n))]


def apa_submatrix(M, pos, w=5, device=0):
    """``M``: the (symmetric) balanced matrix of one chromosome, scipy sparse or dense; ``pos``: iterable of
    ``(i, j)`` bin pairs.  Same filtering as apa.py:18-24; returns an ``ApaWindows``."""
    pos = np.asarray(list(pos), dtype=np.int64).reshape(-1, 2)
    n = M.shape[0]
    inside = (pos[:, 0] - w >= 0) & (pos[:, 0] + w + 1 <= n) & (pos[:, 1] - w >= 0) & (pos[:, 1] + w + 1 <= n)
    span = int(np.abs(pos[inside, 1] - pos[inside, 0]).max()) if inside.any() else 0
    ctx = _capi.Context(device)
    ctx.apa_upload(n, _diagonals(M, span + 2 * w + 1))
    valid, mean = ctx.apa_windows(pos[:, 0], pos[:, 1], w)
    return ApaWindows(ctx, valid, mean, w)


def apa_analysis(apa, w=5, cw=3, device=0):
    """apa.py:30-46.  ``apa``: ``ApaWindows``, a list of ``ApaWindows`` (windows of several chromosomes, in
    order), or an array / list of ``(2w+1, 2w+1)`` arrays."""
    if isinstance(apa, ApaWindows):
        parts = [apa]
    elif isinstance(apa, (list, tuple)) and len(apa) and all(isinstance(a, ApaWindows) for a in apa):
        parts = list(apa)
    else:
        arr = np.ascontiguousarray(np.asarray(apa), dtype=np.float64)
        if arr.ndim != 3 or arr.shape[1] != 2 * w + 1 or arr.shape[2] != 2 * w + 1:
            raise ValueError("apa must be (N, 2w+1, 2w+1)")
        ctx = _capi.Context(device)
        mean = ctx.apa_load_windows(arr, w)
        parts = [ApaWindows(ctx, np.ones(arr.shape[0], dtype=bool), mean, w)]
    # remove outliers (apa.py:33-36)
    mean_arr = np.concatenate([p.mean_arr for p in parts])
    p99 = np.percentile(mean_arr, 99)
    p1 = np.percentile(mean_arr, 1)
    mask = (mean_arr < p99) & (mean_arr > p1)
    side = 2 * w + 1
    acc = np.zeros(side * side, dtype=np.float64)
    count, off, init = int(mask.sum()), 0, True
    for p in parts:
        sel = p.index[mask[off:off + len(p)]]
        off += len(p)
        if sel.size:
            p.ctx.apa_accumulate(sel, acc, init)
            init = False
    with np.errstate(invalid="ignore", divide="ignore"):
        avg = (acc / count if count else acc * np.nan).reshape(side, side)       # apa[mask].mean(axis=0)
        lowerpart = avg[-cw:, :cw]
        upperpart = avg[:cw, -cw:]
        maxi = upperpart.mean() * 5
        score = avg[w, w] / lowerpart.mean()
        z = (avg[w, w] - lowerpart.mean()) / lowerpart.std()
        p = 1 - ndtr(z)
    return avg, score, z, p, maxi
