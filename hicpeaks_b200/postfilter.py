"""Host-side tail of the peak caller: gap filter, donut / lower-left combination, fold thresholds,
merge across (pw, ww) pairs and the anchor-based local clustering.

This runs on the few thousand FDR survivors the GPU returns, so it stays on the host and uses the
same third-party routines as the reference (``sklearn.cluster.dbscan``, ``scipy.signal.find_peaks`` /
``peak_widths``).  Behaviour follows /root/reference/hicpeaks/callers.py:

* ``gap_filter``        <- :291-312 (HiCCUPS) / :557-577 (BH-FDR)
* ``combine_pair``      <- :321-349
* ``find_anchors``      <- :593-634
* ``local_clustering``  <- :636-728 (``_cluster_core`` + ``local_clustering``)

including its quirks (asymmetric gap window ``[x - m, x + m)``, the seed pixel counted twice in the
first centroid, replacement across pairs only when *both* q-values improve).
"""
from __future__ import annotations

from collections import Counter

import numpy as np


def gap_filter(x, y, gaps, m, chrom_len):
    """Boolean keep-mask: no gap bin inside ``[lo(x), hi(x))`` or ``[lo(y), hi(y))`` with
    ``lo(v) = v - m if v > m else 0`` and ``hi(v) = v + m if v + m < chrom_len else chrom_len - 1``."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    if not gaps.any():
        return np.ones(x.size, dtype=bool)
    cum = np.concatenate([[0], np.cumsum(gaps.astype(np.int64))])   # cum[k] = gaps in [0, k)

    def hits(v):
        lo = np.where(v > m, v - m, 0)
        hi = np.where(v + m < chrom_len, v + m, chrom_len - 1)
        hi = np.maximum(hi, lo)                                     # empty range when hi <= lo
        return cum[hi] - cum[lo]

    return (hits(x) == 0) & (hits(y) == 0)


def combine_pair(table, res, K, Y, cemy_nonzero, double_fold, single_fold, use_raw):
    """Merge the donut (K) and lower-left (Y) survivors of one (pw, ww) pair into ``table``.

    K / Y: dicts of equally long arrays x, y, ice, obs, fold, p, q (gap-filtered survivors).
    cemy_nonzero: set of (x, y) whose lower-left expected-ratio entry is non-zero (callers.py:330)."""
    donut = {}
    for i in range(K["x"].size):
        lead = K["obs"][i] if use_raw else K["ice"][i]
        donut[(K["x"][i], K["y"][i])] = (lead, K["obs"][i], K["fold"][i], K["p"][i], K["q"][i])
    lower = {}
    for i in range(Y["x"].size):
        lower[(Y["x"][i], Y["y"][i])] = (Y["ice"][i], Y["obs"][i], Y["fold"][i], Y["p"][i], Y["q"][i])
    for pos, dn in donut.items():
        if pos in lower:
            ll = lower[pos]
        elif pos not in cemy_nonzero:
            ll = dn                        # no lower-left model at this pixel: reuse the donut record
        else:
            continue
        if dn[2] > double_fold and ll[2] > double_fold and (dn[2] > single_fold or ll[2] > single_fold):
            key = (pos[0] * res, pos[1] * res)
            rec = key + (0,) + dn + ll[2:]
            old = table.get(key)
            if old is None or (dn[-1] < old[7] and ll[-1] < old[10]):
                table[key] = rec
    return table


def find_anchors(pos, min_count=3, min_dis=20000, wlen=200000, res=10000):
    """1-D anchors (summit, left, right) on the marginal peak counts."""
    from scipy.signal import find_peaks, peak_widths

    min_dis = max(min_dis // res, 1)
    wlen = min(wlen // res, 10)
    count = Counter(pos)
    first = min(count) - 1
    ref = range(first, max(count) + 2)
    signal = np.r_[[count[i] for i in ref]]
    summits = find_peaks(signal, height=min_count, distance=min_dis)[0]
    ranked = sorted(((signal[i], i) for i in summits), reverse=True)

    anchors = set()
    owner = {}
    for _, i in ranked:
        bounds = peak_widths(signal, [i], rel_height=1, wlen=wlen)[2:4]
        lb = ref[int(np.round(bounds[0][0]))]
        rb = ref[int(np.round(bounds[1][0]))]
        summit = ref[i]
        if anchors:
            hit = next((owner[b] for b in range(lb, rb + 1) if b in owner), None)
            if hit is not None:            # overlaps an earlier (higher) anchor: absorb into it
                anchors.remove(hit)
                summit, lb, rb = hit[0], min(lb, hit[1]), max(rb, hit[2])
        rec = (summit, lb, rb)
        anchors.add(rec)
        for b in range(lb, rb + 1):
            owner[b] = rec
    return anchors


def _grow_clusters(ranked, r, visited, out):
    """DBSCAN pre-grouping, then greedy centroid growth from the strongest unassigned pixel."""
    from scipy.spatial.distance import euclidean
    from sklearn.cluster import dbscan

    pts = np.r_[[rec[1] for rec in ranked]]
    if len(pts) < 2:
        return
    _, labels = dbscan(pts, eps=r, min_samples=2)
    taken = set()
    for idx, (_, seed) in enumerate(ranked):
        if seed in taken or labels[idx] == -1:
            continue
        rest = pts[labels == labels[idx]]
        cen, rad = seed, r
        members = [seed]                   # the seed is visited again below: it weighs twice
        n_out_prev = -1
        while len(rest):
            outside = []
            for q in rest:
                tq = tuple(q)
                if tq in taken:
                    continue
                if euclidean(q, cen) <= rad:
                    members.append(tq)
                else:
                    outside.append(tq)
            if len(outside) == n_out_prev:
                break
            n_out_prev = len(outside)
            cen = tuple(np.r_[members].mean(axis=0).round().astype(int))
            rad = np.int32(np.round(max(euclidean(cen, q) for q in members))) + r
            rest = np.r_[outside]
        taken.update(members)
        out.append((seed, cen, rad))
    visited.update(taken)


def local_clustering(Donuts, LL, res, onlysummit=False, min_count=3, r=20000, sumq=1):
    """Returns [(pixel, centroid, radius)] -- clusters inside anchor rectangles, clusters among the
    rest, then isolated pixels passing the q-value rule."""
    out = []
    xs = np.r_[[k[0] for k in Donuts]]
    ys = np.r_[[k[1] for k in Donuts]]
    if xs.size == 0:
        return out
    x_anchors = find_anchors(xs, min_count=min_count, min_dis=r, res=res)
    y_anchors = find_anchors(ys, min_count=min_count, min_dis=r, res=res)
    r = max(r // res, 1)
    visited = set()
    present = set(zip(xs, ys))
    for xa in x_anchors:
        for ya in y_anchors:
            ranked = [(Donuts[(i, j)][0], (i, j)) for i in range(xa[1], xa[2] + 1)
                      for j in range(ya[1], ya[2] + 1) if (i, j) in present]
            ranked.sort(reverse=True)
            _grow_clusters(ranked, r, visited, out)
    ranked = [(Donuts[(i, j)][0], (i, j)) for i, j in zip(xs, ys) if (i, j) not in visited]
    ranked.sort(reverse=True)
    _grow_clusters(ranked, r, visited, out)

    x_summits = {a[0] for a in x_anchors}
    y_summits = {a[0] for a in y_anchors}
    for i, j in zip(xs, ys):
        if (i, j) in visited:
            continue
        if LL is not None:
            ok = Donuts[(i, j)][-1] + LL[(i, j)][-1] <= sumq
        else:
            ok = Donuts[(i, j)][-1] <= sumq / 2
        if ok and (not onlysummit or i in x_summits or j in y_summits):
            out.append(((i, j), (i, j), 0))
    return out
