"""TEST INFRASTRUCTURE ONLY -- loop-level restatement of the reference's post-FDR tail
(/root/reference/hicpeaks/callers.py:289-362 and :593-728) used to finish the oracle's peak table.

Deliberately written with the reference's own data structures (dicts / sets filled in the same
sequence) so that set-iteration-order effects, if any, are reproduced.  Product code must not
import this module.
"""
from __future__ import annotations

from collections import Counter

import numpy as np


def gaps_of(bal):
    return set(np.where(bal.sum(axis=0) == 0)[0])


def gap_keep(xi, yi, gaps, m, chrom_len):
    if len(gaps) == 0:
        return list(range(len(xi)))
    keep = []
    for i in range(len(xi)):
        region = set()
        for v in (xi[i], yi[i]):
            lo = (v - m) if v > m else 0
            hi = (v + m) if (v + m) < chrom_len else (chrom_len - 1)
            region |= set(range(lo, hi))
        if not (region & gaps):
            keep.append(i)
    return keep


def anchors_1d(pos, min_count, min_dis, res, wlen=200000):
    from scipy.signal import find_peaks, peak_widths
    min_dis = max(min_dis // res, 1)
    wlen = min(wlen // res, 10)
    count = Counter(pos)
    refidx = range(min(count) - 1, max(count) + 2)
    signal = np.r_[[count[i] for i in refidx]]
    summits = find_peaks(signal, height=min_count, distance=min_dis)[0]
    order = [(signal[i], i) for i in summits]
    order.sort(reverse=True)
    peaks, records = set(), {}
    for _, i in order:
        w = peak_widths(signal, [i], rel_height=1, wlen=wlen)[2:4]
        lb, rb = refidx[int(np.round(w[0][0]))], refidx[int(np.round(w[1][0]))]
        new = (refidx[i], lb, rb)
        if len(peaks):
            for b in range(lb, rb + 1):
                if b in records:
                    old = records[b]
                    new = (old[0], min(lb, old[1]), max(rb, old[2]))
                    peaks.remove(old)
                    break
        peaks.add(new)
        for b in range(new[1], new[2] + 1):
            records[b] = new
    return peaks


def cluster_core(sort_list, r, visited, final_list):
    from scipy.spatial.distance import euclidean
    from sklearn.cluster import dbscan
    pos = np.r_[[i[1] for i in sort_list]]
    if len(pos) < 2:
        return
    _, labels = dbscan(pos, eps=r, min_samples=2)
    pool = set()
    for i, p in enumerate(sort_list):
        if p[1] in pool or labels[i] == -1:
            continue
        sub = pos[labels == labels[i]]
        cen, rad, local, ini = p[1], r, [p[1]], -1
        while len(sub):
            out = []
            for q in sub:
                if tuple(q) in pool:
                    continue
                (local if euclidean(q, cen) <= rad else out).append(tuple(q))
            if len(out) == ini:
                break
            ini = len(out)
            cen = tuple(np.r_[local].mean(axis=0).round().astype(int))
            rad = np.int32(np.round(max([euclidean(cen, q) for q in local]))) + r
            sub = np.r_[out]
        for q in local:
            pool.add(q)
        final_list.append((p[1], cen, rad))
    visited.update(pool)


def clustering(Donuts, LL, res, onlysummit, min_count, r, sumq):
    final_list = []
    x = np.r_[[i[0] for i in Donuts]]
    y = np.r_[[i[1] for i in Donuts]]
    if x.size == 0:
        return final_list
    xa = anchors_1d(x, min_count, r, res)
    ya = anchors_1d(y, min_count, r, res)
    r = max(r // res, 1)
    visited, lookup = set(), set(zip(x, y))
    for a in xa:
        for b in ya:
            sl = [(Donuts[(i, j)][0], (i, j)) for i in range(a[1], a[2] + 1) for j in range(b[1], b[2] + 1)
                  if (i, j) in lookup]
            sl.sort(reverse=True)
            cluster_core(sl, r, visited, final_list)
    sl = [(Donuts[(i, j)][0], (i, j)) for i, j in zip(x, y) if (i, j) not in visited]
    sl.sort(reverse=True)
    cluster_core(sl, r, visited, final_list)
    xs, ys = set(i[0] for i in xa), set(i[0] for i in ya)
    for i, j in zip(x, y):
        if (i, j) in visited:
            continue
        qpass = (Donuts[(i, j)][-1] + LL[(i, j)][-1] <= sumq) if LL is not None else (Donuts[(i, j)][-1] <= sumq / 2)
        if qpass and ((not onlysummit) or (i in xs) or (j in ys)):
            final_list.append(((i, j), (i, j), 0))
    return final_list


def finish_hiccups(inp, sw, res_by_pf, pw, ww, res, sumq, double_fold, single_fold, use_raw,
                   min_marginal_peaks, onlyanchor):
    """callers.py:289-362 given the oracle's per-(p, background) results."""
    n = inp["n"]
    gaps = gaps_of(sw["bal"])
    table = {}
    for p, w in zip(pw, ww):
        pre = []
        for fl in (0, 1):
            r = res_by_pf[(p, fl)]
            rej = np.where(r["reject"])[0]
            keep = gap_keep(r["x"][rej], r["y"][rej], gaps, min(ww), n)
            idx = rej[keep]
            lead = r["O"] if (use_raw and fl == 0) else r["ice"]
            pre.append(dict(zip(zip(r["x"][idx], r["y"][idx]),
                                zip(lead[idx], r["O"][idx], r["fold"][idx], r["p"][idx], r["q"][idx]))))
        donuts, lls = pre
        cem_nz = set(zip(*res_by_pf[(p, 1)]["cem_nz"]))
        common = set(donuts) & set(lls)
        for pos in set(donuts) - set(lls):
            if pos not in cem_nz:
                common.add(pos)
        for pos in common:
            dn = donuts[pos]
            ll = lls[pos] if pos in lls else dn
            if dn[2] > double_fold and ll[2] > double_fold and (dn[2] > single_fold or ll[2] > single_fold):
                key = (pos[0] * res, pos[1] * res)
                if key not in table or (dn[-1] < table[key][7] and ll[-1] < table[key][10]):
                    table[key] = key + (0,) + dn + ll[2:]
    Donuts = {(k[0] // res, k[1] // res): table[k][3:8] for k in table}
    LL = {(k[0] // res, k[1] // res): table[k][8:] for k in table}
    final = {}
    for pixel, cen, rad in clustering(Donuts, LL, res, onlyanchor, min_marginal_peaks, 2 * res, sumq):
        key = (pixel[0] * res, pixel[1] * res)
        final[key] = (cen[0] * res, cen[1] * res) + (rad * res,) + table[key][4:]
    return final


def finish_bhfdr(inp, sw, r, ww, res, min_marginal_peaks, onlyanchor):
    """callers.py:555-590: gap filter with m = ww, clustering without the lower-left table, fold > 2."""
    rej = r["reject"]
    x, y = r["x"][rej], r["y"][rej]
    vals = [r[k][rej] for k in ("O", "fold", "p", "q")]
    keep = gap_keep(x, y, gaps_of(sw["bal"]), ww, inp["n"])
    x, y = x[keep], y[keep]
    vals = [v[keep] for v in vals]
    Donuts = dict(zip(zip(x, y), zip(*vals)))
    table = {}
    for pixel, cen, radius in clustering(Donuts, None, res, onlyanchor, min_marginal_peaks, 2 * res, 1):
        donut = Donuts[pixel]
        if donut[1] > 2:
            table[(pixel[0] * res, pixel[1] * res)] = (cen[0] * res, cen[1] * res) + (radius * res,) + donut
    return table
