"""Merging loop calls made at several resolutions -- the host-side step behind ``combine-resolutions``
(/root/reference/hicpeaks/utilities.py:469-552, script /root/reference/scripts/combine-resolutions).

Tiny data (peak lists), no GPU work: like the reference this is plain host code.  Behaviour kept, including the parts
that look accidental:

* resolutions are visited fine to coarse; a call at the finer resolution ``a`` is looked up among the calls of every
  coarser resolution ``b`` on the same chromosome by the Euclidean distance between the two anchor starts
  ``(x1, y1)``; the radius is ``2 * max_res`` when both resolutions are below ``2 * max_res`` and ``5 * max_res``
  otherwise (utilities.py:526-529);
* a call with a partner is kept whatever its own resolution, and its partners are *absorbed*: they are never emitted
  themselves and are skipped when their resolution takes its turn as the finer one (utilities.py:520-521, 531-534);
* a call without a partner (or on a chromosome the coarser set does not have) is kept only if its resolution is at most
  ``max_res`` and it is either at least ``good_res`` or shorter than ``mindis`` (utilities.py:522-525, 535-537); a later
  coarser set can still rescue it through a partner;
* calls of the coarsest resolution go through the same lone-call rule unless absorbed (utilities.py:539-544);
* with a single resolution everything is returned untouched, in input order, and NOT sorted (utilities.py:500-507);
  otherwise the result is the sorted set of ``(chrom, x1, x2, chrom, y1, y2)``.

The reference measures the distance with ``scipy.spatial.distance_matrix`` (float64 square root) and compares it with an
integer radius; coordinates are integers, so ``dx*dx + dy*dy <= radius*radius`` in exact integer arithmetic decides the
same way (the square root of an integer just above ``radius**2`` is above ``radius`` by about ``1 / (2 radius)``, far
beyond float64 resolution at these magnitudes).
"""
from __future__ import annotations

import numpy as np

__all__ = ["combine_annotations"]


def _record(chrom, p):
    return (chrom,) + tuple(p[:2]) + (chrom,) + tuple(p[2:])


def combine_annotations(byres, good_res=10000, mindis=100000, max_res=10000):
    """``byres``: {resolution in bp: {chrom: [(x1, x2, y1, y2), ...]}}.  Returns the merged peak list."""
    if len(byres) == 1:
        (calls,) = byres.values()
        return [_record(c, p) for c in calls for p in calls[c]]

    order = sorted(byres)
    near, far = 2 * max_res, 5 * max_res

    def lone_call_ok(res, p):
        return res <= max_res and (res >= good_res or p[2] - p[0] <= mindis)

    # anchor starts of every (resolution, chromosome) as integer arrays, built once
    starts = {res: {c: np.array([(p[0], p[2]) for p in calls], dtype=np.int64).reshape(-1, 2)
                    for c, calls in byres[res].items()} for res in order}
    kept, absorbed = set(), set()
    for ia, fine in enumerate(order[:-1]):
        for coarse in order[ia + 1:]:
            radius = near if (fine < near and coarse < near) else far
            for chrom, calls in byres[fine].items():
                partners = starts[coarse].get(chrom)
                for p in calls:
                    rec = _record(chrom, p)
                    if rec in absorbed:
                        continue
                    hits = ()
                    if partners is not None and len(partners):
                        dx = partners[:, 0] - p[0]
                        dy = partners[:, 1] - p[2]
                        hits = np.nonzero(dx * dx + dy * dy <= radius * radius)[0]
                    if len(hits):
                        kept.add(rec)
                        theirs = byres[coarse][chrom]
                        absorbed.update(_record(chrom, theirs[k]) for k in hits)
                    elif lone_call_ok(fine, p):
                        kept.add(rec)
    last = order[-1]
    for chrom, calls in byres[last].items():
        for p in calls:
            rec = _record(chrom, p)
            if rec not in absorbed and lone_call_ok(last, p):
                kept.add(rec)
    return sorted(kept)
