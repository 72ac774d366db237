"""Drop-in replacements for ``hicpeaks.callers.hiccups`` / ``bhfdr`` on top of the CUDA engine.

Same names, argument order, defaults, return value and error behaviour as the reference operator
(/root/reference/hicpeaks/callers.py:44-46 and :364-365), so the per-chromosome worker of
``pyHICCUPS`` (/root/reference/scripts/pyHICCUPS:170-173) can call it unchanged.  The sweep, the
expected values, the lambda-chunk Poisson test and the BH-FDR run on the GPU through the C ABI
(``include/hicpeaks_b200.h``); only the few-thousand-survivor tail (``postfilter.py``) runs here.
There is no CPU implementation of the scoring path in this package.
"""
from __future__ import annotations

import logging
import threading

import numpy as np

from . import _capi
from .postfilter import combine_pair, gap_filter, local_clustering

logger = logging.getLogger(__name__)

_tls = threading.local()


def get_context(device: int = 0, max_chunks: int = 52) -> _capi.Context:
    """One cached engine context per (thread, device)."""
    pool = getattr(_tls, "pool", None)
    if pool is None:
        pool = _tls.pool = {}
    ctx = pool.get(device)
    if ctx is None or ctx.max_chunks < max_chunks:
        if ctx is not None:
            ctx.close()
        ctx = pool[device] = _capi.Context(device, max_chunks)
    return ctx


def pw_ww_pairs(pw, ww, maxww):
    """Sweep steps ordered by (w, p) -- callers.py:15-23."""
    steps = sorted((i, p) for p, w in zip(pw, ww) for i in range(w, maxww + 1))
    return [(p, i) for i, p in steps]


def lambdachunk(E):
    """Chunk edges and membership, same contract as callers.py:25-41 (kept for API parity; the
    engine bins on the GPU)."""
    E = np.asarray(E)
    if E.size == 0:
        return []
    numbin = int(np.ceil(np.log(E.max()) / np.log(2) * 3 + 1))
    out = []
    for i in range(1, numbin + 1):
        lv, rv = (0, 1) if i == 1 else (np.power(2, ((i - 2) / 3.)), np.power(2, ((i - 1) / 3.)))
        out.append((lv, rv, np.where((E > lv) & (E < rv))[0]))
    return out


def _as_diags(Diags, cDiags, IR, chromLen, num, min_ww):
    raw = _as_counts(Diags, num, chromLen)
    if _capi.first_nonconforming(cDiags, num - min_ww, chromLen - min_ww, -1, 8, "f") < 0 and len(cDiags) == num - min_ww:
        bal = cDiags                       # already what the engine reads: no per-diagonal Python work
    else:
        bal = [np.ascontiguousarray(c, dtype=np.float64) for c in cDiags]
    if len(bal) != num - min_ww:
        raise ValueError("cDiags must hold offsets min(ww)..num-1 (got %d arrays, expected %d)" % (len(bal), num - min_ww))
    keys = sorted(IR)
    if keys != list(range(min_ww, num)):
        raise ValueError("IR must be keyed by the offsets min(ww)..num-1")
    ir = np.array([IR[k] for k in keys], dtype=np.float64)
    return raw, bal, ir


def _numpy_numbin(e_max, n_valid):
    if n_valid <= 0:
        return 0
    return int(np.ceil(np.log(e_max) / np.log(2) * 3 + 1))


def _as_counts(Diags, num, chromLen=None):
    if chromLen is not None and _capi.first_nonconforming(Diags, num, chromLen, -1, 4, "i") < 0:
        return Diags                       # contiguous int32 diagonals of the right lengths (what `H.diagonal(d)` gives)
    raw = []
    for d in range(num):
        a = np.asarray(Diags[d])
        if a.dtype != np.int32:
            b = a.astype(np.int32)
            if not np.array_equal(b, a):
                raise ValueError("Diags[%d] holds values that are not int32 counts" % d)
            a = b
        raw.append(np.ascontiguousarray(a))
    return raw


def score_chromosome(ctx, chromLen, Diags, cDiags, IR, B1, B2, num, pw, ww, maxww, sig, maxapart_bins,
                     min_local_reads, chrom="", dump=False, weights=None):
    """Upload one chromosome band and run the GPU scoring path.  Returns (summary, survivors, gaps).
    ``weights`` given: worker-level input (``cDiags`` / ``IR`` / biases are derived on the GPU)."""
    min_ww = min(ww)
    if weights is not None:
        ctx.upload_counts(chromLen, num, min_ww, _as_counts(Diags, num, chromLen), weights)
    else:
        raw, bal, ir = _as_diags(Diags, cDiags, IR, chromLen, num, min_ww)
        ctx.upload(chromLen, num, min_ww, raw, bal, ir, B1, B2)
    P = ctx.make_params(pw, ww, maxww, sig, maxapart_bins, min_local_reads, dump=dump)
    try:
        S = ctx.score(P)
    except _capi.EngineError as e:
        if e.code == _capi.HP_ERR_EMPTY_REFIDX:
            # the reference dies here with a ValueError from the empty fancy index (callers.py:205-208)
            raise ValueError(str(e)) from None
        raise
    # the reference evaluates the chunk count with numpy; the engine used the C library -- make sure
    override = None
    nb = [[_numpy_numbin(S.lf[i][fl].e_max, S.lf[i][fl].n_valid) for fl in (0, 1)] for i in range(len(pw))]
    if any(nb[i][fl] != S.lf[i][fl].numbin for i in range(len(pw)) for fl in (0, 1)):
        override = np.array(nb, dtype=np.int32).ravel()
    if max(max(v) for v in nb) > ctx.max_chunks:
        raise _capi.EngineError(_capi.HP_ERR_CHUNK_OVERFLOW, "lambda-chunk count exceeds the context's max_chunks")
    S = ctx.fdr(override)
    return S, ctx.survivors(), ctx.gaps()


def _log_sweep(chrom, S):
    logger.info('Chrom:{0}, Observed Contact Number: {1}'.format(chrom, S.n_pixels))
    logger.info('Chrom:{0}, Two local neighborhoods, two expected matrices ...'.format(chrom))
    for k in range(S.n_steps):
        st = S.steps[k]
        logger.info('Chrom:{0},    Peak width:{1}, Donut width:{2}'.format(chrom, st.p, st.w))
        logger.info('Chrom:{0},    ({1},{2}) Valid Contact Number from This Loop: {3}'.format(chrom, st.p, st.w, st.resolved))
        logger.info('Chrom:{0},    ({1},{2}) Total Valid Ratio after This Loop: {3:.3f}'.format(chrom, st.p, st.w, 1 - st.left_ratio))


def hiccups_from_counts(weights, chromLen, Diags, num, chrom, **kw):
    """The reference's per-chromosome worker minus the cooler fetch (/root/reference/scripts/pyHICCUPS:146-173):
    takes the raw count diagonals and the bin weights, derives the balanced band, ``IR`` and the biases on the GPU
    (bit-identical to the worker's numpy code) and calls the peak caller.  Same keywords and result as ``hiccups``."""
    return hiccups(None, None, None, None, None, chromLen, Diags, None, num, chrom, weights=weights, **kw)


def hiccups(M, cM, B1, B2, IR, chromLen, Diags, cDiags, num, chrom, pw=[2], ww=[5],
            maxww=20, sig=0.1, sumq=0.01, double_fold=1.75, single_fold=2, maxapart=2000000,
            res=10000, use_raw=False, min_marginal_peaks=3, onlyanchor=True, min_local_reads=25,
            device=0, weights=None):
    """HiCCUPS peak calling for one chromosome.  ``M`` and ``cM`` are accepted for signature parity;
    the engine reads the same data from ``Diags`` / ``cDiags``.  Returns
    ``{(x_bp, y_bp): (cx_bp, cy_bp, radius_bp, O, foldK, pK, qK, foldY, pY, qY)}``."""
    pw, ww = list(pw), list(ww)
    try:
        S, sv, gaps = score_chromosome(get_context(device), chromLen, Diags, cDiags, IR, B1, B2, num, pw, ww, maxww, sig,
                                       maxapart // res, min_local_reads, chrom, weights=weights)
    except _capi.EngineError as e:
        if e.code != _capi.HP_ERR_CHUNK_OVERFLOW:
            raise
        # an expected value beyond 2^17 (the default table): once more with every lambda-chunk the engine supports (2^21)
        S, sv, gaps = score_chromosome(get_context(device, 64), chromLen, Diags, cDiags, IR, B1, B2, num, pw, ww, maxww, sig,
                                       maxapart // res, min_local_reads, chrom, weights=weights)
    _log_sweep(chrom, S)
    logger.info('Chrom:{0}, Poisson Models and Benjamini-Hochberg Correcting for lambda chunks ...'.format(chrom))
    for pi, (p, w) in enumerate(zip(pw, ww)):
        for fl in (0, 1):
            L = S.lf[pi][fl]
            logger.info('Chrom:{0},    ({1},{2}), Valid contact number: {3}'.format(chrom, p, w, L.n_valid))
            logger.info('Chrom:{0},    ({1},{2}), Number of chunks: {3}'.format(chrom, p, w, max(L.numbin, 0)))
    logger.info('Chrom:{0}, Perform greedy clustering and additional filtering ...'.format(chrom))
    return assemble_table(sv, gaps, chromLen, pw, ww, res, sumq, double_fold, single_fold, use_raw,
                          min_marginal_peaks, onlyanchor)


def bh_adjust(p, n, alpha):
    """statsmodels ``multipletests(method='fdr_bh')`` restricted to the ``p.size`` smallest of ``n`` p-values
    (all p-values that are not passed are larger than every one that is).  Returns (reject, q)."""
    order = np.argsort(p)
    ps = np.take(p, order)
    ecdf = np.arange(1, ps.size + 1) / float(n)
    rej = ps <= ecdf * alpha
    if rej.any():
        rej[:np.max(np.nonzero(rej)[0])] = True
    q = np.minimum.accumulate((ps / ecdf)[::-1])[::-1]
    q[q > 1] = 1
    reject = np.empty_like(rej)
    qv = np.empty_like(q)
    reject[order] = rej
    qv[order] = q
    return reject, qv


def bhfdr_from_counts(weights, chromLen, Diags, num, chrom, **kw):
    """Worker-level entry of the BH-FDR caller (scripts/pyBHFDR:112-146), see ``hiccups_from_counts``."""
    return bhfdr(None, None, None, None, None, chromLen, Diags, None, num, chrom, weights=weights, **kw)


def bhfdr(M, cM, B1, B2, IR, chromLen, Diags, cDiags, num, chrom, pw=2, ww=5, sig=0.05,
          maxww=20, maxapart=2000000, res=10000, min_marginal_peaks=3, onlyanchor=False, device=0, weights=None):
    """BH-FDR peak calling for one chromosome -- same contract as the reference's ``bhfdr``
    (/root/reference/hicpeaks/callers.py:364-590).  Returns ``{(x_bp, y_bp): (cx_bp, cy_bp, radius_bp, O, fold, p, q)}``.

    GPU: the donut sweep with the hard-coded ``Reads >= 16`` rule (:490), ``E`` (:526-535) and the per-pixel Poisson
    tail (:536-540) for every pixel that can still pass ``sig``.  Host (a few thousand records): the chromosome-wide
    Benjamini-Hochberg step (:545-547), gap filter (:557-577), clustering (:580-582) and ``fold > 2`` (:587)."""
    for max_chunks in (52, 64):
        ctx = get_context(device, max_chunks)
        if weights is not None:
            ctx.upload_counts(chromLen, num, ww, _as_counts(Diags, num, chromLen), weights)
        else:
            raw, bal, ir = _as_diags(Diags, cDiags, IR, chromLen, num, ww)
            ctx.upload(chromLen, num, ww, raw, bal, ir, B1, B2)
        P = ctx.make_params([pw], [ww], maxww, sig, maxapart // res, 16, bhfdr=True)
        try:
            S = ctx.score(P)
        except _capi.EngineError as e:
            if e.code == _capi.HP_ERR_EMPTY_REFIDX:
                raise ValueError(str(e)) from None
            if e.code == _capi.HP_ERR_CHUNK_OVERFLOW and max_chunks < 64:
                continue                  # the engine bins E for its candidate test: once more with its full table (E up to 2^21)
            raise
        break
    _log_sweep(chrom, S)
    S = ctx.fdr()
    n_tests = int(S.lf[0][0].n_valid)
    logger.info('Chrom:{0}, Number of Poisson Models: {1}'.format(chrom, n_tests))
    sv = ctx.survivors()
    sv = sv[np.lexsort((sv["c"], sv["r"]))]
    table = {}
    if sv.size == 0:
        return table
    p = np.ascontiguousarray(sv["p"][:, 0])
    reject, q = bh_adjust(p, n_tests, sig)
    sv, p, q = sv[reject], p[reject], q[reject]
    keep = gap_filter(sv["r"], sv["c"], ctx.gaps(), ww, chromLen)
    sv, p, q = sv[keep], p[keep], q[keep]
    x, y = sv["r"].astype(np.int64), sv["c"].astype(np.int64)
    fold = sv["obs"] / sv["e"][:, 0]
    Donuts = {(int(a), int(b)): (o, f, pp, qq) for a, b, o, f, pp, qq in zip(x, y, sv["obs"], fold, p, q)}
    for pixel, cen, radius in local_clustering(Donuts, None, res, min_count=min_marginal_peaks, r=2 * res,
                                               onlysummit=onlyanchor):
        donut = Donuts[pixel]
        if donut[1] > 2:
            table[(pixel[0] * res, pixel[1] * res)] = (cen[0] * res, cen[1] * res) + (radius * res,) + donut
    return table


def assemble_table(sv, gaps, chromLen, pw, ww, res, sumq, double_fold, single_fold, use_raw,
                   min_marginal_peaks, onlyanchor):
    """callers.py:289-362 on the engine's survivor records (``_capi.SURVIVOR_DTYPE``) and gap mask."""
    sv = sv[np.lexsort((sv["c"], sv["r"]))]
    m = min(ww)
    pixel_table = {}
    for pi in range(len(pw)):
        mine = sv[sv["pair"] == pi]
        side = []
        for fl, rbit in ((0, _capi.SF_REJECT_K), (1, _capi.SF_REJECT_Y)):
            s = mine[(mine["flags"] & rbit) != 0]
            s = s[gap_filter(s["r"], s["c"], gaps, m, chromLen)]
            side.append(dict(x=s["r"].astype(np.int64), y=s["c"].astype(np.int64), ice=s["ice"], obs=s["obs"],
                             fold=s["obs"] / s["e"][:, fl], p=s["p"][:, fl], q=s["q"][:, fl]))
        nz = mine[(mine["flags"] & _capi.SF_CEMY_NONZERO) != 0]
        cemy_nonzero = set(zip(nz["r"].astype(np.int64), nz["c"].astype(np.int64)))
        combine_pair(pixel_table, res, side[0], side[1], cemy_nonzero, double_fold, single_fold, use_raw)

    Donuts = {(k[0] // res, k[1] // res): pixel_table[k][3:8] for k in pixel_table}
    LL = {(k[0] // res, k[1] // res): pixel_table[k][8:] for k in pixel_table}
    peaks = local_clustering(Donuts, LL, res, min_count=min_marginal_peaks, r=2 * res, sumq=sumq,
                             onlysummit=onlyanchor)
    final_table = {}
    for pixel, cen, radius in peaks:
        key = (pixel[0] * res, pixel[1] * res)
        final_table[key] = (cen[0] * res, cen[1] * res) + (radius * res,) + pixel_table[key][4:]
    return final_table
