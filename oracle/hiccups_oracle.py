"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference HiCCUPS scoring path.

Nothing in the product (``hicpeaks_b200/``) may import this module; it is the *checker* used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  It restates, on dense diagonal-major band arrays and with the reference's exact
floating-point summation order, what these pieces of the reference compute:

* ``pw_ww_steps``        <- /root/reference/hicpeaks/callers.py:15-23   (``pw_ww_pairs``)
* ``step_offsets``       <- callers.py:138-198 (masks :138-141, skip rule :150-152, add rules :179-198;
                            the ``else: subtract`` branches :183-185,:192-196 are unreachable)
* ``sweep``              <- callers.py:101-232 (pixel set :101-104, accumulate :132-198, resolve :203-217,
                            adaptive stop ``frozen_w`` :219-232)
* ``expected_and_fdr``   <- callers.py:239-287 (E :244-256, lambda chunks :25-41, Poisson :268-270, BH :273-279)
* ``bh_fdr``             <- statsmodels ``multipletests(method='fdr_bh')`` (third-party, unpinned; see
                            oracle/shims/statsmodels/sandbox/stats/multicomp.py)

Parity status: PINNED -- ``oracle/make_golden.py`` runs the unmodified reference (imported by path from
/root/reference with the statsmodels shim) in the build container, and ``tests/test_oracle_golden.py``
checks this restatement bit-for-bit against the committed outputs under ``tests/golden/``.

Layout: ``X[d, r]`` holds matrix element ``(r, r + d)``; ``shift(X, a, b)[r, c] = X[r + a, c + b]``
is ``X[d + b - a, r + a]``.
"""
from __future__ import annotations

import numpy as np
from scipy.special import pdtr

K, Y = 0, 1


class EmptyRefIdx(ValueError):
    """The reference raises (callers.py:205-208) when the unresolved set of some p is empty at a step."""


def pw_ww_steps(pw, ww, maxww):
    pool = sorted((i, p) for p, w in zip(pw, ww) for i in range(w, maxww + 1))
    return [(p, i) for i, p in pool]


def step_offsets(p, w, limit, last_p, last_w, min_pw):
    """Ordered (a, b, isY, isR) list one sweep step adds; order = the fp64 addition order."""
    ops = []
    for a in range(-w, w + 1):
        for b in range(-w, w + 1):
            g = max(abs(a), abs(b))
            if limit and ((g <= last_w and g > max(p, last_p)) or g <= min(p, last_p)):
                continue
            if a == 0 or b == 0:
                continue
            if abs(a) <= p and abs(b) <= p:
                continue
            is_y = a > 0 and b < 0
            is_r = is_y and ((not limit) or (p == min_pw and g > last_w))
            ops.append((a, b, is_y, is_r))
    return ops


def step_program(pw, ww, maxww):
    """[(p, w, ops)] for every step in execution order, assuming none is skipped by frozen_w
    (skipped steps are always a suffix because steps are sorted by w)."""
    prog = []
    limit = False
    last_p = last_w = 0
    for p, w in pw_ww_steps(pw, ww, maxww):
        prog.append((p, w, step_offsets(p, w, limit, last_p, last_w, min(pw))))
        limit = True
        last_p, last_w = p, w
    return prog


def dense_band(inp):
    """(raw[num, n] int64, bal[num, n] f64, eb[num, n] f64) zero outside the stored band / chromosome."""
    n, num, mw = inp["n"], inp["num"], inp["min_ww"]
    raw = np.zeros((num, n), dtype=np.int64)
    bal = np.zeros((num, n), dtype=np.float64)
    eb = np.zeros((num, n), dtype=np.float64)
    for d in range(num):
        raw[d, : n - d] = inp["Diags"][d]
    for i, d in enumerate(range(mw, num)):
        bal[d, : n - d] = inp["cDiags"][i]
        eb[d, : n - d] = inp["IR"][d]
    return raw, bal, eb


def _add_shift(acc, X, a, b):
    num, n = X.shape
    dd = b - a
    dlo, dhi = max(0, -dd), min(num, num - dd)
    rlo, rhi = max(0, -a), min(n, n - a)
    if dlo < dhi and rlo < rhi:
        acc[dlo:dhi, rlo:rhi] += X[dlo + dd:dhi + dd, rlo + a:rhi + a]


def sweep(inp, pw, ww, maxww, min_local_reads, maxapart_bins, log=None):
    """Accumulate / resolve loop.  Returns dict with pixel coordinates (row-major), per-p snapshots,
    the step index at which each pixel's Reads first reached the threshold, and frozen_w."""
    n, num = inp["n"], inp["num"]
    raw, bal, eb = dense_band(inp)
    d_idx = np.arange(num)[:, None]
    r_idx = np.arange(n)[None, :]
    pix = (raw != 0) & (d_idx >= min(ww)) & (d_idx <= maxapart_bins) & (r_idx + d_idx < n)
    vx, vd = np.nonzero(pix.T)                   # row-major over (r, c)
    total = vx.size
    prog = step_program(pw, ww, maxww)

    bS = [np.zeros((num, n)), np.zeros((num, n))]
    bE = [np.zeros((num, n)), np.zeros((num, n))]
    reads = np.zeros((num, n), dtype=np.int64)
    bSV = {p: [np.zeros(total), np.zeros(total)] for p in pw}
    bEV = {p: [np.zeros(total), np.zeros(total)] for p in pw}
    ref = {p: np.arange(total) for p in pw}
    ini = {p: total for p in pw}
    first_step = np.full(total, len(prog), dtype=np.int64)     # s*: first step with Reads >= thr
    res_w = {p: np.zeros(total, dtype=np.int64) for p in pw}   # w at which p resolved (0 = never)
    frozen = maxww
    executed = []
    for s, (p, w, ops) in enumerate(prog):
        if w > frozen:
            continue
        executed.append(s)
        for a, b, is_y, is_r in ops:
            _add_shift(bS[K], bal, a, b)
            _add_shift(bE[K], eb, a, b)
            if is_y:
                _add_shift(bS[Y], bal, a, b)
                _add_shift(bE[Y], eb, a, b)
            if is_r:
                _add_shift(reads, raw, a, b)
        T = ref[p]
        if T.size == 0:
            raise EmptyRefIdx("unresolved set of p=%d is empty at step (%d,%d)" % (p, p, w))
        rn = reads[vd[T], vx[T]]
        okm = rn >= min_local_reads
        hit = T[okm]
        allr = reads[vd, vx] >= min_local_reads
        first_step[allr & (first_step == len(prog))] = s
        for fl in (K, Y):
            bSV[p][fl][hit] = bS[fl][vd[hit], vx[hit]]
            bEV[p][fl][hit] = bE[fl][vd[hit], vx[hit]]
        res_w[p][hit] = w
        valid = hit.size / float(ini[p])
        ref[p] = T[~okm]
        ini[p] = ref[p].size
        left = ini[p] / float(total)
        if log is not None:
            log.append((p, w, int(hit.size), valid, left))
        if w >= max(ww) and (valid < 0.3 or left < 0.03):
            frozen = w
    return dict(vx=vx, vy=vx + vd, vd=vd, bSV=bSV, bEV=bEV, first_step=first_step, res_w=res_w,
                frozen=frozen, executed=executed, nsteps=len(prog), total=total, raw=raw, bal=bal)


def chunk_edges(numbin):
    """(lv, rv) per chunk i=1..numbin -- callers.py:32-37."""
    out = []
    for i in range(1, numbin + 1):
        if i == 1:
            out.append((0, 1))
        else:
            out.append((np.power(2, ((i - 2) / 3.)), np.power(2, ((i - 1) / 3.))))
    return out


def bh_fdr(p):
    n = p.size
    order = np.argsort(p)
    ps = np.take(p, order)
    ecdf = np.arange(1, n + 1) / float(n)
    q = np.minimum.accumulate((ps / ecdf)[::-1])[::-1]
    q[q > 1] = 1
    out = np.empty_like(q)
    out[order] = q
    return out


def expected_and_fdr(inp, sw, p, w0, fl, sig):
    """callers.py:244-287 for one (p, fl).  Returns row-major arrays of the valid pixels and the
    survivors mask (q <= sig)."""
    vx, vy, vd = sw["vx"], sw["vy"], sw["vd"]
    B = inp["biases"]
    ir = np.zeros(inp["num"])
    for d, v in inp["IR"].items():
        ir[d] = v
    bs, be = sw["bSV"][p][fl], sw["bEV"][p][fl]
    m = (be != 0) & (vd >= w0)
    x, y, d = vx[m], vy[m], vd[m]
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = bs[m] / be[m]
        cem = ir[d] * ratio
        keep0 = (cem != 0) & (ratio != 0)          # lil assignment of a 0 ratio stores nothing
        E = cem * B[x] * B[y]
        keep = keep0 & (E > 0)
    cem_nz = (x[keep0], y[keep0])                  # where the reference's cEM matrix has a non-zero entry
    x, y, d, E, cem = x[keep], y[keep], d[keep], E[keep], cem[keep]
    O = sw["raw"][d, x].astype(np.float64)
    ice = sw["bal"][d, x]
    fold = O / E
    pv = np.ones(x.size)
    qv = np.ones(x.size)
    chunk = np.zeros(x.size, dtype=np.int64)       # 0 = in no chunk
    numbin = 0
    if E.size:
        numbin = int(np.ceil(np.log(E.max()) / np.log(2) * 3 + 1))
        for i, (lv, rv) in enumerate(chunk_edges(numbin), start=1):
            idx = np.where((E > lv) & (E < rv))[0]
            if idx.size:
                cp = 1 - pdtr(np.floor(O[idx]), rv)
                pv[idx] = cp
                qv[idx] = bh_fdr(cp)
                chunk[idx] = i
    return dict(x=x, y=y, E=E, O=O, ice=ice, fold=fold, p=pv, q=qv, chunk=chunk, numbin=numbin,
                reject=qv <= sig, cem_nz=cem_nz)


def score(inp, pw, ww, maxww=20, sig=0.1, maxapart_bins=200, min_local_reads=25):
    """Whole scoring path (rows a-2 .. a-6 of SURVEY 8a).  Returns (sweep dict, {(p, fl): result})."""
    sw = sweep(inp, pw, ww, maxww, min_local_reads, maxapart_bins)
    res = {}
    for p, w0 in zip(pw, ww):
        for fl in (K, Y):
            res[(p, fl)] = expected_and_fdr(inp, sw, p, w0, fl, sig)
    return sw, res


def score_bhfdr(inp, pw, ww, maxww=20, sig=0.05, maxapart_bins=200):
    """callers.py:364-553 (``bhfdr``): the donut sweep of one (pw, ww) pair with the hard-coded Reads >= 16 rule (:490)
    and ``break`` on valid < 0.3 or left < 0.03 (:505-511) -- for a single pair the same executed steps as ``sweep`` --
    then per-pixel Poisson tails with the pixel's own rate (:536-540) and one BH over the chromosome (:545-547)."""
    sw = sweep(inp, [pw], [ww], maxww, 16, maxapart_bins)
    vx, vy, vd = sw["vx"], sw["vy"], sw["vd"]
    B = inp["biases"]
    ir = np.zeros(inp["num"])
    for d, v in inp["IR"].items():
        ir[d] = v
    bs, be = sw["bSV"][pw][K], sw["bEV"][pw][K]
    m = be != 0
    x, y, d = vx[m], vy[m], vd[m]
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = bs[m] / be[m]
        cem = ir[d] * ratio
        E = cem * B[x] * B[y]
        keep = (cem != 0) & (ratio != 0) & (E > 0)
    x, y, d, E = x[keep], y[keep], d[keep], E[keep]
    O = sw["raw"][d, x].astype(np.float64)
    p = 1 - pdtr(np.floor(O), E)
    n = p.size
    order = np.argsort(p)
    ps = np.take(p, order)
    ecdf = np.arange(1, n + 1) / float(n)
    rej = ps <= ecdf * sig
    if rej.any():
        rej[:np.max(np.nonzero(rej)[0])] = True
    reject = np.empty(n, dtype=bool)
    reject[order] = rej
    return sw, dict(x=x, y=y, E=E, O=O, p=p, q=bh_fdr(p) if n else p, reject=reject, fold=O / E)
