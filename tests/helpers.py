"""Shared test plumbing: run the CUDA engine through the C ABI and the oracle on the same inputs."""
from __future__ import annotations

import numpy as np

from hicpeaks_b200 import _capi
from oracle import hiccups_oracle as ho


def engine_inputs(inp):
    Diags = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
    cDiags = [np.ascontiguousarray(c, dtype=np.float64) for c in inp["cDiags"]]
    mw, num = inp["min_ww"], inp["num"]
    ir = np.array([inp["IR"][d] for d in range(mw, num)], dtype=np.float64)
    return Diags, cDiags, ir


def run_engine(ctx, inp, pw, ww, maxww, sig, maxapart_bins, min_local_reads, dump=True, generic_kernel=False):
    Diags, cDiags, ir = engine_inputs(inp)
    ctx.upload(inp["n"], inp["num"], inp["min_ww"], Diags, cDiags, ir, inp["biases"], inp["biases"])
    P = ctx.make_params(pw, ww, maxww, sig, maxapart_bins, min_local_reads, dump=dump, generic_kernel=generic_kernel)
    S1 = ctx.score(P)
    S = ctx.fdr()
    return S1, S


def compare_with_oracle(ctx, inp, pw, ww, maxww, sig, maxapart_bins, min_local_reads, q_tol=1e-6, generic_kernel=False):
    """Full cut-point comparison (levels, bS/bE/E bit-exact, histograms exact, p/q within q_tol,
    survivor coordinates exact).  Returns a dict of small statistics."""
    sw, res = ho.score(inp, pw, ww, maxww=maxww, sig=sig, maxapart_bins=maxapart_bins,
                       min_local_reads=min_local_reads)
    S1, S = run_engine(ctx, inp, pw, ww, maxww, sig, maxapart_bins, min_local_reads, dump=True,
                       generic_kernel=generic_kernel)
    stats = {"spec_kernel": int(S1.spec_kernel)}
    vx, vd = sw["vx"], sw["vd"]
    # --- levels (a-4) ---
    assert S.n_pixels == sw["total"]
    assert S.frozen_w == sw["frozen"], (S.frozen_w, sw["frozen"])
    assert S.n_steps == len(sw["executed"])
    lv = ctx.dump_levels()
    got = lv[vd, vx].astype(np.int64)
    exp = sw["first_step"].copy()
    # the oracle only observes steps it executed; beyond them the engine's level is >= nexec or never
    seen = exp < sw["nsteps"]
    assert np.array_equal(got[seen], exp[seen])
    assert np.all((got[~seen] >= len(sw["executed"])))
    mask = np.ones_like(lv, dtype=bool)
    mask[vd, vx] = False
    d_idx = np.arange(inp["num"])[:, None]
    inband = (d_idx >= min(ww)) & (d_idx <= min(maxapart_bins, inp["num"] - 1))
    assert np.all(lv[mask & inband] == 0xFF)
    # --- sums and expected values (a-3, a-5) ---
    max_dq = 0.0
    max_dp = 0.0
    for pi, (p, w0) in enumerate(zip(pw, ww)):
        for fl in (0, 1):
            bS = ctx.dump_plane(pi, fl, 0)[vd, vx]
            bE = ctx.dump_plane(pi, fl, 1)[vd, vx]
            Ep = ctx.dump_plane(pi, fl, 2)
            sel = (sw["res_w"][p] > 0) & (vd >= w0)
            assert np.array_equal(bS[sel], sw["bSV"][p][fl][sel]), "bS differs (pair %d fl %d)" % (pi, fl)
            assert np.array_equal(bE[sel], sw["bEV"][p][fl][sel]), "bE differs (pair %d fl %d)" % (pi, fl)
            assert np.all(np.isnan(bS[~sel]))
            r = res[(p, fl)]
            assert np.array_equal(Ep[r["y"] - r["x"], r["x"]], r["E"]), "E differs"
            with np.errstate(invalid="ignore"):
                assert int((Ep > 0).sum()) == r["x"].size
            L = S.lf[pi][fl]
            assert L.n_valid == r["x"].size
            assert L.e_max == (r["E"].max() if r["E"].size else 0.0)
            assert L.numbin == r["numbin"]
            # --- histogram / p / q tables (a-6) ---
            nb, widths, off, hist, ptab, qtab = ctx.chunk_table(pi, fl)
            assert nb == max(0, r["numbin"])
            for ci in range(1, nb + 1):
                m = r["chunk"] == ci
                W = int(widths[ci - 1])
                kb = np.minimum(r["O"][m].astype(np.int64), W - 1)
                assert np.array_equal(np.bincount(kb, minlength=W), hist[off[ci - 1]:off[ci]]), "hist chunk %d" % ci
                if m.any():
                    dp = np.abs(ptab[off[ci - 1] + kb] - r["p"][m]).max()
                    dq = np.abs(qtab[off[ci - 1] + kb] - r["q"][m]).max()
                    max_dp, max_dq = max(max_dp, dp), max(max_dq, dq)
            assert L.n_reject == int(r["reject"].sum())
    assert max_dq <= q_tol, max_dq
    stats["max_dp"], stats["max_dq"] = max_dp, max_dq
    # --- survivors ---
    sv = ctx.survivors()
    for pi, (p, w0) in enumerate(zip(pw, ww)):
        for fl, (vbit, rbit) in enumerate(((_capi.SF_VALID_K, _capi.SF_REJECT_K), (_capi.SF_VALID_Y, _capi.SF_REJECT_Y))):
            r = res[(p, fl)]
            s = sv[(sv["pair"] == pi) & ((sv["flags"] & rbit) != 0)]
            s = s[np.lexsort((s["c"], s["r"]))]
            rej = r["reject"]
            assert np.array_equal(s["r"], r["x"][rej]) and np.array_equal(s["c"], r["y"][rej]), "survivor coordinates"
            assert np.array_equal(s["e"][:, fl], r["E"][rej])
            assert np.array_equal(s["obs"], r["O"][rej])
            assert np.array_equal(s["ice"], r["ice"][rej])
            if s.size:
                assert np.abs(s["q"][:, fl] - r["q"][rej]).max() <= q_tol
                assert np.abs(s["p"][:, fl] - r["p"][rej]).max() <= q_tol
    stats["n_survivors"] = int(sv.size)
    stats["frozen"] = int(S.frozen_w)
    stats["n_pixels"] = int(S.n_pixels)
    # --- gaps (callers.py:238) ---
    bal = sw["bal"]
    gaps = bal.sum(axis=0) == 0
    assert np.array_equal(ctx.gaps(), gaps)
    return stats


def compare_survivor_path_with_oracle(ctx, inp, pw, ww, maxww, sig, maxapart_bins, min_local_reads, q_tol=1e-6, counts=False,
                                      expect_fast=None):
    """The cut points of ``compare_with_oracle`` that do not need the per-pixel dump planes: pixel count, frozen width,
    valid counts, E.max, numbin, (chunk, O) histograms, rejected counts, survivor coordinates / O / ice / E (bit-exact)
    and p / q within ``q_tol``.  Without the dump the engine takes its default path -- for single-pair programs the
    re-associated kernel with exact re-evaluation -- so this is the oracle check of that path."""
    sw, res = ho.score(inp, pw, ww, maxww=maxww, sig=sig, maxapart_bins=maxapart_bins, min_local_reads=min_local_reads)
    Diags, cDiags, ir = engine_inputs(inp)
    if counts:
        ctx.upload_counts(inp["n"], inp["num"], inp["min_ww"], Diags, inp["weights"])
    else:
        ctx.upload(inp["n"], inp["num"], inp["min_ww"], Diags, cDiags, ir, inp["biases"], inp["biases"])
    P = ctx.make_params(pw, ww, maxww, sig, maxapart_bins, min_local_reads)
    S1 = ctx.score(P)
    S = ctx.fdr()
    if expect_fast is not None:
        assert bool(S1.fast_kernel) == expect_fast
    assert S.n_pixels == sw["total"] and S.frozen_w == sw["frozen"] and S.n_steps == len(sw["executed"])
    sv = ctx.survivors()
    for pi, (p, w0) in enumerate(zip(pw, ww)):
        for fl, rbit in enumerate((_capi.SF_REJECT_K, _capi.SF_REJECT_Y)):
            r = res[(p, fl)]
            L = S.lf[pi][fl]
            assert L.n_valid == r["x"].size
            assert L.e_max == (r["E"].max() if r["E"].size else 0.0)
            assert L.numbin == r["numbin"]
            nb, widths, off, hist, ptab, qtab = ctx.chunk_table(pi, fl)
            assert nb == max(0, r["numbin"])
            for ci in range(1, nb + 1):
                m = r["chunk"] == ci
                W = int(widths[ci - 1])
                kb = np.minimum(r["O"][m].astype(np.int64), W - 1)
                assert np.array_equal(np.bincount(kb, minlength=W), hist[off[ci - 1]:off[ci]]), "hist chunk %d" % ci
            assert L.n_reject == int(r["reject"].sum())
            s = sv[(sv["pair"] == pi) & ((sv["flags"] & rbit) != 0)]
            s = s[np.lexsort((s["c"], s["r"]))]
            rej = r["reject"]
            assert np.array_equal(s["r"], r["x"][rej]) and np.array_equal(s["c"], r["y"][rej]), "survivor coordinates"
            assert np.array_equal(s["e"][:, fl], r["E"][rej]), "survivor E is not the reference's fp64 value"
            assert np.array_equal(s["obs"], r["O"][rej]) and np.array_equal(s["ice"], r["ice"][rej])
            if s.size:
                assert np.abs(s["q"][:, fl] - r["q"][rej]).max() <= q_tol
                assert np.abs(s["p"][:, fl] - r["p"][rej]).max() <= q_tol
    return dict(n_pixels=int(S.n_pixels), frozen=int(S.frozen_w), n_survivors=int(sv.size), fast=int(S1.fast_kernel),
                n_exact=int(S1.n_exact), ms_score=float(S1.ms_score))
