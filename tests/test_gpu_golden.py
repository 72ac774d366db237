"""GPU: the drop-in ``hicpeaks_b200.callers.hiccups`` on the committed golden inputs against the
outputs of the unmodified reference (peak coordinates exact, values within the stated tolerance)."""
import numpy as np
import pytest

import golden_util as gu
from hicpeaks_b200 import _capi, callers

pytestmark = pytest.mark.gpu

Q_TOL = 1e-6          # BASELINE.json north_star: q-values within 1e-6


@pytest.mark.parametrize("name", gu.names("hiccups"))
def test_hiccups_matches_reference(name):
    z, inp, kw, res = gu.load(name)
    args = (None, None, inp["biases"], inp["biases"], dict(inp["IR"]), inp["n"], inp["Diags"], inp["cDiags"],
            inp["num"], "21")
    if str(z["raises"]):
        with pytest.raises(ValueError):           # same failure mode as the reference (callers.py:205-208)
            callers.hiccups(*args, res=res, **kw)
        return
    table = callers.hiccups(*args, res=res, **kw)
    got, exp = gu.table_rows(table), z["table"]
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert np.array_equal(got[:, :6], exp[:, :6])                       # pixel, centroid, radius, O: exact
    assert np.allclose(got[:, 6], exp[:, 6], rtol=1e-12, atol=0)        # fold (E is bit-exact)
    assert np.allclose(got[:, 9], exp[:, 9], rtol=1e-12, atol=0)
    for col in (7, 8, 10, 11):                                          # p, q
        assert np.abs(got[:, col] - exp[:, col]).max() <= Q_TOL if got.size else True


@pytest.mark.parametrize("name", [n for n in gu.names("hiccups") if "crash" not in n])
def test_survivors_match_reference(name):
    z, inp, kw, res = gu.load(name)
    pw, ww = kw["pw"], kw["ww"]
    ctx = callers.get_context(0)
    S, sv, gaps = callers.score_chromosome(ctx, inp["n"], inp["Diags"], inp["cDiags"], inp["IR"], inp["biases"],
                                           inp["biases"], inp["num"], pw, ww, kw["maxww"], kw["sig"],
                                           kw["maxapart"] // res, kw["min_local_reads"])
    assert S.n_pixels == int(z["n_pixels"])
    for pi, p in enumerate(pw):
        for fl, nm, rbit in ((0, "K", _capi.SF_REJECT_K), (1, "Y", _capi.SF_REJECT_Y)):
            exp = z["p%d%s_surv" % (p, nm)]
            s = sv[(sv["pair"] == pi) & ((sv["flags"] & rbit) != 0)]
            s = s[np.lexsort((s["c"], s["r"]))]
            assert np.array_equal(s["r"], exp[0]) and np.array_equal(s["c"], exp[1])
            assert np.array_equal(s["obs"], exp[2])
            assert np.array_equal(s["e"][:, fl], exp[3])                # expected values: bit-exact
            if s.size:
                assert np.abs(s["p"][:, fl] - exp[4]).max() <= Q_TOL
                assert np.abs(s["q"][:, fl] - exp[5]).max() <= Q_TOL
            assert S.lf[pi][fl].n_valid == int(z["p%d%s_n_valid" % (p, nm)])
            assert S.lf[pi][fl].numbin == int(z["p%d%s_numbin" % (p, nm)])


@pytest.mark.parametrize("scope", ["chrom", "genome"])
def test_genome_runner_matches_hiccups(scope):
    """The dispatcher (chromosome sharding; optional merged-histogram FDR) on one rank: with a single chromosome both
    scopes must reproduce the drop-in hiccups() table, i.e. the reference's."""
    from hicpeaks_b200 import dispatch
    name = "chr21_25k_p1w3"
    z, inp, kw, res = gu.load(name)
    runner = dispatch.GenomeRunner(engine=dispatch.CudaEngine(0), fdr_scope=scope)
    out = runner.run({"21": lambda: inp}, {"21": (inp["n"], inp["num"])}, res=res, **kw)
    got, exp = gu.table_rows(out["21"]), z["table"]
    assert got.shape == exp.shape
    assert np.array_equal(got[:, :6], exp[:, :6])
    for col in (7, 8, 10, 11):
        assert np.abs(got[:, col] - exp[:, col]).max() <= Q_TOL


def test_genome_scope_merges_histograms():
    """Two chromosomes, genome scope: every context runs BH on the summed histogram (checked through the chunk tables)."""
    from hicpeaks_b200 import dispatch
    from hicpeaks_b200.synth import synth_chromosome
    inps = {c: synth_chromosome(500, 60, 5, maxww=10, seed=s, scale=60.0) for c, s in (("a", 3), ("b", 4))}
    prm = dict(dispatch.DEFAULTS, pw=[2], ww=[5], maxww=10, sig=0.1, maxapart=60 * 10000, res=10000, min_local_reads=16)
    eng = dispatch.CudaEngine(0)
    hs = {c: eng.score(c, inps[c], prm) for c in inps}
    hists = {c: eng.hist(hs[c]) for c in inps}
    total = hists["a"] + hists["b"]
    assert total.sum() == hists["a"].sum() + hists["b"].sum() > 0
    hs["a"]["ctx"].hist_import(total)
    assert np.array_equal(hs["a"]["ctx"].hist_export(), total)
    for h in hs.values():
        h["ctx"].close()


@pytest.mark.parametrize("name", gu.names("bhfdr"))
def test_bhfdr_matches_reference(name):
    """pyBHFDR's operator: donut sweep + per-pixel Poisson tails on the GPU, chromosome-wide BH on the host."""
    z, inp, kw, res = gu.load(name)
    table = callers.bhfdr(None, None, inp["biases"], inp["biases"], dict(inp["IR"]), inp["n"], inp["Diags"], inp["cDiags"],
                          inp["num"], "21", res=res, **kw)
    got = np.array([list(k) + [float(v) for v in table[k]] for k in sorted(table)], dtype=np.float64).reshape(len(table), 9)
    exp = z["table"]
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert np.array_equal(got[:, :6], exp[:, :6])                       # pixel, centroid, radius, O: exact
    assert np.allclose(got[:, 6], exp[:, 6], rtol=1e-12, atol=0)        # fold (E is bit-exact)
    assert np.abs(got[:, 7] - exp[:, 7]).max() <= Q_TOL and np.abs(got[:, 8] - exp[:, 8]).max() <= Q_TOL
    # the rejected set before the gap filter / clustering
    ctx = callers.get_context(0)
    sv = ctx.survivors()
    sv = sv[np.lexsort((sv["c"], sv["r"]))]
    reject, q = callers.bh_adjust(np.ascontiguousarray(sv["p"][:, 0]), int(z["n_tests"]), kw["sig"])
    s = sv[reject]
    surv = z["surv"]
    assert np.array_equal(s["r"], surv[0]) and np.array_equal(s["c"], surv[1])
    assert np.array_equal(s["e"][:, 0], surv[3])
    assert np.abs(s["p"][:, 0] - surv[4]).max() <= Q_TOL and np.abs(q[reject] - surv[5]).max() <= Q_TOL


@pytest.mark.parametrize("name", ["chr21_25k_p1w3", "synth_p2w5", "synth_union_124", "synth_p4w7"])
def test_device_input_prep_is_bit_identical(name):
    """K0: balanced band, IR (numpy's pairwise mean over the non-NaN entries) and biases derived on the GPU from raw
    counts + weights equal the worker's host arrays bit for bit (scripts/pyHICCUPS:143-166)."""
    z, inp, kw, res = gu.load(name)
    n, num, mw = inp["n"], inp["num"], inp["min_ww"]
    ctx = callers.get_context(0)
    ctx.upload_counts(n, num, mw, [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]], inp["weights"])
    ir = ctx.dump_band(0)
    exp = np.array([inp["IR"][d] for d in range(mw, num)])
    assert np.array_equal(ir[mw:], exp, equal_nan=True) and np.all(ir[:mw] == 0)
    assert np.array_equal(ctx.dump_band(1), inp["biases"])
    bal = ctx.dump_band(2)
    for i, d in enumerate(range(mw, num)):
        assert np.array_equal(bal[d, : n - d], inp["cDiags"][i]), d
    assert np.all(bal[:mw] == 0)
    if str(z["raises"]):
        return
    a = callers.hiccups_from_counts(inp["weights"], n, inp["Diags"], num, "21", res=res, **kw)
    b = callers.hiccups(None, None, inp["biases"], inp["biases"], dict(inp["IR"]), n, inp["Diags"], inp["cDiags"], num, "21",
                        res=res, **kw)
    assert a == b and np.array_equal(gu.table_rows(a)[:, :6], z["table"][:, :6])
