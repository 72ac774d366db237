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
