"""CPU: the oracle restatement against the committed outputs of the unmodified reference."""
import numpy as np
import pytest

import golden_util as gu
from oracle import glue_oracle, hiccups_oracle as ho


@pytest.mark.parametrize("name", gu.names("hiccups"))
def test_oracle_reproduces_reference(name):
    z, inp, kw, res = gu.load(name)
    pw, ww = kw["pw"], kw["ww"]
    args = dict(maxww=kw["maxww"], sig=kw["sig"], maxapart_bins=kw["maxapart"] // res,
                min_local_reads=kw["min_local_reads"])
    if str(z["raises"]):
        with pytest.raises(ValueError):
            ho.score(inp, pw, ww, **args)
        return
    sw, out = ho.score(inp, pw, ww, **args)
    assert sw["total"] == int(z["n_pixels"])
    assert gu.sha(np.stack([sw["vx"], sw["vy"]]).astype(np.int64)) == str(z["sha_pixels"])
    for p in pw:
        for fl, nm in ((0, "K"), (1, "Y")):
            pre = "p%d%s_" % (p, nm)
            r = out[(p, fl)]
            assert gu.sha(sw["bSV"][p][fl]) == str(z[pre + "sha_bSV"]), "bSV bits"
            assert gu.sha(sw["bEV"][p][fl]) == str(z[pre + "sha_bEV"]), "bEV bits"
            assert r["x"].size == int(z[pre + "n_valid"])
            assert r["numbin"] == int(z[pre + "numbin"])
            assert gu.sha(np.stack([r["x"], r["y"]]).astype(np.int64)) == str(z[pre + "sha_xy"])
            assert gu.sha(r["E"]) == str(z[pre + "sha_E"]), "E bits"
            assert gu.sha(r["chunk"].astype(np.int64)) == str(z[pre + "sha_chunk"])
            assert gu.sha(r["p"]) == str(z[pre + "sha_p"]), "p bits"
            assert gu.sha(r["q"]) == str(z[pre + "sha_q"]), "q bits"
            rej = r["reject"]
            surv = np.stack([r["x"][rej], r["y"][rej], r["O"][rej], r["E"][rej], r["p"][rej], r["q"][rej]])
            assert np.array_equal(surv, z[pre + "surv"])
    final = glue_oracle.finish_hiccups(inp, sw, out, pw, ww, res, kw["sumq"], kw["double_fold"], kw["single_fold"],
                                       kw["use_raw"], kw["min_marginal_peaks"], kw["onlyanchor"])
    assert np.array_equal(gu.table_rows(final), z["table"])


@pytest.mark.parametrize("name", gu.names("apa"))
def test_apa_oracle_reproduces_reference(name):
    from oracle import apa_oracle as ao
    z, n, Diags, weights = gu.load_apa(name)
    w, cw = int(z["w"]), int(z["cw"])
    diags = ao.balanced_diags(Diags, weights)
    apa, valid = ao.apa_submatrix(diags, n, [tuple(p) for p in z["pos"]], w=w)
    assert len(apa) == int(z["n_windows"]) == int(valid.sum())
    assert np.array_equal(apa[0], z["win_first"]) and np.array_equal(apa[-1], z["win_last"])
    avg, score, zz, p, maxi, mean_arr, mask = ao.apa_analysis(apa, w=w, cw=cw)
    assert np.array_equal(mean_arr, z["mean_arr"])                      # bit-exact: the outlier cut selects on these
    assert np.array_equal(avg, z["avg"])
    assert np.array_equal(np.array([score, zz, p, maxi]), z["stats"])
    # the explicit pairwise model the CUDA kernel follows equals numpy's mean on these windows
    for a in (apa[0], apa[len(apa) // 2], apa[-1]):
        assert ao.pairwise_sum(a.ravel()) / a.size == a.mean()


@pytest.mark.parametrize("name", gu.names("bhfdr"))
def test_bhfdr_oracle_reproduces_reference(name):
    z, inp, kw, res = gu.load(name)
    sw, r = ho.score_bhfdr(inp, kw["pw"], kw["ww"], maxww=kw["maxww"], sig=kw["sig"], maxapart_bins=kw["maxapart"] // res)
    assert r["x"].size == int(z["n_tests"])
    assert gu.sha(np.stack([r["x"], r["y"]]).astype(np.int64)) == str(z["sha_xy"])
    for key in ("E", "O", "p", "q"):
        assert gu.sha(r[key]) == str(z["sha_" + key]), key + " bits"
    rej = r["reject"]
    assert int(rej.sum()) == int(z["n_reject"])
    assert np.array_equal(np.stack([r["x"][rej], r["y"][rej], r["O"][rej], r["E"][rej], r["p"][rej], r["q"][rej]]), z["surv"])
    final = glue_oracle.finish_bhfdr(inp, sw, r, kw["ww"], res, kw.get("min_marginal_peaks", 3), kw.get("onlyanchor", False))
    rows = np.array([list(k) + [float(v) for v in final[k]] for k in sorted(final)], dtype=np.float64).reshape(len(final), 9)
    assert np.array_equal(rows, z["table"])
