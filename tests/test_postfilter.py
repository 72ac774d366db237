"""CPU: the product's host tail (postfilter.py / callers.assemble_table) fed with the oracle's
survivors must reproduce the reference's final peak table."""
import numpy as np
import pytest

import golden_util as gu
from hicpeaks_b200 import _capi, callers
from oracle import hiccups_oracle as ho


def survivors_from_oracle(out, pw):
    recs = {}
    for pi, p in enumerate(pw):
        k, y = out[(p, 0)], out[(p, 1)]
        nzy = set(zip(y["cem_nz"][0].tolist(), y["cem_nz"][1].tolist()))
        for fl, r in ((0, k), (1, y)):
            for i in np.where(r["reject"])[0]:
                key = (pi, int(r["x"][i]), int(r["y"][i]))
                rec = recs.setdefault(key, dict(r=key[1], c=key[2], pair=pi, flags=0, obs=r["O"][i], ice=r["ice"][i],
                                                e=[0.0, 0.0], p=[1.0, 1.0], q=[1.0, 1.0]))
                rec["flags"] |= (_capi.SF_REJECT_K | _capi.SF_VALID_K) if fl == 0 else (_capi.SF_REJECT_Y | _capi.SF_VALID_Y)
                rec["e"][fl], rec["p"][fl], rec["q"][fl] = r["E"][i], r["p"][i], r["q"][i]
        for key, rec in recs.items():
            if key[0] == pi and (key[1], key[2]) in nzy:
                rec["flags"] |= _capi.SF_CEMY_NONZERO
    sv = np.zeros(len(recs), dtype=_capi.SURVIVOR_DTYPE)
    for j, rec in enumerate(recs.values()):
        for f in ("r", "c", "pair", "flags", "obs", "ice"):
            sv[f][j] = rec[f]
        sv["e"][j], sv["p"][j], sv["q"][j] = rec["e"], rec["p"], rec["q"]
    return sv


@pytest.mark.parametrize("name", [n for n in gu.names("hiccups") if "crash" not in n])
def test_host_tail_reproduces_reference_table(name):
    z, inp, kw, res = gu.load(name)
    pw, ww = kw["pw"], kw["ww"]
    sw, out = ho.score(inp, pw, ww, maxww=kw["maxww"], sig=kw["sig"], maxapart_bins=kw["maxapart"] // res,
                       min_local_reads=kw["min_local_reads"])
    sv = survivors_from_oracle(out, pw)
    gaps = sw["bal"].sum(axis=0) == 0
    table = callers.assemble_table(sv, gaps, inp["n"], pw, ww, res, kw["sumq"], kw["double_fold"], kw["single_fold"],
                                   kw["use_raw"], kw["min_marginal_peaks"], kw["onlyanchor"])
    assert np.array_equal(gu.table_rows(table), z["table"])


def test_gap_filter_matches_loop_form():
    rng = np.random.default_rng(0)
    n = 200
    gaps = rng.random(n) < 0.1
    x = rng.integers(0, n, 500)
    y = np.minimum(x + rng.integers(1, 30, 500), n - 1)
    from hicpeaks_b200.postfilter import gap_filter
    from oracle.glue_oracle import gap_keep
    for m in (1, 3, 7):
        keep = gap_filter(x, y, gaps, m, n)
        ref = gap_keep(x, y, set(np.where(gaps)[0]), m, n)
        assert np.array_equal(np.where(keep)[0], np.array(ref, dtype=np.int64))
