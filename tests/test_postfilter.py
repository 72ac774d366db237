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


def test_host_helpers_of_the_mirror_match_the_restated_reference():
    """pw_ww_pairs (callers.py:15-23), lambdachunk (:25-41) and the truncated Benjamini-Hochberg used by the BH-FDR
    caller (:545-547 on the pixels the GPU returns) against the oracle's restatements."""
    from hicpeaks_b200 import callers
    from oracle import hiccups_oracle as ho
    rng = np.random.default_rng(3)
    for pw, ww, maxww in [([2], [5], 10), ([1, 2, 4], [3, 5, 7], 10), ([4, 1], [7, 3], 12), ([2], [5], 20)]:
        assert callers.pw_ww_pairs(pw, ww, maxww) == ho.pw_ww_steps(pw, ww, maxww)
    E = np.exp(rng.uniform(np.log(0.05), np.log(3000.0), 5000))
    E[:4] = [1.0, 2.0, 2 ** (1 / 3.), 4.0]                     # values sitting on chunk edges belong to no chunk
    chunks = callers.lambdachunk(E)
    edges = ho.chunk_edges(len(chunks))
    assert len(chunks) == int(np.ceil(np.log(E.max()) / np.log(2) * 3 + 1))
    seen = np.zeros(E.size, dtype=int)
    for i, (lv, rv, idx) in enumerate(chunks):
        assert (lv, rv) == tuple(edges[i])
        assert np.all((E[idx] > lv) & (E[idx] < rv))
        seen[idx] += 1
    assert np.all(seen[4:] == 1) and np.all(seen[:4] == 0)
    assert callers.lambdachunk(np.array([])) == []
    # BH over n tests of which only the p-values <= alpha are known (what the GPU hands back, callers.py:536-547): every
    # rejected test is among them, and a larger p-value can never lower the q of one that is rejected
    for n, alpha, power in [(5000, 0.05, 3), (300, 0.1, 6), (1000, 0.05, 1), (50, 0.01, 8), (200, 1.0, 1)]:
        p = rng.uniform(0, 1, n) ** power
        q_full = ho.bh_fdr(p)
        known = np.nonzero(p <= alpha)[0]
        known = known[rng.permutation(known.size)]
        if known.size == 0:
            continue
        reject, q = callers.bh_adjust(p[known], n, alpha)
        full_rej = q_full[known] <= alpha
        assert np.array_equal(reject, full_rej)
        assert np.array_equal(q[reject], q_full[known][reject])
        assert np.all(q >= q_full[known])
