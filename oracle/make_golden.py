"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported by path; see oracle/ref_harness.py) in the build container.

    python oracle/make_golden.py            # rewrites every fixture

Each fixture stores the inputs (raw band, weights, parameters) and the reference's outputs at four
cut points: sha256 of the per-pixel arrays (pixel set, bSV, bEV, E, p, q, chunk -- a bit-exact pin
for the oracle), the FDR survivors with their values (checked against the CUDA path with the q
tolerance), and the final peak table.  The GPU box has no /root/reference: tests only read the npz.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hicpeaks_b200.synth import band_from_dense, synth_chromosome  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def simple_ice(counts, iters=200, low_frac=0.1):
    """Plain iterative correction (not cooler's; any deterministic weight vector is a valid input)."""
    m = counts.astype(np.float64)
    marg = m.sum(axis=1)
    ok = marg > 0
    ok &= marg > low_frac * np.median(marg[ok])
    m = m * ok[:, None] * ok[None, :]
    bias = np.ones(m.shape[0])
    for _ in range(iters):
        s = m.sum(axis=1)
        s = s / s[ok].mean()
        s[~ok] = 1
        bias *= s
        m = m / s[:, None] / s[None, :]
        if s[ok].var() < 1e-10:
            break
    w = 1.0 / bias
    w[~ok] = np.nan
    scale = np.nanmean((counts * np.nan_to_num(w)[:, None] * np.nan_to_num(w)[None, :]).sum(axis=1)[ok])
    return w / np.sqrt(scale)


def chr21_example(num, min_ww):
    n = 1869                                     # ceil(46709983 / 25000), example/hg38.chromsizes
    dat = np.loadtxt("/root/reference/example/25K/21_21.txt", dtype=np.int64)
    counts = np.zeros((n, n), dtype=np.int64)
    counts[dat[:, 0], dat[:, 1]] = dat[:, 2]
    counts = np.maximum(counts, counts.T)
    w = simple_ice(counts)
    return band_from_dense(counts, w, num, min_ww)


def table_rows(table):
    rows = [list(k) + [float(v) for v in table[k]] for k in sorted(table)]
    return np.array(rows, dtype=np.float64).reshape(len(rows), 12)


def pack_inputs(inp):
    n, num = inp["n"], inp["num"]
    band = np.zeros((num, n), dtype=np.int32)
    for d in range(num):
        band[d, : n - d] = inp["Diags"][d]
    return dict(in_n=n, in_num=num, in_min_ww=inp["min_ww"], in_band=band, in_weights=inp["weights"])


def make_hiccups(name, inp, res, kw):
    out = pack_inputs(inp)
    out["kind"] = "hiccups"
    out["res"] = res
    for k, v in kw.items():
        out["kw_" + k] = np.array(v)
    try:
        table, cap = ref_harness.run_hiccups(inp, res, **kw)
    except Exception as e:                       # the reference's own failure mode is part of the contract
        out["raises"] = type(e).__name__
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "reference raises", type(e).__name__)
        return
    out["raises"] = ""
    vx, vy = cap["pixels"]
    out["n_pixels"] = vx.size
    out["sha_pixels"] = sha(np.stack([vx, vy]).astype(np.int64))
    for p in kw["pw"]:
        for fl, nm in ((0, "K"), (1, "Y")):
            c = cap[(p, nm)]
            pre = "p%d%s_" % (p, nm)
            out[pre + "n_valid"] = c["x"].size
            out[pre + "numbin"] = c["numbin"]
            for key in ("bSV", "bEV", "E", "p", "q"):
                out[pre + "sha_" + key] = sha(c[key].astype(np.float64))
            out[pre + "sha_xy"] = sha(np.stack([c["x"], c["y"]]).astype(np.int64))
            out[pre + "sha_chunk"] = sha(c["chunk"].astype(np.int64))
            rej = c["q"] <= float(kw["sig"])
            out[pre + "surv"] = np.stack([c["x"][rej], c["y"][rej], c["O"][rej], c["E"][rej], c["p"][rej], c["q"][rej]]).astype(np.float64)
    out["table"] = table_rows(table)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "pixels", vx.size, "peaks", len(table), "survivors",
          {k[:-5]: out[k].shape[1] for k in out if k.endswith("_surv")})


def table_rows_bh(table):
    rows = [list(k) + [float(v) for v in table[k]] for k in sorted(table)]
    return np.array(rows, dtype=np.float64).reshape(len(rows), 9)


def make_bhfdr(name, inp, res, kw):
    out = pack_inputs(inp)
    out["kind"] = "bhfdr"
    out["res"] = res
    for k, v in kw.items():
        out["kw_" + k] = np.array(v)
    table, c = ref_harness.run_bhfdr(inp, res, **kw)
    out["n_tests"] = c["x"].size
    out["sha_xy"] = sha(np.stack([c["x"], c["y"]]).astype(np.int64))
    for key in ("E", "O", "p", "q"):
        out["sha_" + key] = sha(c[key].astype(np.float64))
    rej = c["reject"]
    out["n_reject"] = int(rej.sum())
    out["surv"] = np.stack([c["x"][rej], c["y"][rej], c["O"][rej], c["E"][rej], c["p"][rej], c["q"][rej]]).astype(np.float64)
    out["table"] = table_rows_bh(table)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "tests", c["x"].size, "rejected", int(rej.sum()), "peaks", len(table))


def main():
    os.makedirs(GOLD, exist_ok=True)
    cli = dict(maxww=10, sig=0.05, sumq=0.01, double_fold=1.75, single_fold=2, use_raw=False,
               min_marginal_peaks=2, min_local_reads=16)
    # config 1: the example that ships with the reference (README.rst:200-204)
    inp = chr21_example(2000000 // 25000 + 10 + 1, 3)
    make_hiccups("chr21_25k_p1w3", inp, 25000, dict(cli, pw=[1], ww=[3], maxapart=2000000, onlyanchor=False))
    make_hiccups("chr21_25k_union", inp, 25000, dict(cli, pw=[1, 2, 4], ww=[3, 5, 7], maxapart=2000000, onlyanchor=True))
    # synthetic: adaptive width, union multiplicities, (4,7), use_raw, the reference's crash
    inp = synth_chromosome(500, 60, 5, maxww=10, seed=1, scale=40.0)
    make_hiccups("synth_p2w5", inp, 10000, dict(cli, pw=[2], ww=[5], maxapart=600000, onlyanchor=False, sig=0.1))
    inp = synth_chromosome(400, 80, 3, maxww=8, seed=1, scale=100.0, decay=1.3)
    make_hiccups("synth_union_12", inp, 10000, dict(cli, pw=[1, 2], ww=[3, 5], maxww=8, maxapart=800000,
                                                    onlyanchor=False, sig=0.1, min_local_reads=25, use_raw=True))
    inp = synth_chromosome(500, 80, 3, maxww=10, seed=1, scale=40.0)
    make_hiccups("synth_union_124", inp, 10000, dict(cli, pw=[1, 2, 4], ww=[3, 5, 7], maxapart=800000, onlyanchor=True, sig=0.1))
    inp = synth_chromosome(300, 40, 7, maxww=12, seed=4, scale=60.0, decay=1.2)
    make_hiccups("synth_p4w7", inp, 5000, dict(cli, pw=[4], ww=[7], maxww=12, maxapart=200000, onlyanchor=False,
                                               sig=0.1, min_local_reads=30, min_marginal_peaks=3))
    inp = synth_chromosome(300, 60, 3, maxww=8, seed=2, scale=300.0)
    make_hiccups("synth_union_crash", inp, 10000, dict(cli, pw=[1, 2], ww=[3, 5], maxww=8, maxapart=600000, onlyanchor=False))


def main_bhfdr():
    inp = chr21_example(2000000 // 25000 + 10 + 1, 3)
    make_bhfdr("bh_chr21_25k", inp, 25000, dict(pw=1, ww=3, sig=0.05, maxww=10, maxapart=2000000))
    inp = synth_chromosome(500, 60, 5, maxww=10, seed=1, scale=40.0)
    make_bhfdr("bh_synth_p2w5", inp, 10000, dict(pw=2, ww=5, sig=0.05, maxww=10, maxapart=600000))
    inp = synth_chromosome(400, 70, 5, maxww=20, seed=3, scale=25.0, decay=1.0)
    make_bhfdr("bh_synth_deep", inp, 10000, dict(pw=2, ww=5, sig=0.1, maxww=20, maxapart=700000, min_marginal_peaks=2,
                                                 onlyanchor=True))


if __name__ == "__main__":
    if "--bhfdr-only" not in sys.argv:
        main()
    main_bhfdr()
