"""GPU, BASELINE.json sizes.  Two kinds of checks:

* against the ORACLE on one chromosome of every BASELINE shape (the numpy port needs 15 - 40 s for each): cfg2 (20 000
  bins, 5 Mb band, (2,5)) at every cut point including the per-pixel fp64 sums, the same chromosome through the default
  (re-associated + exact re-evaluation) path, a cfg3-shaped union chromosome, a cfg4-shaped (num = 2011, (4,7)) one, and
  the cfg5 APA pile-up (50 000 anchors, w = 20) against apa_oracle;
* size-independent properties: the specialised and the table-driven kernels (different tiling, different summation
  machinery; they share the per-pixel tail) agree bit for bit, the worker-level upload equals the operator-level one,
  conservation laws of the histograms hold, and a second run reproduces the first."""
import numpy as np
import pytest

from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import band_pixels, synth_chromosome

pytestmark = pytest.mark.gpu

N, BAND = 20000, 500


@pytest.fixture(scope="module")
def big():
    return synth_chromosome(N, BAND, 5, maxww=10, seed=17)


def _run(ctx, inp, pw, ww, generic=False, counts=False, sig=0.1, band=None):
    Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
    if counts:
        ctx.upload_counts(inp["n"], inp["num"], min(ww), Dg, inp["weights"])
    else:
        cD = [np.ascontiguousarray(c, dtype=np.float64) for c in inp["cDiags"]]
        ir = np.array([inp["IR"][d] for d in range(inp["min_ww"], inp["num"])])
        ctx.upload(inp["n"], inp["num"], inp["min_ww"], Dg, cD, ir, inp["biases"], inp["biases"])
    P = ctx.make_params(pw, ww, 10, sig, band or BAND, 16, generic_kernel=generic)
    S = ctx.hiccups(P)
    sv = ctx.survivors()
    sv = sv[np.lexsort((sv["pair"], sv["c"], sv["r"]))]
    tabs = [ctx.chunk_table(pi, fl) for pi in range(len(pw)) for fl in (0, 1)]
    return S, sv, tabs


def _same(a, b):
    Sa, sva, ta = a
    Sb, svb, tb = b
    assert (Sa.n_pixels, Sa.frozen_w, Sa.n_steps, Sa.n_survivors) == (Sb.n_pixels, Sb.frozen_w, Sb.n_steps, Sb.n_survivors)
    assert sva.tobytes() == svb.tobytes()                       # coordinates, O, ice, E, p, q: every bit
    for x, y in zip(ta, tb):
        assert x[0] == y[0] and np.array_equal(x[3], y[3]) and np.array_equal(x[5], y[5])       # numbin, hist, q tables


def test_cfg2_kernels_agree_and_conserve(big):
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        spec = _run(c1, big, [2], [5])
        gen = _run(c2, big, [2], [5], generic=True)
        assert spec[0].spec_kernel == 1 and gen[0].spec_kernel == 0
        assert spec[0].band_pixels == band_pixels(N, 5, BAND) == 9794760
        _same(spec, gen)
        S = spec[0]
        assert sum(S.steps[k].resolved for k in range(S.n_steps)) <= S.n_pixels
        for fl, (nb, widths, off, hist, p, q) in enumerate(spec[2]):
            assert 0 < hist.sum() <= S.lf[0][fl].n_valid <= S.n_pixels       # every valid pixel is in at most one chunk
            assert np.all((q >= 0) & (q <= 1)) and np.all((p >= 0) & (p <= 1))
        _same(spec, _run(c1, big, [2], [5]))                              # idempotent
        _same(spec, _run(c2, big, [2], [5], counts=True))                 # worker-level input == operator-level input


def test_union_program_kernels_agree_at_scale():
    inp = synth_chromosome(6000, 500, 3, maxww=10, seed=23)
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        a = _run(c1, inp, [1, 2, 4], [3, 5, 7])
        b = _run(c2, inp, [1, 2, 4], [3, 5, 7], generic=True)
        assert a[0].spec_kernel == 1 and b[0].spec_kernel == 0
        _same(a, b)


def test_cfg4_shape_wide_band_p4w7():
    """BASELINE configs[3] shape: 5 kb bins, 10 Mb band (num = 2011 stored diagonals), (p, w) = (4, 7); a chr21-sized
    chromosome.  Both kernels, both input boundaries."""
    n, band = 9342, 2000
    inp = synth_chromosome(n, band, 7, maxww=10, seed=3)
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        a = _run(c1, inp, [4], [7], band=band)
        b = _run(c2, inp, [4], [7], generic=True, band=band)
        assert a[0].spec_kernel == 1 and b[0].spec_kernel == 0
        assert a[0].band_pixels == band_pixels(n, 7, band)
        _same(a, b)
        _same(a, _run(c2, inp, [4], [7], counts=True, band=band))


def test_cfg3_shape_chr1_union_worker_input():
    """BASELINE configs[2] shape: hg38 chr1 @10 kb (24 896 bins), 5 Mb band, union (1,3)/(2,5)/(4,7), through the
    worker-level upload (narrowed counts, band / IR / biases derived on the GPU) against the operator-level one."""
    n = 24896
    inp = synth_chromosome(n, BAND, 3, maxww=10, seed=29)
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        a = _run(c1, inp, [1, 2, 4], [3, 5, 7], counts=True)
        b = _run(c2, inp, [1, 2, 4], [3, 5, 7])
        assert a[0].spec_kernel == 1 and a[0].band_pixels == band_pixels(n, 3, BAND)
        _same(a, b)


# ---- one chromosome of every BASELINE shape against the oracle ------------------------------------------------------
def test_cfg2_chromosome_matches_oracle_at_every_cut_point(big):
    """configs[1] at full size: levels, per-pixel bS / bE / E (bit-exact), histograms, p / q, survivors (exact-order kernel,
    the dump planes need it), then the same chromosome through the default path."""
    from helpers import compare_survivor_path_with_oracle, compare_with_oracle
    with _capi.Context(0) as ctx:
        st = compare_with_oracle(ctx, big, [2], [5], 10, 0.1, BAND, 16)
        assert st["spec_kernel"] == 1 and st["n_pixels"] > 5_000_000
        st2 = compare_survivor_path_with_oracle(ctx, big, [2], [5], 10, 0.1, BAND, 16, counts=True, expect_fast=True)
        assert st2["n_survivors"] == st["n_survivors"]


def test_cfg3_shape_union_chromosome_matches_oracle():
    from helpers import compare_with_oracle
    inp = synth_chromosome(8200, BAND, 3, maxww=10, seed=31)
    with _capi.Context(0) as ctx:
        st = compare_with_oracle(ctx, inp, [1, 2, 4], [3, 5, 7], 10, 0.1, BAND, 16)
        assert st["spec_kernel"] == 1 and st["n_pixels"] > 2_000_000


def test_cfg4_shape_chromosome_matches_oracle():
    from helpers import compare_survivor_path_with_oracle, compare_with_oracle
    n, band = 4100, 2000
    inp = synth_chromosome(n, band, 7, maxww=10, seed=5)
    with _capi.Context(0) as ctx:
        st = compare_with_oracle(ctx, inp, [4], [7], 10, 0.1, band, 16)
        assert st["spec_kernel"] == 1
        compare_survivor_path_with_oracle(ctx, inp, [4], [7], 10, 0.1, band, 16, expect_fast=True)


def test_cfg5_apa_50000_anchors_matches_oracle():
    """configs[4]: 41 x 41 pile-up over 50 000 anchors @10 kb against the numpy restatement of apa.py:11-46."""
    import time
    from hicpeaks_b200 import apa as hapa
    from oracle import apa_oracle as ao
    from test_gpu_apa import BandMatrix
    n, band, w, cw = 20000, 500, 20, 3
    inp = synth_chromosome(n, band + 2 * w, 5, maxww=10, seed=41)
    rng = np.random.default_rng(7)
    i = rng.integers(w, n - band - w - 1, 50000)
    d = rng.integers(10 + w, band - w, 50000)
    pos = [(int(a), int(a + b)) for a, b in zip(i, d)]
    diags = ao.balanced_diags(inp["Diags"], inp["weights"])
    t0 = time.perf_counter()
    exp, valid = ao.apa_submatrix(diags, n, pos, w=w)
    ref = ao.apa_analysis(np.asarray(exp), w=w, cw=cw)
    t_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = hapa.apa_submatrix(BandMatrix(diags, n), pos, w=w)
    res = hapa.apa_analysis(got, w=w, cw=cw)
    t_gpu = time.perf_counter() - t0
    assert len(got) == len(exp) > 10000
    assert np.array_equal(got.mean_arr, np.array([x.mean() for x in exp]))          # numpy's pairwise mean, bit for bit
    assert np.array_equal(res[0], ref[0])                                           # the 41 x 41 pile-up
    assert tuple(float(x) for x in res[1:5]) == tuple(float(x) for x in ref[1:5])   # score, z, p, maxi
    print("cfg5 APA: %d of %d windows kept, numpy port %.2f s, engine %.3f s (upload + kernels + download)" % (
        len(got), len(pos), t_cpu, t_gpu))
