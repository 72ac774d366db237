"""GPU: the re-associated score kernel + exact re-evaluation (csrc/hp_score_fast.cuh, hp_exact.cuh) against the kernels
that add every pixel's cells in the reference's fp64 order, and against the oracle.  The fast path is only allowed to
differ in HOW it reaches a decision; every number that leaves the engine -- valid counts, E.max(), numbin, the (chunk, O)
histograms, q tables, survivor coordinates, their E, p, q and flags -- must be identical bit for bit."""
import numpy as np
import pytest

from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import _finish, synth_chromosome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _capi.Context(0)
    yield c
    c.close()


def _upload(ctx, inp, counts=False):
    Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
    if counts:
        ctx.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"])
    else:
        cD = [np.ascontiguousarray(c, dtype=np.float64) for c in inp["cDiags"]]
        ir = np.array([inp["IR"][d] for d in range(inp["min_ww"], inp["num"])], dtype=np.float64)
        ctx.upload(inp["n"], inp["num"], inp["min_ww"], Dg, cD, ir, inp["biases"], inp["biases"])


def _run(ctx, inp, p, w, maxww, sig, band, thr, exact, counts=False):
    """p, w: one pair or the lists of a union program."""
    _upload(ctx, inp, counts)
    pw, ww = (list(p), list(w)) if isinstance(p, (list, tuple)) else ([p], [w])
    P = ctx.make_params(pw, ww, maxww, sig, band, thr, exact_sums=exact)
    S1 = ctx.score(P)
    S = ctx.fdr()
    sv = ctx.survivors()
    sv = sv[np.lexsort((sv["pair"], sv["c"], sv["r"]))]
    tabs = [ctx.chunk_table(pi, fl) for pi in range(len(pw)) for fl in (0, 1)]
    return S1, S, sv, tabs


def _assert_same(a, b):
    S1a, Sa, sva, ta = a
    S1b, Sb, svb, tb = b
    assert (Sa.n_pixels, Sa.frozen_w, Sa.n_steps) == (Sb.n_pixels, Sb.frozen_w, Sb.n_steps)
    assert len(ta) == len(tb)
    for lf in range(len(ta)):
        La, Lb = Sa.lf[lf // 2][lf % 2], Sb.lf[lf // 2][lf % 2]
        assert (La.n_valid, La.e_max, La.numbin, La.n_reject) == (Lb.n_valid, Lb.e_max, Lb.numbin, Lb.n_reject), lf
    for x, y in zip(ta, tb):
        assert x[0] == y[0]
        assert np.array_equal(x[3], y[3]), "histograms differ"
        assert np.array_equal(x[5], y[5]), "q tables differ"
    assert Sa.n_survivors == Sb.n_survivors
    assert sva.tobytes() == svb.tobytes(), "survivor records differ"


CASES = [
    # n, band, p, w, maxww, scale, decay, thr, seed
    (600, 60, 2, 5, 8, 300.0, 1.08, 16, 1),
    (500, 60, 2, 5, 10, 40.0, 1.08, 16, 1),
    (129, 30, 2, 5, 7, 80.0, 1.1, 16, 5),              # ragged: n just above two tiles
    (1000, 200, 2, 5, 10, 300.0, 1.08, 16, 7),         # several tiles and strips
    (777, 150, 2, 5, 10, 12.0, 1.0, 16, 9),            # sparse: deep levels, frozen_w = 10 (the FM = 10 kernel)
    (700, 90, 1, 3, 10, 20.0, 1.0, 16, 3),
    (300, 40, 4, 7, 10, 60.0, 1.2, 30, 4),
    (2600, 300, 2, 5, 10, 300.0, 1.08, 16, 21),        # more than one work item per strip (row-tile carry-over)
    (90, 70, 2, 5, 10, 300.0, 1.08, 16, 2),            # chromosome shorter than band + windows: pixels next to both ends
    (640, 90, 3, 6, 10, 40.0, 1.05, 20, 12),           # a pair with no compiled exact-order kernel: table-driven k_score vs fast
    (520, 70, 0, 2, 9, 30.0, 1.1, 16, 13),             # p = 0
    (600, 80, 5, 7, 10, 50.0, 1.08, 16, 14),           # p > 4: one pair through the general-form kernel
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_b%d_p%dw%d" % (c[0], c[1], c[2], c[3]))
def test_fast_equals_exact_order(ctx, case):
    n, band, p, w, maxww, scale, decay, thr, seed = case
    inp = synth_chromosome(n, band, w, maxww=maxww, seed=seed, scale=scale, decay=decay)
    fast = _run(ctx, inp, p, w, maxww, 0.1, band, thr, exact=False)
    exact = _run(ctx, inp, p, w, maxww, 0.1, band, thr, exact=True)
    assert exact[0].fast_kernel == 0
    assert fast[0].fast_kernel == 1, "the re-associated kernel did not run"
    assert 0 < fast[0].n_exact
    _assert_same(fast, exact)
    # worker-level input (band built on the GPU) through the fast path too
    _assert_same(_run(ctx, inp, p, w, maxww, 0.1, band, thr, exact=False, counts=True), exact)


UNION_CASES = [
    # n, band, pw, ww, maxww, scale, decay, thr, seed
    (900, 120, (1, 2, 4), (3, 5, 7), 10, 300.0, 1.08, 16, 31),     # the BASELINE cfg3 program (compiled exact-order kernel)
    (700, 100, (1, 2, 4), (3, 5, 7), 10, 15.0, 1.0, 16, 32),       # sparse: deep levels, re-added rings at every width
    (640, 90, (1, 2), (3, 5), 10, 40.0, 1.05, 16, 33),             # two pairs: table-driven k_score on the exact side
    (500, 80, (2, 4), (5, 7), 9, 60.0, 1.1, 20, 34),
    (129, 40, (1, 3), (4, 6), 8, 80.0, 1.1, 16, 35),               # ragged, FM = 8 kernel
    (90, 70, (1, 2, 4), (3, 5, 7), 10, 300.0, 1.08, 16, 36),       # pixels next to both chromosome ends
    (2300, 260, (4, 2, 1), (7, 5, 3), 10, 200.0, 1.08, 16, 37),    # pairs given in another order, several tiles per strip
]


@pytest.mark.parametrize("case", UNION_CASES, ids=lambda c: "n%d_b%d_p%s_w%s" % (c[0], c[1], "".join(map(str, c[2])), "".join(map(str, c[3]))))
def test_union_fast_equals_exact_order(ctx, case):
    """Union programs: one launch of the general-form fast kernel per pair (ring multiplicities as coefficients of the
    quadrant boxes) against the exact-order kernels -- every pair's counts, E.max(), histograms, q tables and survivors."""
    n, band, pw, ww, maxww, scale, decay, thr, seed = case
    inp = synth_chromosome(n, band, min(ww), maxww=maxww, seed=seed, scale=scale, decay=decay)
    try:
        exact = _run(ctx, inp, pw, ww, maxww, 0.1, band, thr, exact=True)
    except _capi.EngineError as e:                       # the reference's crash (empty unresolved set): same on both routes
        assert e.code == _capi.HP_ERR_EMPTY_REFIDX
        with pytest.raises(_capi.EngineError):
            _run(ctx, inp, pw, ww, maxww, 0.1, band, thr, exact=False)
        return
    fast = _run(ctx, inp, pw, ww, maxww, 0.1, band, thr, exact=False)
    assert exact[0].fast_kernel == 0
    assert fast[0].fast_kernel == 1, "the re-associated kernel did not run"
    _assert_same(fast, exact)
    _assert_same(_run(ctx, inp, pw, ww, maxww, 0.1, band, thr, exact=False, counts=True), exact)


@pytest.mark.parametrize("case", CASES[:5], ids=lambda c: "n%d_b%d_p%dw%d" % (c[0], c[1], c[2], c[3]))
def test_fast_path_matches_oracle(ctx, case):
    """The same cut points compare_with_oracle checks, minus the per-pixel planes (the fast path keeps none)."""
    from helpers import compare_survivor_path_with_oracle
    n, band, p, w, maxww, scale, decay, thr, seed = case
    inp = synth_chromosome(n, band, w, maxww=maxww, seed=seed, scale=scale, decay=decay)
    st = compare_survivor_path_with_oracle(ctx, inp, [p], [w], maxww, 0.1, band, thr, expect_fast=True)
    assert st["n_pixels"] > 0


def _flat_chromosome(n, band, count, eps, seed=0):
    """Every count equal, weights 1 +- eps: every interior pixel has bS / bE ~ 1 and E ~ count.  With count a power of two
    E sits on (eps = 0: exactly on) a lambda-chunk edge 2^k -- strict membership (callers.py:38) decides every pixel."""
    rng = np.random.default_rng(seed)
    num = band + 10 + 1
    Diags = [np.full(n - d, count, dtype=np.int32) for d in range(num)]
    w = 1.0 + eps * rng.standard_normal(n)
    return _finish(n, num, 5, Diags, w)


@pytest.mark.parametrize("eps", [0.0, 1e-15, 1e-9, 1e-7, 3e-6, 1e-4])
def test_expected_values_on_a_chunk_edge(ctx, eps):
    inp = _flat_chromosome(700, 80, 64, eps)             # E ~ 64 = 2^6 = the upper edge of chunk 19
    fast = _run(ctx, inp, 2, 5, 10, 0.1, 80, 16, exact=False)
    exact = _run(ctx, inp, 2, 5, 10, 0.1, 80, 16, exact=True)
    assert fast[0].fast_kernel == 1
    _assert_same(fast, exact)
    if eps <= 1e-7:                                      # far inside the guard band: (almost) every pixel is settled exactly
        assert fast[0].n_exact > 0.9 * fast[1].lf[0][0].n_valid


def test_inputs_outside_the_bound_fall_back(ctx):
    """A negative balanced value (a negative weight) is outside the domain of the error bound: the fast kernel flags it
    and the chromosome is scored by the exact-order kernel -- same numbers as asking for it."""
    inp = synth_chromosome(500, 60, 5, maxww=10, seed=3, scale=40.0)
    w = inp["weights"].copy()
    w[100] = -w[100]
    inp2 = _finish(inp["n"], inp["num"], 5, inp["Diags"], w)
    a = _run(ctx, inp2, 2, 5, 10, 0.1, 60, 16, exact=False)
    b = _run(ctx, inp2, 2, 5, 10, 0.1, 60, 16, exact=True)
    assert a[0].fast_kernel == 0
    _assert_same(a, b)


def test_cfg2_fast_equals_exact_order():
    """BASELINE configs[1]: 20 000-bin chromosome, 5 Mb band, (2, 5)."""
    inp = synth_chromosome(20000, 500, 5, maxww=10, seed=17)
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        fast = _run(c1, inp, 2, 5, 10, 0.1, 500, 16, exact=False, counts=True)
        exact = _run(c2, inp, 2, 5, 10, 0.1, 500, 16, exact=True, counts=True)
        assert fast[0].fast_kernel == 1 and exact[0].fast_kernel == 0
        _assert_same(fast, exact)
        assert fast[0].n_exact < 0.02 * fast[1].n_pixels, "too many records left to the exact kernel"
        print("cfg2: n_exact %d of %d pixels, ms_score fast %.3f exact %.3f, ms_exact %.3f" % (
            fast[0].n_exact, fast[1].n_pixels, fast[0].ms_score, exact[0].ms_score, fast[0].ms_exact))


def test_cfg4_shape_fast_equals_exact_order():
    n, band = 9342, 2000
    inp = synth_chromosome(n, band, 7, maxww=10, seed=3)
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        fast = _run(c1, inp, 4, 7, 10, 0.1, band, 16, exact=False)
        exact = _run(c2, inp, 4, 7, 10, 0.1, band, 16, exact=True)
        assert fast[0].fast_kernel == 1
        _assert_same(fast, exact)
        print("cfg4 shape: n_exact %d of %d pixels, ms_score fast %.3f exact %.3f, ms_exact %.3f" % (
            fast[0].n_exact, fast[1].n_pixels, fast[0].ms_score, exact[0].ms_score, fast[0].ms_exact))
