"""GPU parity: the CUDA path (through the C ABI) against the oracle on seeded synthetic inputs."""
import numpy as np
import pytest

from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import synth_chromosome

from helpers import compare_with_oracle

pytestmark = pytest.mark.gpu

# sweep programs compiled into the library (hp_api.cu g_specs): these must take the specialised kernel
COMPILED = {((2,), (5,)), ((1,), (3,)), ((4,), (7,)), ((1, 2, 4), (3, 5, 7))}

CASES = [
    # n, band, pw, ww, maxww, scale, decay, thr, seed
    (600, 60, [2], [5], 8, 300.0, 1.08, 16, 1),          # dense: freezes at the first level
    (500, 60, [2], [5], 10, 40.0, 1.08, 16, 1),          # two levels
    (500, 80, [1, 2, 4], [3, 5, 7], 10, 40.0, 1.08, 16, 1),   # union mode, re-added rings
    (400, 80, [1, 2], [3, 5], 8, 100.0, 1.3, 25, 1),     # union, deeper levels
    (700, 90, [1], [3], 10, 20.0, 1.0, 16, 3),           # (1,3), many levels
    (300, 40, [4], [7], 12, 60.0, 1.2, 30, 4),           # (4,7)
    (129, 30, [2], [5], 7, 80.0, 1.1, 16, 5),            # ragged: n just above one tile
    (1000, 200, [2], [5], 10, 300.0, 1.08, 16, 7),       # several tiles in both directions
    (777, 150, [2], [5], 10, 12.0, 1.0, 16, 9),          # sparse: deep levels, many zero pixels
    (515, 100, [1, 2, 4], [3, 5, 7], 10, 15.0, 1.0, 16, 10),  # union, sparse
    (800, 120, [2], [5], 20, 300.0, 1.08, 16, 11),       # the operator's default maxww = 20 (callers.py:45): freezes early,
                                                         # the executed prefix is a compiled-in program
]


@pytest.fixture(scope="module")
def ctx():
    c = _capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("kernel", ["spec", "generic"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_b%d_%s" % (c[0], c[1], "-".join(map(str, c[2]))))
def test_cutpoints_match_oracle(ctx, case, kernel):
    n, band, pw, ww, maxww, scale, decay, thr, seed = case
    inp = synth_chromosome(n, band, min(ww), maxww=maxww, seed=seed, scale=scale, decay=decay)
    st = compare_with_oracle(ctx, inp, pw, ww, maxww, 0.1, band, thr, generic_kernel=(kernel == "generic"))
    print(case, st)
    assert st["n_pixels"] > 0
    if kernel == "generic":
        assert st["spec_kernel"] == 0
    elif (tuple(pw), tuple(ww)) in COMPILED and st["frozen"] <= 10:
        assert st["spec_kernel"] == 1, "a compiled-in sweep program fell back to the generic kernel"


def test_poisson_tail_matches_scipy(ctx):
    from scipy.special import pdtr
    rng = np.random.default_rng(0)
    mu = np.exp(rng.uniform(np.log(1e-3), np.log(5e4), 100000))
    k = np.floor(np.maximum(0, mu + rng.normal(0, 1, mu.size) * np.sqrt(mu) * 4 + rng.uniform(0, 3, mu.size)))
    got = ctx.poisson_sf(k, mu)
    ref = 1 - pdtr(k, mu)
    assert np.abs(got - ref).max() < 2e-14


def test_narrowed_count_upload_restores_every_count():
    """hp_band_upload_counts sends each count diagonal as u8 / u16 / i32, whichever holds it exactly; the band the
    device rebuilds (balanced = (w[r] * w[c]) * count, scripts/pyHICCUPS:143) must not depend on the format taken."""
    from hicpeaks_b200 import _capi
    n, num, mw = 1237, 131, 3
    rng = np.random.default_rng(5)
    w = np.exp(rng.normal(0, 0.2, n))
    w[rng.choice(n, 9, replace=False)] = np.nan
    Dg = [rng.poisson(40.0 / (d + 1), n - d).astype(np.int32) for d in range(num)]
    Dg[0][:] += 300                       # every value above a byte
    Dg[4][n - 5] = 255                    # the last value a byte holds, in the final (partial) quad
    Dg[5][17] = 256                       # one value over
    Dg[9][n - 10] = 65535
    Dg[11][0] = 65536
    Dg[12][3] = 2 ** 31 - 1
    Dg[13][:] = 0
    with _capi.Context(0) as ctx:
        ctx.upload_counts(n, num, mw, Dg, w)
        bal = ctx.dump_band(2)
        sent = ctx.upload_bytes()
    differs = 0
    for d in range(mw, num):
        exp = w[: n - d] * w[d:] * Dg[d].astype(np.float64)          # cooler: (bias1[row] * bias2[col]) * data
        exp[Dg[d] == 0] = 0.0
        exp[np.isnan(exp)] = 0.0
        assert np.array_equal(bal[d, : n - d], exp), d
        with np.errstate(invalid="ignore"):
            other = np.nan_to_num(Dg[d].astype(np.float64) * w[: n - d] * w[d:])
        differs += int((other != exp).sum())
    # fp64 multiplication is not associative: (count * w[r]) * w[c] is a different number for many pixels, so this test
    # does pin the order cooler uses (api.py, sparse branch of matrix(): bias1[row] * bias2[col] * data, left to right)
    assert differs > 100
    assert sent < 2 * sum(a.size for a in Dg)          # most diagonals travelled as bytes


def test_chromosomes_in_flight_match_one_at_a_time():
    """Several host threads, each with its own context and stream, upload (shared pack pool) and score different
    chromosomes at once -- the way the dispatcher and bench.py drive a GPU -- and get what a lone call gets."""
    from concurrent.futures import ThreadPoolExecutor
    from hicpeaks_b200 import _capi
    inputs = [synth_chromosome(2500 + 37 * i, 300, 5, maxww=10, seed=40 + i) for i in range(6)]
    P = _capi.Context.make_params([2], [5], 10, 0.1, 300, 16)

    def one(ctx, inp):
        Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
        ctx.upload_counts(inp["n"], inp["num"], 5, Dg, inp["weights"])
        S = ctx.hiccups(P)
        sv = ctx.survivors()
        sv = sv[np.lexsort((sv["c"], sv["r"]))]
        return (S.n_pixels, S.frozen_w, S.n_survivors), sv.tobytes(), ctx.gaps().tobytes()

    ctxs = [_capi.Context(0) for _ in inputs]
    try:
        alone = [one(c, i) for c, i in zip(ctxs, inputs)]
        with ThreadPoolExecutor(len(inputs)) as pool:
            for _ in range(3):
                together = list(pool.map(lambda a: one(*a), zip(ctxs, inputs)))
                assert together == alone
        assert alone[0][0][2] > 0
    finally:
        for c in ctxs:
            c.close()


def _staircase_chromosome(n=640, band=100, maxww=20, thr=16, seed=3):
    """Counts whose density falls in steps with the distance, tuned so that every sweep step resolves >= 30 % of what is
    left (callers.py:219-232) until w = 11: the adaptive width runs past the widths compiled into the library."""
    from hicpeaks_b200.synth import _finish
    num = band + maxww + 1
    rng = np.random.default_rng(seed)
    fr = np.array([0.30, 0.50, 0.64, 0.74, 0.81, 0.86, 0.90, 0.93, 0.95, 0.965, 0.975, 0.985, 1.0])
    Diags = []
    for d in range(num):
        f = min(max((d - 5 + 8) / (band - 5), 0), 0.999)
        wt = 5 + int(np.searchsorted(fr, f, side="right"))
        Diags.append(rng.poisson(thr / (wt * wt - 4), n - d).astype(np.int32) + (1 if d < 5 else 0))
    w = np.exp(rng.normal(0, 0.2, n))
    w[rng.choice(n, 6, replace=False)] = np.nan
    return _finish(n, num, 5, Diags, w)


def test_sweep_beyond_compiled_widths(ctx):
    """maxww = 20 (the operator's default, callers.py:45) with an input that keeps widening to w = 11: no compiled-in
    program covers the executed steps, the table-driven kernel must take over and match the oracle at every cut point."""
    inp = _staircase_chromosome()
    st = compare_with_oracle(ctx, inp, [2], [5], 20, 0.1, 100, 16)
    assert st["frozen"] > 10 and st["spec_kernel"] == 0 and st["n_pixels"] > 0
