"""GPU: error behaviour of the C ABI (status codes + messages, no exceptions across the boundary, no silent fallback)
and of the Python mirror (the reference's own failure modes)."""
import numpy as np
import pytest

from hicpeaks_b200 import _capi, callers
from hicpeaks_b200.synth import synth_chromosome

pytestmark = pytest.mark.gpu


def _upload(ctx, inp):
    Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
    ctx.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"])


def test_call_order_and_arguments():
    inp = synth_chromosome(300, 40, 5, maxww=8, seed=1, scale=60.0)
    with _capi.Context(0) as ctx:
        P = ctx.make_params([2], [5], 8, 0.1, 40, 16)
        with pytest.raises(_capi.EngineError) as e:
            ctx.score(P)                                        # nothing uploaded yet
        assert e.value.code == _capi.HP_ERR_STATE and "upload" in str(e.value)
        for early in (ctx.upload_bytes, ctx.timer_stop):        # measurement aids keep the same conventions
            with pytest.raises(_capi.EngineError) as e:
                early()
            assert e.value.code == _capi.HP_ERR_STATE
        _upload(ctx, inp)
        assert 0 < ctx.upload_bytes() < 4 * sum(d.size for d in inp["Diags"]) + 8 * inp["n"] + 16 * inp["num"] + 1
        ctx.timer_start()
        assert ctx.timer_stop() >= 0.0
        with pytest.raises(_capi.EngineError) as e:
            ctx.fdr()                                           # fdr before score
        assert e.value.code == _capi.HP_ERR_STATE
        for bad in (ctx.make_params([2], [9], 8, 0.1, 40, 16),          # ww > maxww
                    ctx.make_params([2], [5], 25, 0.1, 40, 16),         # maxww beyond HP_MAX_WW
                    ctx.make_params([2, 2], [5, 6], 8, 0.1, 40, 16),    # duplicate pw
                    ctx.make_params([2], [4], 8, 0.1, 40, 16)):         # band uploaded for min(ww) = 5
            with pytest.raises(_capi.EngineError) as e:
                ctx.score(bad)
            assert e.value.code == _capi.HP_ERR_INVALID
        S = ctx.hiccups(P)                                      # the context is still usable after errors
        assert S.n_pixels > 0
        with pytest.raises(_capi.EngineError) as e:
            ctx.make_params([2], [5], 8, 0.1, 40, 16, bhfdr=True).npw and ctx.score(ctx.make_params([1, 2], [3, 5], 8, 0.1, 40, 16, bhfdr=True))
        assert e.value.code == _capi.HP_ERR_INVALID
    with pytest.raises(ValueError):
        _capi.Context.make_params([1] * 9, [3] * 9, 10, 0.1, 40, 16)   # more pairs than HP_MAX_PW
    with pytest.raises(_capi.EngineError) as e:
        _capi.Context(99)                                       # no such device
    assert e.value.code == _capi.HP_ERR_INVALID


def test_chunk_overflow_is_reported_not_clipped():
    """An expected value beyond the last lambda-chunk edge of the context must raise, never be mis-binned."""
    inp = synth_chromosome(300, 40, 5, maxww=8, seed=1, scale=3000.0)
    with _capi.Context(0, max_chunks=12) as ctx:               # edges up to 2^(11/3) = 12.7
        _upload(ctx, inp)
        with pytest.raises(_capi.EngineError) as e:
            ctx.score(ctx.make_params([2], [5], 8, 0.1, 40, 16))
        assert e.value.code == _capi.HP_ERR_CHUNK_OVERFLOW


def test_reference_failure_modes_are_kept():
    """No band pixel at all / every pixel resolved before the sweep ends: the reference raises (callers.py:205-208)."""
    inp = synth_chromosome(300, 40, 5, maxww=8, seed=1, scale=60.0)
    inp["Diags"] = [np.zeros_like(d) for d in inp["Diags"]]
    with pytest.raises(ValueError):
        callers.hiccups_from_counts(inp["weights"], inp["n"], inp["Diags"], inp["num"], "1", pw=[2], ww=[5], maxww=8,
                                    maxapart=400000, res=10000, min_local_reads=16)
    with pytest.raises(ValueError):
        callers.hiccups(None, None, inp["biases"], inp["biases"], {0: 1.0}, inp["n"], inp["Diags"], inp["cDiags"], inp["num"], "1",
                        pw=[2], ww=[5], maxww=8, maxapart=400000, res=10000)       # IR keyed wrongly
