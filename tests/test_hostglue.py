"""Host glue of the Python mirror: the per-diagonal array lists of the reference's worker (scripts/pyHICCUPS:146-157)
become pointer tables through a small CPython helper (csrc/hp_pyhelper.c); a pure-Python loop is the stand-in when
the helper is not built.  Both must accept and reject the same inputs."""
import ctypes as C

import numpy as np
import pytest

from hicpeaks_b200 import _capi


def _both(seq, count, first, step, size, kind, want_table=True):
    out = []
    for fast in (True, False):
        saved = _capi._hpfast
        if not fast:
            _capi._hpfast = None
        elif saved is None:
            pytest.skip("_hpfast not built")
        try:
            tab = (C.c_void_p * max(count, 1))() if want_table else None
            r = _capi.first_nonconforming(seq, count, first, step, size, kind, tab)
            out.append((r, [tab[i] for i in range(count)] if (want_table and r < 0) else None))
        finally:
            _capi._hpfast = saved
    assert out[0] == out[1]
    return out[0]


def test_conforming_diagonals_give_their_addresses():
    n, num = 300, 40
    Dg = [np.arange(n - d, dtype=np.int32) for d in range(num)]
    r, ptrs = _both(Dg, num, n, -1, 4, "i")
    assert r == -1 and ptrs == [a.ctypes.data for a in Dg]
    cD = tuple(np.zeros(n - 3 - i) for i in range(num - 3))
    r, ptrs = _both(cD, num - 3, n - 3, -1, 8, "f")
    assert r == -1 and ptrs == [a.ctypes.data for a in cD]
    assert _both(Dg, num, n, -1, 4, "i", want_table=False)[0] == -1


@pytest.mark.parametrize("bad", ["dtype", "float", "length", "strided", "list", "2d", "byteorder", "empty-ok"])
def test_first_offender_is_reported(bad):
    n, num = 64, 9
    Dg = [np.zeros(n - d, dtype=np.int32) for d in range(num)]
    k = 5
    if bad == "dtype":
        Dg[k] = Dg[k].astype(np.int64)
    elif bad == "float":
        Dg[k] = Dg[k].astype(np.float32)
    elif bad == "length":
        Dg[k] = np.zeros(n - k + 1, dtype=np.int32)
    elif bad == "strided":
        Dg[k] = np.zeros(2 * (n - k), dtype=np.int32)[::2]
    elif bad == "list":
        Dg[k] = [0] * (n - k)
    elif bad == "2d":
        Dg[k] = np.zeros((1, n - k), dtype=np.int32)
    elif bad == "byteorder":
        Dg[k] = np.zeros(n - k, dtype=">i4")
    else:
        # zero-length arrays conform (a band as wide as the chromosome ends in an empty diagonal)
        assert _both([np.zeros(1, np.int32), np.zeros(0, np.int32)], 2, 1, -1, 4, "i")[0] == -1
        return
    assert _both(Dg, num, n, -1, 4, "i")[0] == k


def test_short_or_foreign_sequences():
    Dg = [np.zeros(4, dtype=np.int32)]
    assert _both(Dg, 3, 4, -1, 4, "i")[0] == 1          # fewer items than diagonals: the first missing index
    assert _both(iter(Dg), 1, 4, -1, 4, "i")[0] == 0    # not a list / tuple


def test_upload_narrowing_is_lossless_on_the_host():
    """hp_narrow_diagonal (host only, no GPU): every count diagonal is sent as the narrowest of u8 / u16 / i32 that
    holds it exactly -- never clipped, whatever the length (SIMD body + scalar tail) and wherever the large value sits."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        n = int(rng.integers(0, 20000)) if trial else 0
        a = rng.integers(0, 200, n).astype(np.int32)
        kind = trial % 5
        if n and kind == 1:
            a[rng.integers(0, n)] = 255
        if n and kind == 2:
            a[rng.integers(0, n)] = 256 + int(rng.integers(0, 65280))
        if n and kind == 3:
            a[rng.integers(0, n)] = 65536 + int(rng.integers(0, 2 ** 30))
        if n and kind == 4:
            a[rng.integers(0, n)] = -1 - int(rng.integers(0, 1000))
        out = _capi.narrow_diagonal(a, misalign=(0, 16, 4, 1)[trial % 4])      # streaming stores (aligned) and plain stores
        want = np.int32 if (n and (a.min() < 0 or a.max() > 65535)) else np.uint16 if (n and a.max() > 255) else np.uint8
        assert out.dtype == want, (trial, n, out.dtype, want)
        assert np.array_equal(out.astype(np.int64), a.astype(np.int64))
    edge = np.zeros(8192 * 3 + 5, dtype=np.int32)
    edge[-1] = 70000                                     # the offending value in the scalar tail of the last block
    assert _capi.narrow_diagonal(edge).dtype == np.int32
    edge[-1] = 300
    assert _capi.narrow_diagonal(edge).dtype == np.uint16
