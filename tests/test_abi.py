"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from hicpeaks_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(hp_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    decl = declared_symbols()
    assert decl, "no declarations found"
    for name in sorted(decl):
        assert hasattr(lib, name), "missing export: " + name
    assert decl == set(_capi.SYMBOLS), decl ^ set(_capi.SYMBOLS)
    assert _capi.load_library().hp_abi_version() == 2


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_capi.HiccupsParams) == 104
    assert ctypes.sizeof(_capi.StepStat) == 32
    assert ctypes.sizeof(_capi.LfStat) == 32
    assert ctypes.sizeof(_capi.HiccupsSummary) == 24 + 160 * 32 + 16 * 32 + 16 + 16 + 8 + 16
    assert _capi.SURVIVOR_DTYPE.itemsize == 80


def test_no_gpu_means_loud_failure():
    import pytest
    n = _capi.device_count()
    if n > 0:
        pytest.skip("GPU present")
    with pytest.raises(_capi.EngineError):
        _capi.Context(0)
