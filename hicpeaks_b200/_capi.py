"""ctypes binding of ``include/hicpeaks_b200.h`` (the C-ABI drop-in boundary).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a only) as
``hicpeaks_b200/libhicpeaks_b200.so``.  There is no Python / CPU implementation behind this module:
if the library is missing, or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HP_MAX_PW = 8
HP_MAX_WW = 20
HP_MAX_STEPS = 160

HP_OK = 0
HP_ERR_INVALID = -1
HP_ERR_CUDA = -2
HP_ERR_NO_DEVICE = -3
HP_ERR_STATE = -4
HP_ERR_EMPTY_REFIDX = -5
HP_ERR_CHUNK_OVERFLOW = -6
HP_ERR_CAPACITY = -7

PF_GENERIC_KERNEL = 1
PF_BHFDR = 2
PF_EXACT_SUMS = 4
SF_VALID_K, SF_VALID_Y, SF_REJECT_K, SF_REJECT_Y, SF_CEMY_NONZERO = 1, 2, 4, 8, 16

LIB_NAME = "libhicpeaks_b200.so"
LIB_PATH = os.environ.get("HICPEAKS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

# every symbol include/hicpeaks_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "hp_abi_version", "hp_device_count", "hp_ctx_create", "hp_ctx_destroy", "hp_last_error",
    "hp_band_upload", "hp_band_upload_counts", "hp_upload_bytes", "hp_narrow_diagonal", "hp_program_dump", "hp_timer_start", "hp_timer_stop", "hp_dump_band", "hp_hiccups_score", "hp_hiccups_fdr", "hp_hiccups", "hp_get_survivors",
    "hp_hist_bins", "hp_hist_export", "hp_hist_import", "hp_get_summary", "hp_comm_unique_id", "hp_comm_init", "hp_comm_destroy", "hp_allreduce_hist", "hp_ctx_trim", "hp_get_gaps", "hp_dump_levels", "hp_dump_plane", "hp_get_chunk_table", "hp_poisson_sf",
    "hp_apa_upload", "hp_apa_windows", "hp_apa_load_windows", "hp_apa_accumulate", "hp_apa_get_windows",
]


class BandDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("num", C.c_int32), ("bal_first", C.c_int32),
                ("raw_diags", C.POINTER(C.c_void_p)), ("bal_diags", C.POINTER(C.c_void_p)),
                ("ir", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p)]


class CountsDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("num", C.c_int32), ("bal_first", C.c_int32),
                ("raw_diags", C.POINTER(C.c_void_p)), ("weights", C.c_void_p)]


class ApaDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("num", C.c_int32), ("reserved", C.c_int32), ("bal_diags", C.POINTER(C.c_void_p))]


class HiccupsParams(C.Structure):
    _fields_ = [("npw", C.c_int32), ("pw", C.c_int32 * HP_MAX_PW), ("ww", C.c_int32 * HP_MAX_PW),
                ("maxww", C.c_int32), ("min_local_reads", C.c_int32), ("maxapart_bins", C.c_int64),
                ("sig", C.c_double), ("dump", C.c_int32), ("flags", C.c_int32)]


class StepStat(C.Structure):
    _fields_ = [("p", C.c_int32), ("w", C.c_int32), ("resolved", C.c_int64),
                ("valid_ratio", C.c_double), ("left_ratio", C.c_double)]


class LfStat(C.Structure):
    _fields_ = [("n_valid", C.c_int64), ("e_max", C.c_double), ("numbin", C.c_int32),
                ("reserved", C.c_int32), ("n_reject", C.c_int64)]


class HiccupsSummary(C.Structure):
    _fields_ = [("n_pixels", C.c_int64), ("band_pixels", C.c_int64), ("frozen_w", C.c_int32),
                ("n_steps", C.c_int32), ("steps", StepStat * HP_MAX_STEPS),
                ("lf", (LfStat * 2) * HP_MAX_PW), ("n_candidates", C.c_int64), ("n_survivors", C.c_int64),
                ("ms_levels", C.c_float), ("ms_score", C.c_float), ("ms_fdr", C.c_float), ("ms_total", C.c_float),
                ("launches", C.c_int32), ("spec_kernel", C.c_int32),
                ("ms_exact", C.c_float), ("fast_kernel", C.c_int32), ("n_exact", C.c_int64)]


SURVIVOR_DTYPE = np.dtype([("r", "<i4"), ("c", "<i4"), ("pair", "<i4"), ("flags", "<u4"), ("obs", "<f8"),
                           ("ice", "<f8"), ("e", "<f8", (2,)), ("p", "<f8", (2,)), ("q", "<f8", (2,))])
assert SURVIVOR_DTYPE.itemsize == 80


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hicpeaks_b200 [%d]: %s" % (code, msg))
        self.code = code


_lib = None


def load_library(path: str | None = None):
    """Load the C-ABI library (no GPU needed to load; compute calls need a B200)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineError(HP_ERR_NO_DEVICE, "%s not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.hp_abi_version.restype = C.c_int
    lib.hp_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.hp_ctx_create.argtypes = [C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.hp_ctx_destroy.argtypes = [vp]
    lib.hp_ctx_destroy.restype = None
    lib.hp_last_error.argtypes = [vp]
    lib.hp_last_error.restype = C.c_char_p
    lib.hp_band_upload.argtypes = [vp, C.POINTER(BandDesc)]
    lib.hp_band_upload_counts.argtypes = [vp, C.POINTER(CountsDesc)]
    lib.hp_dump_band.argtypes = [vp, i32, vp, i64]
    lib.hp_upload_bytes.argtypes = [vp, C.POINTER(i64)]
    lib.hp_narrow_diagonal.argtypes = [vp, i64, vp, C.POINTER(i32)]
    lib.hp_program_dump.argtypes = [C.POINTER(HiccupsParams), C.POINTER(i32), vp, vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64)]
    lib.hp_timer_start.argtypes = [vp]
    lib.hp_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    lib.hp_hiccups_score.argtypes = [vp, C.POINTER(HiccupsParams), C.POINTER(HiccupsSummary)]
    lib.hp_hiccups_fdr.argtypes = [vp, vp, C.POINTER(HiccupsSummary)]
    lib.hp_hiccups.argtypes = [vp, C.POINTER(HiccupsParams), C.POINTER(HiccupsSummary)]
    lib.hp_get_survivors.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.hp_hist_bins.argtypes = [vp, C.POINTER(i64)]
    lib.hp_hist_export.argtypes = [vp, vp, i64]
    lib.hp_hist_import.argtypes = [vp, vp, i64]
    lib.hp_get_summary.argtypes = [vp, C.POINTER(HiccupsSummary)]
    lib.hp_comm_unique_id.argtypes = [vp]
    lib.hp_comm_init.argtypes = [vp, i32, i32, vp]
    lib.hp_comm_destroy.argtypes = [vp]
    lib.hp_allreduce_hist.argtypes = [vp, vp, i32, i32, C.POINTER(C.c_float)]
    lib.hp_ctx_trim.argtypes = [vp]
    lib.hp_get_gaps.argtypes = [vp, vp, i64]
    lib.hp_dump_levels.argtypes = [vp, vp, i64]
    lib.hp_dump_plane.argtypes = [vp, i32, i32, i32, vp, i64]
    lib.hp_get_chunk_table.argtypes = [vp, i32, i32, C.POINTER(i32), vp, vp, vp, vp, i64]
    lib.hp_poisson_sf.argtypes = [vp, vp, vp, vp, i64]
    lib.hp_apa_upload.argtypes = [vp, C.POINTER(ApaDesc)]
    lib.hp_apa_windows.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.hp_apa_load_windows.argtypes = [vp, vp, i64, i32, vp]
    lib.hp_apa_accumulate.argtypes = [vp, vp, i64, vp, i32]
    lib.hp_apa_get_windows.argtypes = [vp, vp, i64, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("hp_ctx_destroy", "hp_last_error"):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def program_dump(pw, ww, maxww):
    """Host-only: [(p, w, [(a, b, is_y, is_r), ...])] -- the sweep program the engine builds for (pw, ww, maxww)."""
    lib = load_library()
    P = Context.make_params(pw, ww, maxww, 0.1, 1, 1)
    sp = np.zeros(HP_MAX_STEPS, dtype=np.int32)
    sw = np.zeros(HP_MAX_STEPS, dtype=np.int32)
    oe = np.zeros(HP_MAX_STEPS, dtype=np.int32)
    ns, no = C.c_int32(), C.c_int64()
    rc = lib.hp_program_dump(C.byref(P), C.byref(ns), _ptr(sp), _ptr(sw), _ptr(oe), 0, None, None, None, None, C.byref(no))
    if rc != HP_OK:
        raise EngineError(rc, (lib.hp_last_error(None) or b"").decode())
    a = np.zeros(max(no.value, 1), dtype=np.int8)
    b = np.zeros_like(a)
    y = np.zeros(a.size, dtype=np.uint8)
    r = np.zeros(a.size, dtype=np.uint8)
    rc = lib.hp_program_dump(C.byref(P), C.byref(ns), _ptr(sp), _ptr(sw), _ptr(oe), a.size, _ptr(a), _ptr(b), _ptr(y), _ptr(r), C.byref(no))
    if rc != HP_OK:
        raise EngineError(rc, (lib.hp_last_error(None) or b"").decode())
    out, lo = [], 0
    for s in range(ns.value):
        hi = int(oe[s])
        out.append((int(sp[s]), int(sw[s]), [(int(a[i]), int(b[i]), bool(y[i]), bool(r[i])) for i in range(lo, hi)]))
        lo = hi
    return out


def narrow_diagonal(counts, misalign: int = 0) -> np.ndarray:
    """Host-only: the narrowed form (uint8 / uint16 / int32 array) hp_band_upload_counts sends for one count diagonal.
    ``misalign``: byte offset of the destination from a 64-byte boundary (0: the streaming-store path of the library,
    anything not a multiple of 16: its plain-store path)."""
    a = np.ascontiguousarray(counts, dtype=np.int32)
    raw = np.empty(max(a.size, 1) * 4 + 128, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64 + int(misalign)
    buf = raw[off: off + max(a.size, 1) * 4]
    es = C.c_int32()
    rc = load_library().hp_narrow_diagonal(_ptr(a), a.size, buf.ctypes.data_as(C.c_void_p), C.byref(es))
    if rc != HP_OK:
        raise EngineError(rc, "hp_narrow_diagonal failed")
    dt = {1: np.uint8, 2: np.uint16, 4: np.int32}[es.value]
    return buf[: a.size * es.value].copy().view(dt)


def chunk_edges(max_chunks: int) -> np.ndarray:
    """Upper lambda-chunk edges rv_i, i = 1..max_chunks, evaluated exactly as the reference does
    (/root/reference/hicpeaks/callers.py:33-37)."""
    out = np.empty(max_chunks, dtype=np.float64)
    for i in range(1, max_chunks + 1):
        out[i - 1] = 1 if i == 1 else np.power(2, ((i - 1) / 3.))
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


try:                                     # built by __graft_entry__.build() (csrc/hp_pyhelper.c); host glue only
    from . import _hpfast
except ImportError:                      # pragma: no cover - same checks, ~50x slower
    _hpfast = None

_KIND = {"i": np.signedinteger, "f": np.floating}


def first_nonconforming(seq, count, first_len, len_step, itemsize, kind, table=None):
    """Index of the first of ``seq[:count]`` that is not a C-contiguous 1-D native array of ``itemsize``-byte
    ``kind`` ('i' / 'f') elements with ``first_len + i * len_step`` elements, or -1 when all conform.  ``table``
    (a ctypes ``c_void_p`` array) receives the data pointers of the conforming prefix."""
    if _hpfast is not None:
        return _hpfast.collect(seq, count, first_len, len_step, itemsize, kind, C.addressof(table) if table is not None else 0)
    if not isinstance(seq, (list, tuple)):
        return 0
    if len(seq) < count:
        return len(seq)
    for i in range(count):
        a = seq[i]
        if (not isinstance(a, np.ndarray) or a.ndim != 1 or a.dtype.itemsize != itemsize or not a.dtype.isnative
                or not np.issubdtype(a.dtype, _KIND[kind]) or not a.flags.c_contiguous or a.size != first_len + i * len_step):
            return i
        if table is not None:
            table[i] = a.ctypes.data
    return -1


class Context:
    """One engine context = one GPU + one stream (see include/hicpeaks_b200.h)."""

    def __init__(self, device: int = 0, max_chunks: int = 52):
        self.lib = load_library()
        self.max_chunks = max_chunks
        self._h = C.c_void_p()
        edges = chunk_edges(max_chunks)
        rc = self.lib.hp_ctx_create(device, max_chunks, _ptr(edges), C.byref(self._h))
        if rc != HP_OK:
            raise EngineError(rc, (self.lib.hp_last_error(None) or b"").decode())
        self.n = self.num = 0
        self.params = None

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None
            self.lib.hp_ctx_destroy(h)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != HP_OK:
            raise EngineError(rc, (self.lib.hp_last_error(self._h) or b"").decode())

    # -- input ---------------------------------------------------------------------------------
    def upload(self, n, num, bal_first, Diags, cDiags, ir, b1, b2):
        """Diags: ``num`` int32 arrays (length n-d); cDiags: ``num - bal_first`` float64 arrays;
        ir: float64[num - bal_first]; b1, b2: float64[n].  Arrays must already be C-contiguous of the
        right dtype (the caller normalises); they are only read during this call."""
        keep = []
        rp = (C.c_void_p * num)()
        d = first_nonconforming(Diags, num, n, -1, 4, "i", rp)
        if d >= 0:
            raise ValueError("Diags[%d] must be contiguous int32 of length %d" % (d, n - d))
        nb = num - bal_first
        bp = (C.c_void_p * nb)()
        i = first_nonconforming(cDiags, nb, n - bal_first, -1, 8, "f", bp)
        if i >= 0:
            raise ValueError("cDiags[%d] must be contiguous float64 of length %d" % (i, n - bal_first - i))
        ir = np.ascontiguousarray(ir, dtype=np.float64)
        b1 = np.ascontiguousarray(b1, dtype=np.float64)
        b2 = np.ascontiguousarray(b2, dtype=np.float64)
        if ir.size != nb or b1.size != n or b2.size != n:
            raise ValueError("ir / bias length mismatch")
        keep += [ir, b1, b2]
        desc = BandDesc(n, num, bal_first, C.cast(rp, C.POINTER(C.c_void_p)), C.cast(bp, C.POINTER(C.c_void_p)),
                        ir.ctypes.data, b1.ctypes.data, b2.ctypes.data)
        self._check(self.lib.hp_band_upload(self._h, C.byref(desc)))
        self.n, self.num = int(n), int(num)

    def upload_counts(self, n, num, bal_first, Diags, weights):
        """Worker-level input: ``Diags`` (``num`` contiguous int32 arrays) and the bin weights (float64[n], NaN for
        masked bins); the balanced band, IR and the biases are derived on the GPU (scripts/pyHICCUPS:143-166)."""
        rp = (C.c_void_p * num)()
        d = first_nonconforming(Diags, num, n, -1, 4, "i", rp)
        if d >= 0:
            raise ValueError("Diags[%d] must be contiguous int32 of length %d" % (d, n - d))
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.size != n:
            raise ValueError("weights length mismatch")
        desc = CountsDesc(n, num, bal_first, C.cast(rp, C.POINTER(C.c_void_p)), w.ctypes.data)
        self._check(self.lib.hp_band_upload_counts(self._h, C.byref(desc)))
        self.n, self.num = int(n), int(num)

    def upload_bytes(self):
        """Bytes the last upload moved over PCIe (count diagonals travel narrowed to u8 / u16 where they fit)."""
        v = C.c_int64()
        self._check(self.lib.hp_upload_bytes(self._h, C.byref(v)))
        return v.value

    def timer_start(self):
        self._check(self.lib.hp_timer_start(self._h))

    def timer_stop(self):
        """Milliseconds on the device clock since timer_start() (CUDA events on this context's stream)."""
        ms = C.c_float()
        self._check(self.lib.hp_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def dump_band(self, what):
        shape = {0: (self.num,), 1: (self.n,), 2: (self.num, self.n)}[what]
        out = np.empty(shape, dtype=np.float64)
        self._check(self.lib.hp_dump_band(self._h, what, _ptr(out), out.size))
        return out

    # -- scoring -------------------------------------------------------------------------------
    @staticmethod
    def make_params(pw, ww, maxww, sig, maxapart_bins, min_local_reads, dump=False, generic_kernel=False, bhfdr=False,
                    exact_sums=False):
        if len(pw) != len(ww) or not 1 <= len(pw) <= HP_MAX_PW:
            raise ValueError("need 1..%d (pw, ww) pairs" % HP_MAX_PW)
        P = HiccupsParams()
        P.npw = len(pw)
        for i, (p, w) in enumerate(zip(pw, ww)):
            P.pw[i], P.ww[i] = int(p), int(w)
        P.maxww, P.min_local_reads = int(maxww), int(min_local_reads)
        P.maxapart_bins, P.sig, P.dump = int(maxapart_bins), float(sig), int(bool(dump))
        P.flags = ((PF_GENERIC_KERNEL if generic_kernel else 0) | (PF_BHFDR if bhfdr else 0) |
                   (PF_EXACT_SUMS if exact_sums else 0))
        return P

    def score(self, P):
        S = HiccupsSummary()
        self.params = P
        self._check(self.lib.hp_hiccups_score(self._h, C.byref(P), C.byref(S)))
        return S

    def fdr(self, numbin_override=None):
        S = HiccupsSummary()
        ov = None
        if numbin_override is not None:
            ov = np.ascontiguousarray(numbin_override, dtype=np.int32)
        self._check(self.lib.hp_hiccups_fdr(self._h, None if ov is None else _ptr(ov), C.byref(S)))
        return S

    def hiccups(self, P):
        S = HiccupsSummary()
        self.params = P
        self._check(self.lib.hp_hiccups(self._h, C.byref(P), C.byref(S)))
        return S

    def survivors(self):
        cnt = C.c_int64()
        self._check(self.lib.hp_get_survivors(self._h, None, 0, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=SURVIVOR_DTYPE)
        if cnt.value:
            self._check(self.lib.hp_get_survivors(self._h, _ptr(out), cnt.value, C.byref(cnt)))
        return out

    def hist_export(self):
        """(npw * 2, total_bins) int64 histograms of the last score() call."""
        tb = C.c_int64()
        self._check(self.lib.hp_hist_bins(self._h, C.byref(tb)))
        out = np.zeros((self.params.npw * 2, tb.value), dtype=np.int64)
        self._check(self.lib.hp_hist_export(self._h, _ptr(out), out.size))
        return out

    def hist_import(self, hist):
        h = np.ascontiguousarray(hist, dtype=np.int64)
        self._check(self.lib.hp_hist_import(self._h, _ptr(h), h.size))

    # -- genome-wide FDR: one NCCL all-reduce of the histograms (include/hicpeaks_b200.h) ----------------------------
    def comm_init(self, nranks=1, rank=0, unique_id=None):
        """Bind this context to rank ``rank`` of an ``nranks``-GPU communicator (``unique_id``: bytes from
        ``comm_unique_id()`` on rank 0; not needed for nranks == 1)."""
        buf = None
        if nranks > 1:
            if unique_id is None or len(unique_id) != COMM_ID_BYTES:
                raise ValueError("unique_id must be the %d bytes of comm_unique_id()" % COMM_ID_BYTES)
            buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self.lib.hp_comm_init(self._h, int(nranks), int(rank), buf))

    def comm_destroy(self):
        self._check(self.lib.hp_comm_destroy(self._h))

    def allreduce_hist(self, ctxs, npw):
        """Merge the histograms / E.max / valid counts of the scored contexts ``ctxs`` (this process, this GPU) and of all
        ranks; returns the device milliseconds of the merge.  Every rank calls it with the run's number of (pw, ww) pairs;
        ``ctxs`` may be empty."""
        arr = (C.c_void_p * max(1, len(ctxs)))(*[c._h for c in ctxs])
        ms = C.c_float()
        self._check(self.lib.hp_allreduce_hist(self._h, arr, len(ctxs), int(npw), C.byref(ms)))
        return float(ms.value)

    def summary(self):
        """Summary of the last score() as it stands now (after allreduce_hist: genome-wide e_max / numbin / n_valid)."""
        S = HiccupsSummary()
        self._check(self.lib.hp_get_summary(self._h, C.byref(S)))
        return S

    def trim(self):
        """Give back the upload scratch (landing zone, pinned staging); the band and the last results stay valid."""
        self._check(self.lib.hp_ctx_trim(self._h))

    def gaps(self):
        out = np.empty(self.n, dtype=np.bool_)
        self._check(self.lib.hp_get_gaps(self._h, _ptr(out), self.n))
        return out

    # -- inspection ----------------------------------------------------------------------------
    def dump_levels(self):
        out = np.empty((self.num, self.n), dtype=np.uint8)
        self._check(self.lib.hp_dump_levels(self._h, _ptr(out), out.size))
        return out

    def dump_plane(self, pair, background, what):
        out = np.empty((self.num, self.n), dtype=np.float64)
        self._check(self.lib.hp_dump_plane(self._h, pair, background, what, _ptr(out), out.size))
        return out

    def chunk_table(self, pair, background):
        nb = C.c_int32()
        widths = np.zeros(64, dtype=np.int32)
        self._check(self.lib.hp_get_chunk_table(self._h, pair, background, C.byref(nb), _ptr(widths), None, None, None, 0))
        widths = widths[:nb.value]
        tot = int(widths.sum())
        hist = np.zeros(tot, dtype=np.int64)
        p = np.zeros(tot, dtype=np.float64)
        q = np.zeros(tot, dtype=np.float64)
        if tot:
            self._check(self.lib.hp_get_chunk_table(self._h, pair, background, C.byref(nb), _ptr(widths), _ptr(hist),
                                                    _ptr(p), _ptr(q), tot))
        off = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
        return nb.value, widths, off, hist, p, q

    # -- APA ----------------------------------------------------------------------------------
    def apa_upload(self, n, diags):
        """diags[d]: balanced diagonal d (float64, length n - d, NaN kept), d = 0 .. len(diags) - 1."""
        num = len(diags)
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in diags]
        ptrs = (C.c_void_p * num)()
        for d, a in enumerate(keep):
            if a.size != n - d:
                raise ValueError("diagonal %d must have %d entries" % (d, n - d))
            ptrs[d] = a.ctypes.data
        desc = ApaDesc(n, num, 0, C.cast(ptrs, C.POINTER(C.c_void_p)))
        self._check(self.lib.hp_apa_upload(self._h, C.byref(desc)))

    def apa_windows(self, pos_i, pos_j, w):
        pi = np.ascontiguousarray(pos_i, dtype=np.int32)
        pj = np.ascontiguousarray(pos_j, dtype=np.int32)
        valid = np.zeros(pi.size, dtype=np.uint8)
        mean = np.zeros(pi.size, dtype=np.float64)
        self._check(self.lib.hp_apa_windows(self._h, _ptr(pi), _ptr(pj), pi.size, int(w), _ptr(valid), _ptr(mean)))
        return valid.astype(bool), mean

    def apa_load_windows(self, wins, w):
        a = np.ascontiguousarray(wins, dtype=np.float64)
        mean = np.zeros(a.shape[0], dtype=np.float64)
        self._check(self.lib.hp_apa_load_windows(self._h, _ptr(a), a.shape[0], int(w), _ptr(mean)))
        return mean

    def apa_accumulate(self, sel, acc, init):
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        self._check(self.lib.hp_apa_accumulate(self._h, _ptr(sel), sel.size, _ptr(acc), int(bool(init))))

    def apa_get_windows(self, sel, w):
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        side = 2 * w + 1
        out = np.empty((sel.size, side, side), dtype=np.float64)
        if sel.size:
            self._check(self.lib.hp_apa_get_windows(self._h, _ptr(sel), sel.size, _ptr(out)))
        return out

    def poisson_sf(self, k, mu):
        k = np.ascontiguousarray(k, dtype=np.float64)
        mu = np.ascontiguousarray(np.broadcast_to(mu, k.shape), dtype=np.float64)
        out = np.empty_like(k)
        self._check(self.lib.hp_poisson_sf(self._h, _ptr(k), _ptr(mu), _ptr(out), k.size))
        return out


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """NCCL unique id (rank 0 creates it and hands it to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.hp_comm_unique_id(buf)
    if rc != HP_OK:
        raise EngineError(rc, (lib.hp_last_error(None) or b"").decode())
    return buf.raw


def device_count() -> int:
    lib = load_library()
    n = C.c_int()
    lib.hp_device_count(C.byref(n))
    return n.value
