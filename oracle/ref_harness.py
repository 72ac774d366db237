"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference (``/root/reference/hicpeaks/callers.py``)
in the build container and captures its intermediates.  Used by ``oracle/make_golden.py`` only; the
reference tree does not exist on the GPU box, so nothing under ``tests/`` or ``bench.py`` imports this.

How: the module is imported by path with the ``statsmodels`` stand-in on
``sys.path``; ``callers.lambdachunk`` and ``callers.multipletests`` are wrapped by spies that copy the
caller frame's locals (``pi, wi, fl, xi, yi, Evalues, Ovalues, bSV, bEV, vxi, vyi``) -- the reference
source itself is not touched.  Inputs are built the way the worker does (scripts/pyHICCUPS:146-166).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import warnings

import numpy as np
from scipy import sparse

REF_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def load_reference_callers():
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    spec = importlib.util.spec_from_file_location("ref_callers", os.path.join(REF_ROOT, "hicpeaks", "callers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_apa():
    spec = importlib.util.spec_from_file_location("ref_apa", os.path.join(REF_ROOT, "hicpeaks", "apa.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_matrices(inp):
    """M and cM as the worker builds them (pyHICCUPS:148, :160)."""
    num, mw = inp["num"], inp["min_ww"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = sparse.diags([d.astype(float) for d in inp["Diags"]], np.arange(num), format="csr")
        cM = sparse.diags(inp["cDiags"], np.arange(mw, num), format="csr")
    return M, cM


def run_hiccups(inp, res, chrom="1", **kw):
    """Call reference ``hiccups`` and return (final_table, captures).

    captures[(pi, fl)] = dict(x, y, E, O, p, q, bSV, bEV) in the reference's own pixel order;
    captures['pixels'] = (vxi, vyi)."""
    ref = load_reference_callers()
    M, cM = reference_matrices(inp)
    cap = {}
    state = {}
    orig_lc, orig_mt = ref.lambdachunk, ref.multipletests

    def spy_lc(E):
        f = sys._getframe(1).f_locals
        chunks = orig_lc(E)
        key = (int(f["pi"]), f["fl"])
        cap["pixels"] = (np.array(f["vxi"]), np.array(f["vyi"]))
        cap[key] = dict(x=np.array(f["xi"]), y=np.array(f["yi"]), E=np.array(f["Evalues"]),
                        O=np.array(f["Ovalues"][f["fl"]]),
                        bSV=np.array(f["bSV"][f["pi"]][f["fl"]]), bEV=np.array(f["bEV"][f["pi"]][f["fl"]]),
                        p=np.ones(len(f["xi"])), q=np.ones(len(f["xi"])),
                        chunk=np.zeros(len(f["xi"]), dtype=np.int64), numbin=len(chunks))
        state["key"] = key
        state["todo"] = [(i + 1, c[2]) for i, c in enumerate(chunks) if c[2].size > 0]
        return chunks

    def spy_mt(pvals, *a, **k):
        out = orig_mt(pvals, *a, **k)
        if "key" in state and state["todo"]:
            ci, idx = state["todo"].pop(0)
            c = cap[state["key"]]
            c["p"][idx] = pvals
            c["q"][idx] = out[1]
            c["chunk"][idx] = ci
        return out

    ref.lambdachunk, ref.multipletests = spy_lc, spy_mt
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            table = ref.hiccups(M, cM, inp["biases"], inp["biases"], dict(inp["IR"]), inp["n"],
                                inp["Diags"], [c.copy() for c in inp["cDiags"]], inp["num"], chrom,
                                res=res, **kw)
    finally:
        ref.lambdachunk, ref.multipletests = orig_lc, orig_mt
    return table, cap


def run_bhfdr(inp, res, chrom="1", **kw):
    """Call reference ``bhfdr``; captures = dict(x, y, E? , p, q) via the multipletests spy."""
    ref = load_reference_callers()
    M, cM = reference_matrices(inp)
    cap = {}
    orig_mt = ref.multipletests

    def spy_mt(pvals, *a, **k):
        f = sys._getframe(1).f_locals
        out = orig_mt(pvals, *a, **k)
        cap.update(x=np.array(f["xi"]), y=np.array(f["yi"]), E=np.array(f["Evalues"]),
                   O=np.array(f["Ovalues"]), p=np.array(pvals), q=np.array(out[1]), reject=np.array(out[0]))
        return out

    ref.multipletests = spy_mt
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            table = ref.bhfdr(M, cM, inp["biases"], inp["biases"], dict(inp["IR"]), inp["n"],
                              inp["Diags"], [c.copy() for c in inp["cDiags"]], inp["num"], chrom,
                              res=res, **kw)
    finally:
        ref.multipletests = orig_mt
    return table, cap
