"""compute-sanitizer target: the worker-level upload + one scoring call + downloads on a small chromosome."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import synth_chromosome
inp = synth_chromosome(1237, 131, 5, maxww=10, seed=11, scale=60.0)
Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
Dg[0][:] += 300; Dg[7][5] = 70000
with _capi.Context(0) as ctx:
    ctx.upload_counts(inp["n"], inp["num"], 5, Dg, inp["weights"])
    P = ctx.make_params([2], [5], 10, 0.1, 120, 16)
    S = ctx.hiccups(P)
    sv = ctx.survivors(); g = ctx.gaps()
    print("ok", S.n_pixels, S.frozen_w, S.n_survivors, int(g.sum()), S.spec_kernel)
    print("fast kernel", S.fast_kernel, "exact records", S.n_exact)
    P = ctx.make_params([2], [5], 10, 0.1, 120, 16, exact_sums=True)
    S = ctx.hiccups(P)
    print("ok exact-order spec kernel", S.n_pixels, S.n_survivors, S.fast_kernel)
    P = ctx.make_params([2], [5], 10, 0.1, 120, 16)
    S = ctx.score(P)
    ctx.comm_init(1, 0, None)
    print("device merge ms", ctx.allreduce_hist([ctx], 1))
    S = ctx.fdr()
    print("ok genome-scope merge", S.n_survivors)
    P = ctx.make_params([2], [5], 10, 0.1, 120, 16, generic_kernel=True)
    S = ctx.hiccups(P)
    print("ok generic", S.n_pixels, S.n_survivors)
