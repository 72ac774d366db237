"""One cfg2 chromosome through the engine a few times (for ncu / timing of the score kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import synth_chromosome

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
exact = len(sys.argv) > 3 and sys.argv[3] == "exact"
union = len(sys.argv) > 3 and sys.argv[3] == "union"
pw, ww = ([1, 2, 4], [3, 5, 7]) if union else ([2], [5])
inp = synth_chromosome(n, 500, min(ww), maxww=10, seed=17 if not union else 333)
Dg = [np.ascontiguousarray(d, dtype=np.int32) for d in inp["Diags"]]
with _capi.Context(0) as ctx:
    ctx.upload_counts(inp["n"], inp["num"], min(ww), Dg, inp["weights"])
    P = ctx.make_params(pw, ww, 10, 0.1, 500, 16, exact_sums=exact)
    for k in range(reps):
        S = ctx.hiccups(P)
        print("rep %d: frozen %d steps %d fast %d n_exact %d ms_levels %.3f ms_score %.3f ms_exact %.3f ms_fdr %.3f cand %d surv %d" % (
            k, S.frozen_w, S.n_steps, S.fast_kernel, S.n_exact, S.ms_levels, S.ms_score, S.ms_exact, S.ms_fdr, S.n_candidates, S.n_survivors))
