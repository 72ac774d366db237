"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- times the UNMODIFIED reference ``hiccups()`` (/root/reference, imported by
path with the statsmodels stand-in, oracle/ref_harness.py) on one BASELINE configs[1] chromosome (20 000 bins @10 kb,
5 Mb band, p=2 w=5) in the build container, one process, the ``hiccups()`` call only, and the
oracle port on the same input.  Writes profiles/r02_reference_cfg2_timing.json.  Not run on the GPU box (no reference there).

    python oracle/time_reference_cfg2.py [n_bins]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hicpeaks_b200.synth import band_pixels, synth_chromosome  # noqa: E402
from oracle import glue_oracle, hiccups_oracle as ho, ref_harness  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
band = 500
inp = synth_chromosome(n, band, 5, maxww=10, seed=17)
kw = dict(pw=[2], ww=[5], maxww=10, sig=0.1, sumq=0.01, double_fold=1.75, single_fold=2, maxapart=band * 10000, res=10000,
          use_raw=False, min_marginal_peaks=2, onlyanchor=False, min_local_reads=16)
ref = ref_harness.load_reference_callers()
M, cM = ref_harness.reference_matrices(inp)
t0 = time.perf_counter()
table = ref.hiccups(M, cM, inp["biases"], inp["biases"], inp["IR"], inp["n"], inp["Diags"], inp["cDiags"], inp["num"], "1", **kw)
t_ref = time.perf_counter() - t0
t0 = time.perf_counter()
sw, out = ho.score(inp, [2], [5], maxww=10, sig=0.1, maxapart_bins=band, min_local_reads=16)
tab2 = glue_oracle.finish_hiccups(inp, sw, out, [2], [5], 10000, 0.01, 1.75, 2, False, 2, False)
t_port = time.perf_counter() - t0
px = band_pixels(n, 5, band)
res = {"what": "unmodified reference hicpeaks.callers.hiccups() vs the oracle port, one process, build container (no GPU)",
       "workload": "cfg2 chromosome: %d bins @10kb, 5 Mb band, p=2 w=5, maxww 10, min_local_reads 16, sig 0.1 (seed 17)" % n,
       "band_pixels": px, "reference_seconds": t_ref, "reference_pixels_per_s": px / t_ref,
       "oracle_port_seconds": t_port, "oracle_port_pixels_per_s": px / t_port,
       "peaks_reference": len(table), "peaks_port": len(tab2), "same_peak_set": sorted(table) == sorted(tab2),
       "cpu": os.cpu_count()}
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_reference_cfg2_timing.json"), "w"), indent=1)
print(json.dumps(res))
