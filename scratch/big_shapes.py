"""Size smoke for the BASELINE config shapes that are not the bench line: cfg3 (hg38 chr1 @10 kb, union program) and cfg4
(5 kb, 10 Mb band, (4,7)): spec vs generic kernel agreement + timings."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from hicpeaks_b200 import _capi
from hicpeaks_b200.synth import synth_chromosome, band_pixels
from test_gpu_fullsize import _run, _same
import test_gpu_fullsize as T
for name, n, band, pw, ww, scale in (("cfg3 chr1@10kb union", 24896, 500, [1, 2, 4], [3, 5, 7], 300.0),
                                     ("cfg4 chr21@5kb (4,7) band 2000", 9342, 2000, [4], [7], 300.0),
                                     ("cfg4 chr8@5kb (4,7) band 2000", 29028, 2000, [4], [7], 300.0)):
    t = time.perf_counter()
    inp = synth_chromosome(n, band, min(ww), maxww=10, seed=3, scale=scale)
    tg = time.perf_counter() - t
    T.BAND = band
    with _capi.Context(0) as c1, _capi.Context(0) as c2:
        try:
            a = _run(c1, inp, pw, ww)
            b = _run(c2, inp, pw, ww, generic=True)
            _same(a, b)
            a = _run(c1, inp, pw, ww)
            S = a[0]
            print("%s: n=%d pixels=%d spec=%d frozen=%d survivors=%d | ms levels %.3f score %.3f fdr %.3f (generic score %.3f) | gen %.1fs" % (
                name, n, S.band_pixels, S.spec_kernel, S.frozen_w, S.n_survivors, S.ms_levels, S.ms_score, S.ms_fdr, b[0].ms_score, tg))
        except Exception as e:
            print(name, "FAILED:", type(e).__name__, e)
