"""cfg5: APA 41x41 pileup over 50 000 synthetic anchors @10 kb on one B200, next to the oracle (numpy, 1 core) on a sample."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from hicpeaks_b200 import apa
from hicpeaks_b200.synth import synth_chromosome
from oracle import apa_oracle as ao
from test_gpu_apa import BandMatrix
n, band, w = 20000, 540, 20
inp = synth_chromosome(n, band, 1, maxww=0, seed=5)
diags = ao.balanced_diags(inp["Diags"], inp["weights"])
rng = np.random.default_rng(1)
i = rng.integers(w, n - band - w, 50000); j = i + rng.integers(30, band - 40, 50000)
pos = list(zip(i.tolist(), j.tolist()))
M = BandMatrix(diags, n)
wins = apa.apa_submatrix(M, pos[:100], w=w)          # warm-up (context, kernels)
t = time.perf_counter(); wins = apa.apa_submatrix(M, pos, w=w); t1 = time.perf_counter()
avg, score, z, p, maxi = apa.apa_analysis(wins, w=w, cw=3); t2 = time.perf_counter()
print("GPU: submatrix %.1f ms (incl. %.0f MB band upload), analysis %.1f ms, windows %d, score %.4f" % (1e3 * (t1 - t), sum(d.nbytes for d in diags[:band]) / 1e6, 1e3 * (t2 - t1), len(wins), score))
t = time.perf_counter(); ex, valid = ao.apa_submatrix(diags, n, pos[:2000], w=w); t1 = time.perf_counter()
print("oracle: submatrix %.1f ms for 2000 anchors -> %.1f s per 50 000" % (1e3 * (t1 - t), 25 * (t1 - t)))
