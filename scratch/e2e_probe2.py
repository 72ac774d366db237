import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from hicpeaks_b200 import _capi
from bench import make_batch, engine_arrays, WORKLOAD as W
NC = int(sys.argv[1]) if len(sys.argv) > 1 else 8
batch = make_batch(0, NC); arrays = [engine_arrays(i) for i in batch]
ctxs = [_capi.Context(0) for _ in batch]
def up(j):
    c, inp, (Dg, cD, ir) = j
    c.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"])
jobs = list(zip(ctxs, batch, arrays))
for j in jobs: up(j)
print('upload bytes', ctxs[0].upload_bytes(), 'of', sum(a.nbytes for a in arrays[0][0]))
t = time.perf_counter()
for _ in range(5):
    for j in jobs: up(j)
print('upload_counts sequential ms/chrom', (time.perf_counter() - t) / (5 * NC) * 1e3)
pool = ThreadPoolExecutor(NC)
t = time.perf_counter()
for _ in range(5): list(pool.map(up, jobs))
print('upload_counts %d threads ms/step' % NC, (time.perf_counter() - t) / 5 * 1e3)
P = _capi.Context.make_params(W["pw"], W["ww"], W["maxww"], W["sig"], W["band"], W["min_local_reads"])
for c in ctxs: c.hiccups(P)
t = time.perf_counter()
for _ in range(5): list(pool.map(lambda c: c.hiccups(P), ctxs))
print('hiccups %d threads ms/step' % NC, (time.perf_counter() - t) / 5 * 1e3)
t = time.perf_counter()
for _ in range(5): list(pool.map(lambda c: (c.survivors(), c.gaps()), ctxs))
print('tail %d threads ms/step' % NC, (time.perf_counter() - t) / 5 * 1e3)
def full(j):
    up(j); c = j[0]; c.hiccups(P); c.survivors(); c.gaps()
t = time.perf_counter()
for _ in range(5): list(pool.map(full, jobs))
print('full %d threads ms/step' % NC, (time.perf_counter() - t) / 5 * 1e3)
for pt in (1, 2, 4):
    os.environ["HP_PACK_THREADS_PROBE"] = str(pt)
