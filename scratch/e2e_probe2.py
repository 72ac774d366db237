import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from hicpeaks_b200 import _capi
from bench import make_batch, engine_arrays, WORKLOAD as W
import ctypes as C
batch = make_batch(0, 4); arrays = [engine_arrays(i) for i in batch]
ctxs = [_capi.Context(0) for _ in batch]
def up(j):
    c, inp, (Dg, cD, ir) = j
    c.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"])
jobs = list(zip(ctxs, batch, arrays))
for j in jobs: up(j)
t = time.perf_counter()
for _ in range(5):
    for j in jobs: up(j)
print('upload_counts sequential ms/chrom', (time.perf_counter() - t) / 20 * 1e3)
pool = ThreadPoolExecutor(4)
t = time.perf_counter()
for _ in range(5): list(pool.map(up, jobs))
print('upload_counts 4 threads ms/step', (time.perf_counter() - t) / 5 * 1e3)
# python-side marshalling only
def marshal(j):
    c, inp, (Dg, cD, ir) = j
    num, n = inp["num"], inp["n"]
    rp = (C.c_void_p * num)()
    for d in range(num):
        a = Dg[d]
        if a.dtype != np.int32 or not a.flags.c_contiguous or a.size != n - d: raise ValueError
        rp[d] = a.ctypes.data
t = time.perf_counter()
for _ in range(5):
    for j in jobs: marshal(j)
print('marshal ms/chrom', (time.perf_counter() - t) / 20 * 1e3)
P = _capi.Context.make_params(W["pw"], W["ww"], W["maxww"], W["sig"], W["band"], W["min_local_reads"])
for c in ctxs: c.hiccups(P)
t = time.perf_counter()
for _ in range(5): list(pool.map(lambda c: c.hiccups(P), ctxs))
print('hiccups 4 threads ms/step', (time.perf_counter() - t) / 5 * 1e3)
t = time.perf_counter()
for _ in range(5): list(pool.map(lambda c: (c.survivors(), c.gaps()), ctxs))
print('tail 4 threads ms/step', (time.perf_counter() - t) / 5 * 1e3)
