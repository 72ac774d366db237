import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from hicpeaks_b200 import _capi
from bench import make_batch, engine_arrays
NC = 8
batch = make_batch(0, NC); arrays = [engine_arrays(i) for i in batch]
ctxs = [_capi.Context(0) for _ in batch]
def up(j):
    c, inp, (Dg, cD, ir) = j
    t = time.perf_counter()
    c.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"])
    return (time.perf_counter() - t) * 1e3
jobs = list(zip(ctxs, batch, arrays))
for j in jobs: up(j)
print("--- sequential", file=sys.stderr)
for j in jobs[:3]: print("call ms %.3f" % up(j), file=sys.stderr)
print("--- 8 threads", file=sys.stderr)
pool = ThreadPoolExecutor(NC)
t = time.perf_counter(); r = list(pool.map(up, jobs)); print("step ms %.3f" % ((time.perf_counter() - t) * 1e3), ["%.2f" % x for x in r], file=sys.stderr)
