import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from hicpeaks_b200 import _capi
from bench import make_batch, engine_arrays, WORKLOAD as W
NC = 8
batch = make_batch(0, NC); arrays = [engine_arrays(i) for i in batch]
ctxs = [_capi.Context(0) for _ in batch]
P = _capi.Context.make_params(W["pw"], W["ww"], W["maxww"], W["sig"], W["band"], W["min_local_reads"])
for c, inp, (Dg, cD, ir) in zip(ctxs, batch, arrays):
    c.upload_counts(inp["n"], inp["num"], inp["min_ww"], Dg, inp["weights"]); c.hiccups(P)
c = ctxs[0]
for name, fn in (("survivors", c.survivors), ("gaps", c.gaps)):
    fn()
    t = time.perf_counter()
    for _ in range(20): fn()
    print(name, "single thread ms/call", (time.perf_counter() - t) / 20 * 1e3)
pool = ThreadPoolExecutor(NC)
for name in ("survivors", "gaps"):
    t = time.perf_counter()
    for _ in range(10): list(pool.map(lambda c: getattr(c, name)(), ctxs))
    print(name, "8 threads ms/step", (time.perf_counter() - t) / 10 * 1e3)
for mode in ("spin", "yield", "block"):
    pass
