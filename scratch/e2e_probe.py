import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device='cuda')
for _ in range(3): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print('pinned H2D GB/s', 5 * 0.268435456 / (time.perf_counter() - t))
z = torch.empty(256 << 20, dtype=torch.uint8)
t = time.perf_counter(); y.copy_(z); torch.cuda.synchronize(); print('pageable H2D GB/s', 0.268435456 / (time.perf_counter() - t))
a = np.zeros(256 << 20, dtype=np.uint8); b = np.empty_like(a)
t = time.perf_counter(); b[:] = a; print('host memcpy 1 thread GB/s', 0.268435456 / (time.perf_counter() - t))
from hicpeaks_b200 import _capi
from bench import make_batch, engine_arrays, WORKLOAD as W
batch = make_batch(0, 4); arrays = [engine_arrays(i) for i in batch]
ctxs = [_capi.Context(0) for _ in batch]
def up(j):
    c, inp, (Dg, cD, ir) = j
    c.upload(inp["n"], inp["num"], inp["min_ww"], Dg, cD, ir, inp["biases"], inp["biases"])
jobs = list(zip(ctxs, batch, arrays))
for j in jobs: up(j)
t = time.perf_counter()
for _ in range(3):
    for j in jobs: up(j)
print('upload sequential ms/chrom', (time.perf_counter() - t) / 12 * 1e3)
pool = ThreadPoolExecutor(4)
t = time.perf_counter()
for _ in range(3): list(pool.map(up, jobs))
print('upload 4 threads ms/step', (time.perf_counter() - t) / 3 * 1e3)
P = _capi.Context.make_params(W["pw"], W["ww"], W["maxww"], W["sig"], W["band"], W["min_local_reads"])
def sc(c): return c.hiccups(P)
t = time.perf_counter()
for _ in range(3): list(pool.map(sc, ctxs))
print('hiccups 4 threads ms/step', (time.perf_counter() - t) / 3 * 1e3)
def tail(c): return c.survivors().nbytes + c.gaps().size
t = time.perf_counter()
for _ in range(3): list(pool.map(tail, ctxs))
print('survivors+gaps 4 threads ms/step', (time.perf_counter() - t) / 3 * 1e3)
