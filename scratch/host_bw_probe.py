"""Host-side ceiling of the worker-level boundary (e2e): how fast can this box's cores READ int32 count diagonals from
pageable memory?  Two loops over the same 324 MB working set per process (8 cfg2 chromosomes' worth of int32 counts):
  copy   numpy memcpy into a second buffer (read 4 B + write 4 B per count)
  pack   the library's own narrowing (hp_narrow_diagonal: read 4 B, write 1 B per count) -- what hp_band_upload_counts does
with T threads per process, for 1 process and for 8 processes at once (the 8-GPU bench: 8 ranks on one box).
Writes one JSON line; run on the GPU box:  python scratch/host_bw_probe.py > gpurun_out/host_bw_probe.json"""
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(nthreads, seconds, q):
    from hicpeaks_b200 import _capi
    lib = _capi.load_library()
    n, ndiag = 20000, 128
    rng = np.random.default_rng(os.getpid())
    src = [rng.poisson(3.0, n).astype(np.int32) for _ in range(ndiag * nthreads)]       # 10 MB per thread, > L2 per core
    dst = [np.empty(n * 4, dtype=np.uint8) for _ in range(nthreads)]
    dst32 = [np.empty(n, dtype=np.int32) for _ in range(nthreads)]
    out = {}
    for mode in ("copy", "pack"):
        counts = [0] * nthreads
        stop = time.perf_counter() + seconds

        def loop(t):
            es = C.c_int32()
            k = 0
            while time.perf_counter() < stop:
                for a in src[t * ndiag:(t + 1) * ndiag]:
                    if mode == "copy":
                        np.copyto(dst32[t], a)
                    else:
                        lib.hp_narrow_diagonal(a.ctypes.data_as(C.c_void_p), a.size, dst[t].ctypes.data_as(C.c_void_p), C.byref(es))
                    k += a.nbytes
            counts[t] = k

        th = [threading.Thread(target=loop, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        out[mode] = sum(counts) / (time.perf_counter() - t0) / 1e9
    q.put(out)


def run(nproc, nthreads, seconds=2.0):
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(nthreads, seconds, q)) for _ in range(nproc)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    return {m: sum(r[m] for r in res) for m in ("copy", "pack")}


if __name__ == "__main__":
    cores = os.cpu_count() or 1
    rows = []
    for nproc, nth in ((1, 1), (1, 4), (1, 8), (1, min(16, cores)), (1, cores), (8, max(1, cores // 8)), (8, max(1, cores // 4))):
        r = run(nproc, nth)
        rows.append({"processes": nproc, "threads_per_process": nth, "int32_read_GBps_copy": round(r["copy"], 2),
                     "int32_read_GBps_pack": round(r["pack"], 2)})
    print(json.dumps({"host_cores": cores, "what": "aggregate GB/s of int32 counts READ from pageable memory (copy: +4 B written per 4 B read; "
                      "pack: hp_narrow_diagonal, +1 B written)", "rows": rows}))
