"""Chromosome -> GPU dispatcher: the replacement of the reference's ``Pool(nproc).map(worker, Params)``
(/root/reference/scripts/pyHICCUPS:184-198).

Every chromosome is an independent ``hiccups()`` call in the reference (its lambda-chunks and BH
included, callers.py:263-275), so the path shards with no data-path collective: rank ``k`` of
``world`` (one process per GPU) scores the chromosomes a longest-processing-time partition gives
it and the peak tables (KBs) are gathered on the host.

``fdr_scope="genome"`` (NOT reference behaviour -- BASELINE.json's north star asks for it) merges the
(pair, background, lambda-chunk, observed) histograms over all chromosomes and ranks before BH runs: summed on the
device over a GPU's chromosomes and all-reduced over the GPUs with NCCL inside the C ABI (``hp_allreduce_hist``:
u64 sum of the histograms, max of ``E.max()``, sum of the valid counts).  Ranks are processes (``TorchComm`` only
carries the NCCL id and the finished peak tables) or the host threads of one process (``ThreadComm``, the CLI).

The communicator and the engine are small interfaces so that the sharding / merge logic is
testable without a GPU (tests/test_dispatch.py runs it over ``gloo`` with world_size 2 and a fake
engine); on a GPU box the defaults are ``TorchComm`` (NCCL) and ``CudaEngine`` (the C ABI).
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["lpt_partition", "chrom_cost", "LocalComm", "TorchComm", "ThreadComm", "CudaEngine", "GenomeRunner",
           "device_for_worker"]


def chrom_cost(n: int, num: int) -> int:
    """Work estimate of one chromosome: stored band cells."""
    return int(n) * int(num)


def lpt_partition(costs, nparts):
    """Greedy longest-processing-time partition.  Returns ``nparts`` lists of indices into ``costs``
    (each sorted by descending cost); deterministic for equal costs."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * nparts
    parts = [[] for _ in range(nparts)]
    for i in order:
        k = min(range(nparts), key=lambda j: (loads[j], j))
        parts[k].append(i)
        loads[k] += costs[i]
    return parts


def device_for_worker(n_devices: int | None = None) -> int:
    """GPU index for a forked ``Pool`` worker (identity 1..nproc) or a torchrun rank."""
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    import multiprocessing as mp
    ident = getattr(mp.current_process(), "_identity", ()) or (1,)
    if n_devices is None:
        from . import _capi
        n_devices = max(1, _capi.device_count())
    return (ident[0] - 1) % n_devices


class LocalComm:
    """world_size 1."""
    rank, world = 0, 1

    def allreduce_sum(self, a):
        return a

    def allreduce_max(self, a):
        return a

    def gather_objects(self, obj):
        return [obj]


class TorchComm:
    """torch.distributed (NCCL on GPUs, gloo on CPU) -- plumbing only."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def _reduce(self, a, op):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.dist.all_reduce(t, op=op)
        return t.cpu().numpy()

    def allreduce_sum(self, a):
        return self._reduce(a, self.dist.ReduceOp.SUM)

    def allreduce_max(self, a):
        return self._reduce(a, self.dist.ReduceOp.MAX)

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


class ThreadComm:
    """Ranks = host threads of one process, one per GPU (the CLI's ``--gpus N``).  Objects travel through shared lists and
    a barrier; the histograms never pass through here -- they are merged on the devices (``CudaEngine.merge``)."""

    class _Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def group(cls, world):
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def gather_objects(self, obj):
        self.shared.slots[self.rank] = obj
        self.shared.barrier.wait()
        out = list(self.shared.slots)
        self.shared.barrier.wait()
        return out

    def allreduce_sum(self, a):
        return sum(np.asarray(x) for x in self.gather_objects(np.asarray(a)))

    def allreduce_max(self, a):
        return np.maximum.reduce([np.asarray(x) for x in self.gather_objects(np.asarray(a))])


class CudaEngine:
    """One engine context per chromosome on one GPU (a context keeps its candidates until the FDR step).  Worker-level input
    (``weights`` in the band dict: raw count diagonals + bin weights, 1 - 4 B / pixel over PCIe) or operator-level input
    (``cDiags`` / ``IR`` / ``biases``).  The upload scratch of a scored context is given back at once (``hp_ctx_trim``)."""

    def __init__(self, device=0, max_chunks=52):
        self.device, self.max_chunks = device, max_chunks
        self._comm_ctx = None

    def score(self, name, inp, prm):
        from . import _capi, callers
        mw = min(prm["ww"])
        for max_chunks in (self.max_chunks, 64):
            ctx = _capi.Context(self.device, max_chunks)
            try:
                if inp.get("weights") is not None and inp.get("cDiags") is None:
                    ctx.upload_counts(inp["n"], inp["num"], mw, callers._as_counts(inp["Diags"], inp["num"], inp["n"]), inp["weights"])
                else:
                    raw, bal, ir = callers._as_diags(inp["Diags"], inp["cDiags"], inp["IR"], inp["n"], inp["num"], mw)
                    ctx.upload(inp["n"], inp["num"], mw, raw, bal, ir, inp["biases"], inp["biases"])
                P = ctx.make_params(prm["pw"], prm["ww"], prm["maxww"], prm["sig"], prm["maxapart"] // prm["res"],
                                    prm["min_local_reads"])
                S = ctx.score(P)
            except _capi.EngineError as e:
                ctx.close()
                if e.code == _capi.HP_ERR_EMPTY_REFIDX:      # where the reference raises (callers.py:205-208)
                    raise ValueError(str(e)) from None
                if e.code == _capi.HP_ERR_CHUNK_OVERFLOW and max_chunks < 64:
                    continue                                  # E beyond 2^17: once more with every lambda-chunk the engine has
                raise
            break
        ctx.trim()
        emax = np.array([[S.lf[i][fl].e_max for fl in (0, 1)] for i in range(len(prm["pw"]))]).ravel()
        nval = np.array([[S.lf[i][fl].n_valid for fl in (0, 1)] for i in range(len(prm["pw"]))]).ravel()
        return dict(ctx=ctx, emax=emax, nvalid=nval, chromLen=inp["n"], summary=S)

    def hist(self, h):
        return h["ctx"].hist_export()

    def merge(self, handles, npw, nranks=1, rank=0, unique_id=None):
        """Genome-wide merge on the device: sums over this GPU's contexts, one NCCL all-reduce over the ranks
        (``hp_allreduce_hist``).  Every rank calls it.  Returns the device milliseconds."""
        from . import _capi
        if self._comm_ctx is None:
            self._comm_ctx = _capi.Context(self.device, self.max_chunks if not handles else handles[0]["ctx"].max_chunks)
            self._comm_ctx.comm_init(nranks, rank, unique_id)
        ms = self._comm_ctx.allreduce_hist([h["ctx"] for h in handles], npw)
        for h in handles:
            S = h["ctx"].summary()
            npw = len(h["emax"]) // 2
            h["emax"] = np.array([[S.lf[i][fl].e_max for fl in (0, 1)] for i in range(npw)]).ravel()
            h["nvalid"] = np.array([[S.lf[i][fl].n_valid for fl in (0, 1)] for i in range(npw)]).ravel()
        return ms

    def close(self):
        if self._comm_ctx is not None:
            self._comm_ctx.close()
            self._comm_ctx = None

    def finish(self, h, prm, hist=None, numbin=None):
        from . import callers
        ctx = h["ctx"]
        if hist is not None:
            ctx.hist_import(hist)
        if numbin is None:
            numbin = [callers._numpy_numbin(e, n) for e, n in zip(h["emax"], h["nvalid"])]
        if max(numbin) > ctx.max_chunks:
            raise RuntimeError("lambda-chunk count %d exceeds the context's max_chunks" % max(numbin))
        ctx.fdr(np.asarray(numbin, dtype=np.int32))
        table = callers.assemble_table(ctx.survivors(), ctx.gaps(), h["chromLen"], list(prm["pw"]), list(prm["ww"]),
                                       prm["res"], prm["sumq"], prm["double_fold"], prm["single_fold"], prm["use_raw"],
                                       prm["min_marginal_peaks"], prm["onlyanchor"])
        ctx.close()
        return table


DEFAULTS = dict(pw=[2], ww=[5], maxww=20, sig=0.1, sumq=0.01, double_fold=1.75, single_fold=2, maxapart=2000000,
                res=10000, use_raw=False, min_marginal_peaks=3, onlyanchor=True, min_local_reads=25)


class GenomeRunner:
    """Scores a set of chromosomes across ``comm.world`` ranks.

    ``chroms``: ordered mapping ``name -> loader`` where ``loader()`` returns the band input of that
    chromosome (keys n, num, Diags, cDiags, IR, biases) -- only the owning rank calls it.  ``sizes``:
    ``name -> (n, num)`` for the partition.  Returns ``{name: pixel_table}`` on every rank."""

    def __init__(self, comm=None, engine=None, fdr_scope="chrom"):
        if fdr_scope not in ("chrom", "genome"):
            raise ValueError("fdr_scope must be 'chrom' or 'genome'")
        self.comm = comm or LocalComm()
        self.engine = engine
        self.fdr_scope = fdr_scope
        self.merge_ms = None

    def assignment(self, sizes):
        names = list(sizes)
        parts = lpt_partition([chrom_cost(*sizes[n]) for n in names], self.comm.world)
        return [[names[i] for i in p] for p in parts]

    def run(self, chroms, sizes, **params):
        prm = dict(DEFAULTS)
        prm.update(params)
        engine = self.engine or CudaEngine(device_for_worker())
        mine = self.assignment(sizes)[self.comm.rank]
        handles = {name: engine.score(name, chroms[name](), prm) for name in mine}
        tables = {}
        if self.fdr_scope == "chrom":
            for name in mine:
                tables[name] = engine.finish(handles[name], prm)
        elif hasattr(engine, "merge"):
            # the one data-path collective, on the devices: hp_allreduce_hist (NCCL when there is more than one rank)
            uid = None
            if self.comm.world > 1:
                from . import _capi
                uid = self.comm.gather_objects(_capi.comm_unique_id() if self.comm.rank == 0 else None)[0]
            self.merge_ms = engine.merge([handles[name] for name in mine], len(prm["pw"]), self.comm.world, self.comm.rank, uid)
            for name in mine:
                tables[name] = engine.finish(handles[name], prm)
        else:
            # engines without a device-side merge (the CPU stand-ins of tests/test_dispatch.py): host arrays through the communicator
            nlf = 2 * len(prm["pw"])
            total = None
            emax = np.zeros(nlf)
            nval = np.zeros(nlf, dtype=np.int64)
            for name in mine:
                h = engine.hist(handles[name])
                total = h.copy() if total is None else total + h
                emax = np.maximum(emax, handles[name]["emax"])
                nval += np.asarray(handles[name]["nvalid"], dtype=np.int64)
            shape = self.comm.allreduce_max(np.array(total.shape if total is not None else (0, 0), dtype=np.int64))
            if total is None:
                total = np.zeros(tuple(int(x) for x in shape), dtype=np.int64)
            total = self.comm.allreduce_sum(total)
            emax = self.comm.allreduce_max(emax)
            nval = self.comm.allreduce_sum(nval)
            from .callers import _numpy_numbin
            numbin = [_numpy_numbin(e, n) for e, n in zip(emax, nval)]
            for name in mine:
                tables[name] = engine.finish(handles[name], prm, hist=total, numbin=numbin)
        merged = {}
        for part in self.comm.gather_objects(tables):
            merged.update(part)
        return {name: merged[name] for name in chroms if name in merged}
