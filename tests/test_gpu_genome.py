"""GPU: genome-wide FDR -- the lambda-chunk histograms of several chromosomes merged on the device
(``hp_allreduce_hist``) and, with two GPUs, all-reduced with NCCL inside the C ABI.  The merged run must equal the host
merge (``hp_hist_export`` -> numpy sum -> ``hp_hist_import``) and, across ranks, the single-GPU result."""
import threading

import numpy as np
import pytest

from hicpeaks_b200 import _capi, dispatch
from hicpeaks_b200.synth import synth_chromosome

pytestmark = pytest.mark.gpu

PRM = dict(pw=[1, 2, 4], ww=[3, 5, 7], maxww=10, sig=0.1, maxapart=300 * 10000, res=10000, min_local_reads=16,
           min_marginal_peaks=2, onlyanchor=False)


def _chroms(k=5):
    out = {}
    for i in range(k):
        inp = synth_chromosome(1500 + 211 * i, 300, 3, maxww=10, seed=70 + i)
        out["c%d" % i] = dict(n=inp["n"], num=inp["num"], Diags=inp["Diags"], weights=inp["weights"])
    return out


def _survivor_bytes(ctx):
    sv = ctx.survivors()
    return sv[np.lexsort((sv["pair"], sv["c"], sv["r"]))].tobytes()


def test_device_merge_equals_host_merge():
    chroms = _chroms(3)
    prm = dict(dispatch.DEFAULTS)
    prm.update(PRM)
    eng = dispatch.CudaEngine(0)
    # (a) host merge: export, numpy sum, import
    ha = [eng.score(k, v, prm) for k, v in chroms.items()]
    total = sum(eng.hist(h) for h in ha)
    emax = np.maximum.reduce([h["emax"] for h in ha])
    nval = sum(h["nvalid"] for h in ha)
    from hicpeaks_b200.callers import _numpy_numbin
    numbin = [_numpy_numbin(e, n) for e, n in zip(emax, nval)]
    ref = []
    for h in ha:
        h["ctx"].hist_import(total)
        h["ctx"].fdr(np.asarray(numbin, dtype=np.int32))
        ref.append((_survivor_bytes(h["ctx"]), [h["ctx"].chunk_table(pi, fl)[3].tobytes() for pi in range(3) for fl in (0, 1)]))
        h["ctx"].close()
    # (b) device merge through the C ABI
    hb = [eng.score(k, v, prm) for k, v in chroms.items()]
    ms = eng.merge(hb, 3)
    assert ms >= 0
    for h, (sv_ref, tabs_ref) in zip(hb, ref):
        assert np.array_equal(h["emax"], emax) and np.array_equal(h["nvalid"], nval)
        nb = [_numpy_numbin(e, n) for e, n in zip(h["emax"], h["nvalid"])]
        assert nb == numbin
        h["ctx"].fdr(np.asarray(nb, dtype=np.int32))
        assert _survivor_bytes(h["ctx"]) == sv_ref
        assert [h["ctx"].chunk_table(pi, fl)[3].tobytes() for pi in range(3) for fl in (0, 1)] == tabs_ref
        h["ctx"].close()
    eng.close()
    assert len(ref[0][0]) > 0


def _genome_tables(ngpu, chroms):
    sizes = {k: (v["n"], v["num"]) for k, v in chroms.items()}
    comms = dispatch.ThreadComm.group(ngpu) if ngpu > 1 else [dispatch.LocalComm()]
    results, errors, ms = [None] * ngpu, [], [None] * ngpu

    def rank_main(g):
        eng = dispatch.CudaEngine(g)
        try:
            runner = dispatch.GenomeRunner(comm=comms[g], engine=eng, fdr_scope="genome")
            results[g] = runner.run({k: (lambda k=k: chroms[k]) for k in chroms}, sizes, **PRM)
            ms[g] = runner.merge_ms
        except BaseException as e:                      # noqa: BLE001
            errors.append(e)
            if ngpu > 1:
                comms[g].shared.barrier.abort()
        finally:
            eng.close()

    th = [threading.Thread(target=rank_main, args=(g,)) for g in range(ngpu)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise errors[0]
    return results, ms


def test_genome_scope_across_gpus_equals_one_gpu():
    """NCCL all-reduce of the histograms between the GPUs of the box (one host thread = one rank per GPU)."""
    ngpu = _capi.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    ngpu = min(ngpu, 4)
    chroms = _chroms(6)
    one, _ = _genome_tables(1, chroms)
    many, ms = _genome_tables(ngpu, chroms)
    assert all(r == one[0] for r in many), "multi-GPU genome-scope peak tables differ from the single-GPU run"
    assert sum(len(t) for t in one[0].values()) > 0
    print("genome-scope merge over %d GPUs: %s ms (device, per rank)" % (ngpu, ["%.3f" % m for m in ms]))
