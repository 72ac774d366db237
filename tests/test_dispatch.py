"""CPU: chromosome sharding and the genome-scope histogram merge of hicpeaks_b200.dispatch, over gloo with
world_size 2 and a fake engine (the CUDA engine itself is covered by the -m gpu tests)."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from hicpeaks_b200 import dispatch
from hicpeaks_b200.synth import hg38_autosome_bins

NLF, BINS = 2, 50


class FakeEngine:
    """Deterministic stand-in: histogram and E.max() derived from the chromosome name."""

    def __init__(self):
        self.scored = []

    def score(self, name, inp, prm):
        self.scored.append(name)
        rng = np.random.default_rng(abs(hash_name(name)))
        return dict(name=name, hist=rng.integers(0, 1000, size=(NLF, BINS)).astype(np.int64),
                    emax=rng.uniform(1, 200, size=NLF), nvalid=rng.integers(1, 10 ** 6, size=NLF), n=inp["n"])

    def hist(self, h):
        return h["hist"]

    def finish(self, h, prm, hist=None, numbin=None):
        used = h["hist"] if hist is None else hist
        return {(h["n"], 0): (h["name"], int(used.sum()), None if numbin is None else tuple(int(x) for x in numbin))}


def hash_name(name):
    return sum((i + 1) * ord(c) for i, c in enumerate(name))


def make_genome(k=7):
    sizes = {"chr%d" % (i + 1): (1000 + 137 * i, 91) for i in range(k)}
    chroms = {name: (lambda n=n: dict(n=n)) for name, (n, _) in sizes.items()}
    return chroms, sizes


def test_lpt_partition_balances_hg38():
    bins = hg38_autosome_bins(10000)
    costs = [dispatch.chrom_cost(n, 511) for n in bins]
    for parts in (2, 4, 8):
        p = dispatch.lpt_partition(costs, parts)
        assert sorted(i for q in p for i in q) == list(range(len(costs)))
        loads = [sum(costs[i] for i in q) for q in p]
        assert max(loads) / (sum(loads) / parts) < 1.06          # 8-way imbalance 1.043


def test_single_rank_matches_plain_loop():
    chroms, sizes = make_genome()
    eng = FakeEngine()
    out = dispatch.GenomeRunner(engine=eng).run(chroms, sizes, pw=[2], ww=[5])
    assert list(out) == list(chroms)
    assert sorted(eng.scored) == sorted(chroms)
    for name, tab in out.items():
        (key, val), = tab.items()
        assert val[0] == name and val[2] is None


def _worker(rank, world, port, scope, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chroms, sizes = make_genome()
        eng = FakeEngine()
        out = dispatch.GenomeRunner(comm=dispatch.TorchComm(), engine=eng, fdr_scope=scope).run(chroms, sizes, pw=[2], ww=[5])
        q.put((rank, sorted(eng.scored), out))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("scope", ["chrom", "genome"])
def test_two_ranks_gloo(scope):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scope, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    chroms, sizes = make_genome()
    # every chromosome scored exactly once, by the rank the partition names
    assign = dispatch.GenomeRunner().assignment(sizes)
    assert sorted(res[0][1] + res[1][1]) == sorted(chroms)
    parts2 = dispatch.lpt_partition([dispatch.chrom_cost(*sizes[n]) for n in sizes], 2)
    names = list(sizes)
    assert res[0][1] == sorted(names[i] for i in parts2[0]) and res[1][1] == sorted(names[i] for i in parts2[1])
    assert len(assign) == 1
    # both ranks hold the same, complete result
    assert res[0][2] == res[1][2] and list(res[0][2]) == list(chroms)
    single = dispatch.GenomeRunner(engine=FakeEngine(), fdr_scope=scope).run(chroms, sizes, pw=[2], ww=[5])
    assert res[0][2] == single                      # sharding does not change the answer
    if scope == "genome":
        ref = FakeEngine()
        total = sum(ref.score(n, dict(n=sizes[n][0]), None)["hist"].sum() for n in chroms)
        for tab in res[0][2].values():
            (_, val), = tab.items()
            assert val[1] == total and val[2] is not None
