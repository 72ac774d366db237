"""2-rank NCCL check of the dispatcher: chromosome sharding + genome-scope histogram all-reduce on real GPUs."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from hicpeaks_b200 import dispatch
from hicpeaks_b200.synth import synth_chromosome
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
names = ["c%d" % i for i in range(5)]
sizes = {n: (600 + 100 * i, 71) for i, n in enumerate(names)}
chroms = {n: (lambda n=n, i=i: synth_chromosome(sizes[n][0], 60, 5, maxww=10, seed=50 + i, scale=60.0)) for i, n in enumerate(names)}
prm = dict(pw=[2], ww=[5], maxww=10, sig=0.1, maxapart=600000, res=10000, min_local_reads=16, min_marginal_peaks=2, onlyanchor=False)
out = {}
for scope in ("chrom", "genome"):
    multi = dispatch.GenomeRunner(comm=dispatch.TorchComm(), engine=dispatch.CudaEngine(local), fdr_scope=scope).run(chroms, sizes, **prm)
    single = dispatch.GenomeRunner(engine=dispatch.CudaEngine(local), fdr_scope=scope).run(chroms, sizes, **prm)
    same = multi == single
    out[scope] = (same, sum(len(t) for t in multi.values()))
    assert same, scope
if rank == 0:
    print("NCCL dispatcher ok:", out)
dist.destroy_process_group()
