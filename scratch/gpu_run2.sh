cd $GRAFT_REPO_ROOT
N=$1
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${2:-5} --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python scratch/show_bench.py gpurun_out/bench_n$N.json || cat gpurun_out/bench_n$N.json
