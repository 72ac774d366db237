cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_n1.json; python scratch/show_bench.py gpurun_out/bench_n1.json | head -1
for N in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python scratch/show_bench.py gpurun_out/bench_n$N.json
done
