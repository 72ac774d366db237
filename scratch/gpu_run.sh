cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench.json; python scratch/show_bench.py gpurun_out/bench.json; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench.json; python scratch/show_bench.py gpurun_out/bench.json
