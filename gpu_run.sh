cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"
timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "n600 or n129 or n1000" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
