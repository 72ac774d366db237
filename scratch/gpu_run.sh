cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scratch/e2e_probe2.py 8
HP_PACK_THREADS=2 timeout 300 python scratch/e2e_probe2.py 8 | grep -v hiccups
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
