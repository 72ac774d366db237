cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "n600 or n500_b60 or n129 or n1000 or n777 or poisson" 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_spec -s 2 -c 1 -o gpurun_out/prof_score -f python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
