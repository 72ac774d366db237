cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench.json; python scratch/show_bench.py gpurun_out/bench.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_e2e.csv python scratch/e2e_probe2.py 2 > /dev/null 2>&1
