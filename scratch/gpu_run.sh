cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for c in "" "taskset -c 0-3"; do $c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench.json; python scratch/show_bench.py gpurun_out/bench.json; 
python -c "import json; print(json.load(open('gpurun_out/bench.json'))['kernel_ms_per_chromosome_alone'])"; done
