cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for c in 4 8 12 16; do echo "chroms=$c"; $B --chroms $c 2>/dev/null > gpurun_out/b.json; python scratch/show_bench.py gpurun_out/b.json 2>/dev/null | head -1; done
