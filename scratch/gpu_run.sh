cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for m in yield hybrid block; do
for c in "" "taskset -c 0-3"; do
echo "$m [$c]"; HP_SYNC=$m LOCAL_WORLD_SIZE=$([ -z "$c" ] && echo 1 || echo 4) $c $B 2>/dev/null > gpurun_out/b.json; python scratch/show_bench.py gpurun_out/b.json | head -1
done; done
