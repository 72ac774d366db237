cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
S="import json,sys; b=json.loads(sys.stdin.read()); print(b['value'], b['ms_per_step'], b['ms_per_step_host'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'])"
for m in yield spin; do
echo ${m}16; HP_SYNC=$m $B 2>/dev/null | python -c "$S"
echo ${m}8; HP_SYNC=$m taskset -c 0-7 $B 2>/dev/null | python -c "$S"
echo ${m}4; HP_SYNC=$m taskset -c 0-3 $B 2>/dev/null | python -c "$S"
done
