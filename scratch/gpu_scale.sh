cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit code $?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('N', d['n_gpus'], 'value %.3e e2e %.3e e2e_op %.3e frac %.4f' % (d['value'], d['e2e']['value'], d['e2e_operator']['value'], d['roofline']['frac']))
print('by rank', d.get('ms_per_step_by_rank'))
for k in ('cfg3','cfg4'):
    print(k, json.dumps(d.get(k))[:1100])
PY
