cd $GRAFT_REPO_ROOT
N=$1
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python -c "import json,sys; b=json.loads(open('gpurun_out/bench_n$N.json').read()); print('N', b['n_gpus'], 'value', b['value'], 'ms/step', b['ms_per_step'], b['ms_per_step_host'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], 'e2e_op', b['e2e_operator']['value'], b['clocks'])" || cat gpurun_out/bench_n$N.json
