cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.3e e2e %.3e e2e_op %.3e frac %.4f parity %s' % (d['value'], d['e2e']['value'], d['e2e_operator']['value'], d['roofline']['frac'], d.get('parity_checked')))
print(json.dumps(d['kernel_ms_per_chromosome_alone']))
for k in ('cfg3','cfg4','cfg5'):
    print(k, json.dumps(d.get(k))[:900])
PY
