cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print('value %.3e ms/step %.3f e2e %.3e e2e_op %.3e frac %.4f parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_operator']['value'], d['roofline']['frac'], d.get('parity_checked')))
print(json.dumps(d['kernel_ms_per_chromosome_alone']))
for k in ('cfg3','cfg4','cfg5'):
    print(k, json.dumps(d.get(k))[:1200])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; head -c 400 gpurun_out/bench_ref.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_fast -s 1 -c 1 -f -o gpurun_out/r02c_k_score_fast python scratch/prof_fast.py 20000 2 > gpurun_out/r02c_prof.log 2>&1; tail -1 gpurun_out/r02c_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02c_bench_under_ncu.json 2>/dev/null; wc -l gpurun_out/r02c_launches.csv
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scratch/sanitize.py > gpurun_out/r02c_sanitize_memcheck.log 2>&1; tail -2 gpurun_out/r02c_sanitize_memcheck.log
