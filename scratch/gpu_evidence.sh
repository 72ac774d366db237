cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scratch/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
timeout 600 python scratch/host_bw_probe.py > gpurun_out/host_bw_probe.json 2> gpurun_out/host_bw_probe.err; cat gpurun_out/host_bw_probe.json | head -c 1500; echo
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_fast -s 1 -c 1 -o gpurun_out/r02_k_score_fast python scratch/prof_fast.py 20000 2 > gpurun_out/r02_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_under_ncu.json 2>/dev/null
