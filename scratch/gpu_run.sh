cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_spec -s 2 -c 1 -o gpurun_out/prof_score -f python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prep_band -s 2 -c 1 -o gpurun_out/prof_prep -f python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
cat gpurun_out/bench.json
