cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "not (1-2-4 or n700 or n300 or p1w3 or union or p4w7)" 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('new', b['value'], b['kernel_ms_per_chromosome_alone'])"
HICPEAKS_B200_LIB=$PWD/scratch/lib_old.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('old', b['value'], b['kernel_ms_per_chromosome_alone'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_spec -s 2 -c 1 -o gpurun_out/prof_score -f python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
