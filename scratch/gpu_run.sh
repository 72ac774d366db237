cd $GRAFT_REPO_ROOT
for v in fused3 split4; do
HICPEAKS_B200_LIB=$PWD/scratch/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -x -q -m gpu -k "not (1-2-4 or n700 or n300 or union or p1w3 or p4w7 or chr21)" 2>&1 | tail -1
HICPEAKS_B200_LIB=$PWD/scratch/lib_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('$v', b['value'], b['kernel_ms_per_chromosome_alone'])"
HICPEAKS_B200_LIB=$PWD/scratch/lib_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_$v.csv python bench.py --steps 1 --warmup 3 --chroms 2 --no-cpu-baseline > /dev/null 2>&1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('base', b['value'], b['kernel_ms_per_chromosome_alone'])"
