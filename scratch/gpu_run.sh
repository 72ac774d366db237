cd $GRAFT_REPO_ROOT
HICPEAKS_B200_LIB=$PWD/scratch/lib_new.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_errors.py -x -q -m gpu -k "not spec" 2>&1 | tail -2
for i in 1 2; do
HICPEAKS_B200_LIB=$PWD/scratch/lib_new.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('new', b['value'], b['kernel_ms_per_chromosome_alone'])"
HICPEAKS_B200_LIB=$PWD/scratch/lib_old.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('old', b['value'], b['kernel_ms_per_chromosome_alone'])"
done
