cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in A B; do
echo "variant $v"; HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_$v.so timeout 60 python scratch/prof_fast.py 20000 4 2>&1 | tail -2
done
done
HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_B.so timeout 60 python scratch/prof_fast.py 20000 3 union 2>&1 | tail -1
HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_A.so timeout 60 python scratch/prof_fast.py 20000 3 union 2>&1 | tail -1
HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_B.so timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_golden.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
