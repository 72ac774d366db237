cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in A B; do
echo "variant $v"; HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_$v.so timeout 60 python scratch/prof_fast.py 20000 5 2>&1 | tail -2
done
done
HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_B.so timeout 300 python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -2
