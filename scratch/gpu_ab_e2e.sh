cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in A B; do
HICPEAKS_B200_LIB=$GRAFT_REPO_ROOT/scratch/ab/lib_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('variant $v value %.3e e2e %.3e (%.2f ms/step) e2e_op %.3e' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_operator']['value']))"
done
done
