cd $GRAFT_REPO_ROOT
for c in 4 6 8; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chroms $c 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('chroms', $c, 'value', b['value'], 'ms/step', b['ms_per_step'], 'e2e', b['e2e']['value'], 'e2e_op', b['e2e_operator']['value'])"
done
