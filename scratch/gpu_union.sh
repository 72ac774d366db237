cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python scratch/prof_fast.py 20000 3 union 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_union.csv python scratch/prof_fast.py 20000 2 union > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_union.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; data=rows[hi+2:]
kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value')
for r in data:
    if len(r)>mv and ('score' in r[kn] or 'levels' in r[kn] or 'exact' in r[kn]): print(r[kn][:60], r[mn], r[mv])
PY
