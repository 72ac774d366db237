cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_golden.py tests/test_gpu_genome.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 120 python scratch/prof_fast.py 20000 4 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_quick.csv python scratch/prof_fast.py 20000 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_quick.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; data=rows[hi+2:]
kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in data:
    if len(r)>mv: agg[r[kn][:50]].append(float(r[mv].replace(',','')))
for k,v in sorted(agg.items(), key=lambda x:-min(x[1])):
    print('%-52s n=%3d min=%8.1f us'%(k,len(v),min(v)/1000))
PY
