"""Per code region (source-line ranges of one file) instruction counts and stall-reason samples from an ncu report.
usage: ncu_regions.py rep lib kernel_substr file 'name:lo-hi,name:lo-hi,...'"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, lib, kname, fname, spec = sys.argv[1:6]
regions = []
for part in spec.split(','):
    nm, rg = part.split(':'); lo, hi = rg.split('-'); regions.append((nm, int(lo), int(hi)))
td = tempfile.mkdtemp(); subprocess.run(['cuobjdump','-xelf','all',os.path.abspath(lib)],cwd=td,capture_output=True)
cub=[f for f in os.listdir(td) if f.endswith('.cubin')][0]
dis=subprocess.run(['nvdisasm','-g','-c',os.path.join(td,cub)],capture_output=True,text=True).stdout.splitlines()
insec=False; cur=None; lines=[]
for ln in dis:
    if ln.startswith('//---') and '.text.' in ln: insec = kname in ln; continue
    if not insec: continue
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+',ln): lines.append(cur)
sass=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(sass.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hi[0]]; end=hi[1]-1 if len(hi)>1 else len(rows); data=rows[hi[0]+1:end]
iI=h.index('Instructions Executed'); iN=h.index('# Samples')
stalls=[c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
idx={c:h.index(c) for c in stalls}
agg=collections.defaultdict(lambda: collections.Counter())
for r,l in zip(data,lines):
    key='other'
    if l and l[0]==fname:
        for nm,lo,hi_ in regions:
            if lo<=l[1]<=hi_: key=nm; break
    elif l: key='inc:'+l[0]
    a=agg[key]; a['inst']+=int(r[iI]); a['samp']+=int(r[iN]); a['sass']+=1
    for c in stalls: a[c]+=int(r[idx[c]] or 0)
ti=sum(a['inst'] for a in agg.values()); ts=sum(a['samp'] for a in agg.values())
print('total warp inst %d samples %d'%(ti,ts))
for k,a in sorted(agg.items(), key=lambda x:-x[1]['samp']):
    top=sorted(((a[c],c[6:]) for c in stalls), reverse=True)[:5]
    print('%-16s inst%%=%5.1f samp%%=%5.1f sass=%4d | %s'%(k,100*a['inst']/ti,100*a['samp']/ts,a['sass'],' '.join('%s=%.0f%%'%(n,100*v/max(1,a['samp'])) for v,n in top)))
