"""Join ncu per-SASS metrics with nvdisasm line info -> per source line instruction / sample counts (dev tool).
usage: ncu_lines.py rep.ncu-rep lib.so kernel_substr"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, lib, kname = sys.argv[1:4]
td = tempfile.mkdtemp(); subprocess.run(['cuobjdump','-xelf','all',os.path.abspath(lib)],cwd=td,capture_output=True)
cub=[f for f in os.listdir(td) if f.endswith('.cubin')][0]
dis=subprocess.run(['nvdisasm','-g','-c',os.path.join(td,cub)],capture_output=True,text=True).stdout.splitlines()
# collect per-instruction (in order) source line for the kernel section
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
iN=h.index('# Samples'); iI=h.index('Instructions Executed')
print('sass rows',len(data),'disasm instrs',len(lines))
agg=collections.defaultdict(lambda:[0,0,0])
for r,l in zip(data,lines):
    a=agg[l]; a[0]+=int(r[iI]); a[1]+=int(r[iN]); a[2]+=1
tot=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
src={}
for l,a in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    f,n=l if l else ('?',0)
    if f not in src:
        pth=os.path.join('hicpeaks_b200/csrc',f); src[f]=open(pth).read().splitlines() if os.path.exists(pth) else []
    text=src[f][n-1].strip()[:90] if 0<n<=len(src[f]) else ''
    print(f"{f:20s} L{n:4d} inst%={100*a[0]/tot:5.2f} samp%={100*a[1]/ts:5.2f} n={a[2]:4d} | {text}")
