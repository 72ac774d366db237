"""Summarise an .ncu-rep: key raw metrics + per-region SASS profile (dev tool)."""
import csv, re, collections, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum']
for d in data[:1]:
    print(d[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('   ',w, d[hdr.index(w)], units[hdr.index(w)])
    for i,h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active.ratio' in h and 'not_issued' not in h and float(d[i] or 0)>0.05: print('    stall', h.split('stalled_')[1].split('_per')[0], d[i])
sass = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(sass.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hi[0]]; end=hi[1]-1 if len(hi)>1 else len(rows)
data=rows[hi[0]+1:end]
iS=h.index('Source'); iN=h.index('# Samples'); iI=h.index('Instructions Executed'); iT=h.index('Avg. Threads Executed')
tot_i=sum(int(r[iI]) for r in data if r[iI].isdigit()); tot_s=sum(int(r[iN]) for r in data if r[iN].isdigit())
print('sass rows',len(data),'inst',tot_i,'samples',tot_s)
for k in range(0,len(data),B):
    seg=data[k:k+B]
    ins=sum(int(r[iI]) for r in seg); sm=sum(int(r[iN]) for r in seg)
    ops=collections.Counter(re.sub(r'^@!?U?P\d+\s+','',r[iS]).split()[0].split('.')[0] for r in seg)
    thr=sum(float(r[iT])*int(r[iI]) for r in seg)/max(ins,1)
    print(f"{k:6d} inst%={100*ins/tot_i:5.1f} samp%={100*sm/tot_s:5.1f} thr={thr:4.1f} top={ops.most_common(4)}")
