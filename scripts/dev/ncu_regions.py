import csv,re,collections,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
KI=int(sys.argv[2]) if len(sys.argv)>2 else 0
hdr=rows[0]; r=rows[2+KI]
want=['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','launch__registers_per_thread','smsp__average_warp_latency_per_inst_issued.ratio','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct']
for h,v in zip(hdr,r):
    if h in want or re.match(r'smsp__average_warps_issue_stalled_.*_per_issue_active.ratio',h): print(h.replace('smsp__average_warps_issue_stalled_','  stall_').replace('_per_issue_active.ratio',''),v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
names_src=[rows[i][1] if len(rows[i])>1 else '' for i in starts]
rawname=r[hdr.index('Kernel Name')][:40]
KS=[k for k,n in enumerate(names_src) if rawname[:30] in n.replace('hipgp::','')]
KS=[k for k,n in enumerate(names_src) if rawname.split('<')[0].split()[-1] in n]
KS=KS[0] if KS else KI
seg=rows[starts[KS]:starts[KS+1] if len(starts)>KS+1 else None]
print('kernel:',seg[0][1][:80] if len(seg[0])>1 else seg[0])
hdr=seg[1]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
names=[hdr[i] for i in stall_cols]
reg=[];cur=[]
for r in seg[2:]:
    s=r[iS].strip()
    m=re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)',s)
    if not m: continue
    cur.append((s,m.group(2),int(r[iE]),int(r[iSamp]),[int(r[i]) for i in stall_cols]))
    if 'BAR.SYNC' in s or 'BAR.ARV' in s or 'SYNCS' in s or 'WARPSYNC' in s: reg.append(cur);cur=[]
reg.append(cur)
tot=sum(x[2] for rg in reg for x in rg); tots=sum(x[3] for rg in reg for x in rg)
print('total instr',tot,'samples',tots)
for i,rg in enumerate(reg):
    t=sum(x[2] for x in rg); sm=sum(x[3] for x in rg)
    if sm<0.01*tots: continue
    oc=collections.Counter()
    for x in rg: oc[x[1]]+=x[2]
    stt=[sum(x[4][k] for x in rg) for k in range(len(names))]
    top=sorted(zip(stt,names),reverse=True)[:6]
    print(f'region {i} end={rg[-1][0][:40]!r}: static={len(rg)} dyn={100*t/tot:.1f}% samples={100*sm/tots:.1f}%')
    print('   ops:',[(o,round(100*c/max(t,1),1)) for o,c in oc.most_common(7)])
    print('   stalls:',[(n.replace('stall_',''),v) for v,n in top])
