import csv,re,sys,subprocess
rep=sys.argv[1]; reason=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
KS=int(sys.argv[4]) if len(sys.argv)>4 else 0
seg=rows[starts[KS]:starts[KS+1] if len(starts)>KS+1 else None]
hdr=seg[1]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples'); iR=hdr.index(reason)
recs=[(int(r[iR]),idx,r[iS].strip(),int(r[iE]),int(r[iSamp])) for idx,r in enumerate(seg[2:]) if len(r)>iR and r[iR].isdigit()]
tot=sum(x[0] for x in recs)
print('total',reason,tot)
for x in sorted(recs,reverse=True)[:topn]:
    print(x[1], x[0], x[2][:90], 'exec',x[3],'samples',x[4])
