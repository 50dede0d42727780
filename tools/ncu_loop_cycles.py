import csv,sys,collections,subprocess
rep=sys.argv[1]; cyc_per_iter=float(sys.argv[2]); mode=sys.argv[3] if len(sys.argv)>3 else 'fam'
txt=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
mx=max(int(r[ix['Instructions Executed']] or 0) for r in data)
loop=[r for r in data if int(r[ix['Instructions Executed']] or 0)>=mx*0.5]
tot=sum(int(r[ix['# Samples']]) for r in loop)
per=tot/cyc_per_iter
print('loop instrs',len(loop),'exec',mx,'samples',tot)
if mode=='fam':
    fam=collections.Counter(); famc=collections.Counter()
    for r in loop:
        w=r[ix['Source']].split()
        op=w[1] if w[0].startswith('@') else w[0]
        op0=op.split('.')[0]
        fam[op0]+=int(r[ix['Instructions Executed']])/mx; famc[op0]+=int(r[ix['# Samples']])/per
    for k,v in sorted(famc.items(),key=lambda x:-x[1]):
        print("%-10s n=%6.1f cycles=%7.1f  per=%.2f"%(k,fam[k],v,v/fam[k]))
else:
    cum=0
    for r in loop:
        c=int(r[ix['# Samples']])/per; cum+=c
        print("%s %5.1f %6.0f  %s%s"%(r[ix['Address']][-4:],c,cum,r[ix['Source']][:80],'  <<<' if c>=5 else ''))
