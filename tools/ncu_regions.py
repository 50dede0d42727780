import csv,sys,subprocess,collections
rep=sys.argv[1]; cyc=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 10
txt=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines())); hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
ex=[int(r[ix['Instructions Executed']] or 0) for r in data]
common=[v for v,c in collections.Counter(ex).most_common(6) if v>1000]
top=max(common)
loop=[(i,r) for i,r in enumerate(data) if int(r[ix['Instructions Executed']] or 0)>=top*0.9]
regs=[];cur=[loop[0]]
for a,b in zip(loop,loop[1:]):
    if b[0]-a[0]>40: regs.append(cur);cur=[]
    cur.append(b)
regs.append(cur)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for reg in regs:
    tot=sum(int(r[ix['# Samples']]) for _,r in reg)
    print('region',reg[0][1][ix['Address']][-5:],reg[-1][1][ix['Address']][-5:],'instrs',len(reg),'samples',tot)
big=[reg for reg in regs if sum(int(r[ix['# Samples']]) for _,r in reg)>1500]
for reg in big:
    tot=sum(int(r[ix['# Samples']]) for _,r in reg); per=tot/cyc
    agg={s:0 for s in stalls}; fam=collections.Counter(); famn=collections.Counter()
    for _,r in reg:
        for s in stalls: agg[s]+=int(r[ix[s]] or 0)
        w=r[ix['Source']].split(); op=(w[1] if w[0].startswith('@') else w[0]).split('.')[0]
        fam[op]+=int(r[ix['# Samples']])/per; famn[op]+=1
    print('== region',reg[0][1][ix['Address']][-5:],'assuming',cyc,'cycles')
    print(sorted(((k[6:],round(v/per)) for k,v in agg.items() if v),key=lambda x:-x[1])[:7])
    print([(k,famn[k],round(v)) for k,v in fam.most_common(10)])
    for _,r in reg:
        c=int(r[ix['# Samples']])/per
        if c>=thr: print("   %s %6.1f  %s"%(r[ix['Address']][-5:],c,r[ix['Source']][:70]))
