import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
cur_file=None; cur_fn=None; hdr=None
agg=collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': cur_fn=r[1][:40]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None: continue
    if r[0]!='':   # cuda source line row (aggregated)
        iI=hdr.index('Instructions Executed'); iS=hdr.index('# Samples')
        try: ie=float(r[iI]); s=float(r[iS])
        except: continue
        key=(cur_fn,cur_file,int(r[0]),r[1].strip()[:100])
        a=agg.setdefault(key,[0,0]); a[0]+=ie; a[1]+=s
fns=collections.OrderedDict()
for (fn,f,l,src),(ie,s) in agg.items(): fns.setdefault(fn,[]).append((f,l,src,ie,s))
for fn,items in fns.items():
    tot=sum(i[3] for i in items); st=sum(i[4] for i in items)
    print('=====',fn,'inst %.3g samples %d'%(tot,st))
    for f,l,src,ie,s in items:
        if ie>0.006*tot or s>0.012*st: print('%-16s %4d  inst %5.1f%%  stall-samp %5.1f%%  %s'%(f[:16],l,100*ie/tot,100*s/max(st,1),src))
