import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); mu=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    v=float(r[mv].replace(",","")); v = v/1000 if r[mu]=="ns" else (v*1000 if r[mu]=="ms" else v)
    n=r[kn].split("(")[0][-50:]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 14]:
    print("%-52s n=%4d total=%9.1f us  avg=%8.1f us  %5.1f%%"%(k,n,t,t/n,100*t/tot))
