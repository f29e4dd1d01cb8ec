import csv, collections, sys
def summarize(path, title):
    rows=list(csv.reader(open(path, errors='ignore')))
    hdr=None; agg=collections.OrderedDict(); by_id={}
    for r in rows:
        if len(r)>5 and r[0]=='ID': hdr=r; continue
        if hdr and len(r)==len(hdr):
            k=r[hdr.index('Kernel Name')].split('(')[0].replace('<unnamed>::','').replace('void ','')
            m=r[hdr.index('Metric Name')]; v=float(r[hdr.index('Metric Value')].replace(',','')); u=r[hdr.index('Metric Unit')]
            d=by_id.setdefault((r[0],k),{})
            if m=='gpu__time_duration.sum': d['t']=v*{'ns':1e-6,'us':1e-3,'ms':1.0}.get(u,1e-6)
            else:
                d[m]=v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    for (i,k),d in by_id.items():
        a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=d.get('t',0); a[2]+=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0)
    tot=sum(a[1] for a in agg.values())
    out=[title]
    for k,(n,t,b) in sorted(agg.items(), key=lambda x:-x[1][1]):
        out.append(f"{k[:58]:58s} launches {n:4d}  total {t:10.3f} ms  share {100*t/tot:5.1f} %  avg {t/n:9.3f} ms  dram/launch {b/n/1e6:10.2f} MB")
    return "\n".join(out)
print(summarize(sys.argv[1], sys.argv[2]))
