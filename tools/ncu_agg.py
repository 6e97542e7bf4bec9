import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hdr+2:]:
    if len(r)>vi:
        try: agg[r[ki][:70]].append(float(r[vi].replace(',','')))
        except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us  min={min(v)/1000:8.1f}")
