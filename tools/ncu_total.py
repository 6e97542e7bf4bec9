"""Aggregate an ncu launch list (csv) by kernel: total / count / average, sorted by total."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 2:]:
    if len(r) > vi:
        try: agg[r[ki][:80]].append(float(r[vi].replace(',', '')))
        except Exception: pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k:82s} n={len(v):4d} total={sum(v)/1000:9.1f} us ({100*sum(v)/tot:4.1f}%) avg={sum(v)/len(v)/1000:8.1f}")
print(f"total {tot/1000:.1f} us")
