#!/usr/bin/env python
"""Aggregate tools/sass_hotspots.py output (all lines) by file and line ranges.
usage: python tools/sass_regions.py <ncu-rep> <kernel-regex> <cubin> <mangled-substring> file:lo-hi[:label] ..."""
import subprocess, sys, re, collections
rep, kre, cubin, mangled = sys.argv[1:5]
regions = []
for r in sys.argv[5:]:
    parts = r.split(":")
    lo, hi = parts[1].split("-")
    regions.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else r))
out = subprocess.run([sys.executable, "tools/sass_hotspots.py", rep, kre, cubin, mangled, "100000"], capture_output=True, text=True).stdout
agg = collections.Counter(); samp = collections.Counter(); byfile = collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% samp thr/inst\s+[\d.]+\s+(\S+?):(\d+):", ln)
    if not m:
        continue
    pi, ps, f, l = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    byfile[f] += pi
    for (rf, lo, hi, lab) in regions:
        if rf == f and lo <= l <= hi:
            agg[lab] += pi; samp[lab] += ps
            break
    else:
        agg["other:" + f] += pi; samp["other:" + f] += ps
print(out.splitlines()[0]); print(out.splitlines()[1])
for k, v in agg.most_common():
    print(f"{v:6.1f}% inst {samp[k]:6.1f}% samp  {k}")
