#!/usr/bin/env python
"""Attribute executed SASS instructions (ncu source page CSV) to CUDA source lines (nvdisasm -g line info).
usage: python tools/sass_hotspots.py <ncu-rep> <kernel-regex> <cubin> <mangled-substring> [top]
The function section of the cubin is chosen by <mangled-substring> (e.g. shade_bwd_kernelILi4ELb0E); instructions are
matched by order within the kernel, so the instruction counts of the two listings must agree (checked)."""
import collections
import csv
import glob
import re
import subprocess
import sys

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ie, ss, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
te = hdr.index("Thread Instructions Executed")
ins = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > te and r[ie].isdigit():
        ins.append((r[src].strip(), int(r[ie]), int(r[ss]), int(r[te])))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
out, cur, infn = [], None, False
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = mangled in m.group(1)
        continue
    if re.match(r"\s*\.section", ln):
        infn = False
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        out.append((m.group(1).strip(), cur))
print(f"kernel {kname}: ncu {len(ins)} instrs, nvdisasm {len(out)} instrs")
agg = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
ops = collections.Counter()
n = min(len(ins), len(out))
for i in range(n):
    key = out[i][1][:2] if out[i][1] else ("?", 0)
    agg[key] += ins[i][1]; samp[key] += ins[i][2]; thr[key] += ins[i][3]
    ops[ins[i][0].split()[0] if not ins[i][0].startswith("@") else ins[i][0].split()[1]] += ins[i][1]
tot = sum(agg.values()); stot = sum(samp.values())
print(f"total warp insts {tot}, samples {stot}")
files = {}
for (f, l), c in agg.most_common(top):
    if f not in files:
        try:
            p = glob.glob(f"/root/repo/**/{f}", recursive=True)[0]
            files[f] = open(p).read().splitlines()
        except Exception:
            files[f] = []
    text = files[f][l - 1].strip()[:100] if 0 < l <= len(files[f]) else ""
    print(f"{100*c/tot:5.1f}% inst {100*samp[(f,l)]/max(stot,1):5.1f}% samp thr/inst {thr[(f,l)]/max(c,1):4.1f}  {f}:{l}: {text}")
print("opcode mix:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in ops.most_common(25)))
