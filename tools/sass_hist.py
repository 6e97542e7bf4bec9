#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of libhifihr_b200.so (cuobjdump -sass), written to profiles/<tag>_sass.md.
usage: python tools/sass_hist.py <tag>
Shows which kernels carry the Blackwell-specific instructions (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UBLKCP = cp.async.bulk, SYNCS = mbarrier) and which still use global floating-point reductions (RED / ATOMG)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lib = os.path.join(ROOT, "hifihr_b200", "libhifihr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").replace("hfr::", "")
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
cols = [("UTC*MMA", lambda o: o.startswith("UTC") and "MMA" in o), ("LDTM", lambda o: o.startswith("LDTM")),
        ("UBLKCP", lambda o: o.startswith("UBLKCP")), ("UTMA*", lambda o: o.startswith("UTMA")),
        ("SYNCS", lambda o: o.startswith("SYNCS")), ("RED.F32", lambda o: o.startswith("RED") and "F32" in o),
        ("RED/ATOMG int", lambda o: (o.startswith("RED") or o.startswith("ATOMG")) and "F32" not in o),
        ("ATOMS", lambda o: o.startswith("ATOMS")), ("LDG", lambda o: o.startswith("LDG")), ("STG", lambda o: o.startswith("STG")),
        ("LDS", lambda o: o.startswith("LDS")), ("FFMA", lambda o: o.startswith("FFMA")), ("MUFU", lambda o: o.startswith("MUFU")),
        ("BAR", lambda o: o.startswith("BAR")), ("total", lambda o: True)]
lines = [f"# SASS opcode histogram `{tag}` (static instruction counts, `cuobjdump -sass hifihr_b200/libhifihr_b200.so`)", "",
         "| kernel | " + " | ".join(c for c, _ in cols) + " |", "|---|" + "---|" * len(cols)]
for k, h in hist.items():
    lines.append(f"| `{k}` | " + " | ".join(str(sum(v for o, v in h.items() if f(o))) for _, f in cols) + " |")
open(os.path.join(ROOT, "profiles", f"{tag}_sass.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:4] + [l for l in lines[4:] if re.search(r"blend|raster_shade_fwd|shade_bwd_tiled|loss", l)]))
