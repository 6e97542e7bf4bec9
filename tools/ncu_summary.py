#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) and a launch-list CSV into profiles/<tag>_*.{md,csv}.
usage: python tools/ncu_summary.py <tag> [note]"""
import csv, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
# optional: --reps a.ncu-rep b.ncu-rep (instead of gpurun_out/<tag>_prof.ncu-rep; later files win per kernel name),
#           --launches file.csv, --keep-traffic (do not rewrite profiles/traffic.json: the captures are not the C2 step)
opt = {"--reps": [], "--launches": []}
cur = None
for a in sys.argv[2:]:
    if a in opt:
        cur = a
    elif a == "--keep-traffic":
        opt[a] = True
        cur = None
    elif cur:
        opt[cur].append(a)
reps = opt["--reps"] or [os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")]
lau = opt["--launches"][0] if opt["--launches"] else os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
out = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__inst_executed.sum", "warp insts"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_bytes.sum", "L2 bytes")]
lines = [f"# ncu summary `{tag}`", "", note, ""]
reps = [r for r in reps if os.path.isfile(r)]
if reps:
    seen = set()
    traffic = {}
    lines += ["## `ncu --set full --clock-control none` (one launch per kernel; cold-cache, serialised)", "",
              "captures: " + ", ".join(f"`{os.path.relpath(r, ROOT)}`" for r in reps), "",
              "| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---|" * len(WANT)]
    allrows = []
    for rep in reversed(reps):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        allrows += [(r, {h: i for i, h in enumerate(hdr)}, units) for r in rows[2:]]
    for r, idx, units in allrows:
        k = r[idx["Kernel Name"]]
        if k in seen:
            continue
        seen.add(k)
        cells = []
        for m, _ in WANT:
            if m in idx:
                v = r[idx[m]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[idx[m]]}".strip())
            else:
                cells.append("-")
        short = k.split("(")[0].replace("void ", "").replace("hfr::", "").replace("<unnamed>::", "")
        try:   # DRAM bytes of this launch (read + write), for bench.py's roofline.traffic
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]
            traffic[short.split("<")[0]] = float(r[rd]) * mult[units[rd]] + float(r[wr]) * mult[units[wr]]
        except (KeyError, ValueError):
            pass
        lines.append(f"| `{short}` | " + " | ".join(cells) + " |")
    lines.append("")
    import json
    if not opt.get("--keep-traffic"):
        json.dump({"tag": tag, "source": f"profiles/{tag}_summary.md (ncu --set full, one launch, C2 B=64)",
                   "dram_bytes_per_launch": traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
if os.path.isfile(lau):
    rows = [r for r in csv.reader(open(lau)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v_us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
        k = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    lines += ["## launch list (`--metrics gpu__time_duration.sum`, first 400 launches of `bench.py --steps 3 --warmup 3`)", "",
              "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    import shutil
    shutil.copy(lau, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
open(out, "w").write("\n".join(lines) + "\n")
print(open(out).read())
