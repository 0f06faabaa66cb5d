#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
warp-level instructions executed, average active threads, stall samples and the dominant stall reasons.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > /tmp/cs.csv
    python tools/ncu_source_summary.py /tmp/cs.csv [--top 40] [--kernel-index 0]
"""
import argparse
import collections
import csv
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--kernel-index", type=int, default=0, help="which profiled launch of the report (sections repeat per launch)")
ap.add_argument("--by", default="samples", choices=["samples", "inst"])
a = ap.parse_args()

csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(a.csv, newline="")))
# the report is a sequence of sections: ["File Path", path] / ["Function Name", kernel] / header row / then per source
# line one aggregated row (line number, text, "-", "-", metrics...) followed by its SASS rows (empty line number)
sections = collections.OrderedDict()   # kernel name -> list of per-launch dicts (a kernel's files repeat per launch)
cur_file, hdr, per_line, col = None, None, None, {}
seen = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        launches = sections.setdefault(r[1], [])
        seen[(r[1], cur_file)] += 1
        idx = seen[(r[1], cur_file)] - 1
        while len(launches) <= idx:
            launches.append({})
        per_line = launches[idx]
        continue
    if r[0] == "Line No":
        hdr = r
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None or per_line is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def num(name):
        try:
            return float(r[col[name]])
        except (KeyError, ValueError):
            return 0.0
    c = per_line.setdefault((cur_file, line), collections.Counter())
    c["inst"] += num("Instructions Executed")
    c["tinst"] += num("Thread Instructions Executed")
    c["samples"] += num("# Samples")
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = num(h)
            if v:
                c[h] += v
    per_line[(cur_file, line, "text")] = r[1]

kernels = [(k, l[min(a.kernel_index, len(l) - 1)]) for k, l in sections.items() if l]
if not kernels:
    sys.exit("no kernel sections found")
name, per_line = kernels[0]
print(f"kernel: {name}")
items = [(k, v) for k, v in per_line.items() if len(k) == 2 and v["inst"] > 0]
tot_inst = sum(v["inst"] for _, v in items)
tot_tinst = sum(v["tinst"] for _, v in items)
tot_samples = sum(v["samples"] for _, v in items)
print(f"total warp instructions {tot_inst:.3g}, avg active threads {tot_tinst / max(tot_inst, 1):.2f}, stall samples {tot_samples:.0f}")
by_file = collections.Counter()
for (f, _), v in items:
    by_file[f] += v["samples"]
print("samples by file:", ", ".join(f"{f} {100 * s / max(tot_samples, 1):.1f}%" for f, s in by_file.most_common()))
items.sort(key=lambda kv: -kv[1][a.by])
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'thr':>5s}  top stalls | source")
for (f, ln), v in items[:a.top]:
    stalls = sorted(((h[6:], s) for h, s in v.items() if h.startswith("stall_")), key=lambda x: -x[1])[:3]
    st = " ".join(f"{h}:{100 * s / max(v['samples'], 1):.0f}" for h, s in stalls)
    text = per_line.get((f, ln, "text"), "")
    text = text.strip()[:90] if isinstance(text, str) else ""
    print(f"{f + ':' + str(ln):28s} {100 * v['samples'] / max(tot_samples, 1):6.2f} {100 * v['inst'] / max(tot_inst, 1):6.2f} "
          f"{v['tinst'] / max(v['inst'], 1):5.1f}  {st:34s} | {text}")
