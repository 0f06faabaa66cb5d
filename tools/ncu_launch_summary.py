#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, average).
    python tools/ncu_launch_summary.py gpurun_out/launches_bench.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], newline="")))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
col = {h: i for i, h in enumerate(hdr)}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("b200rt::", "").replace("void ", "")
    ns = float(r[col["Metric Value"]])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns; tot += ns
print(f"{len(rows) - start} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised by ncu: compare shares, not absolutes)")
print(f"{'kernel':64s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k[:64]:64s} {n:8d} {ns / 1e6:10.2f} {100 * ns / tot:6.2f}% {ns / n / 1e3:9.1f}")
