#!/usr/bin/env python3
"""Per-launch DRAM bytes of the trace kernel from
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pt_trace --csv ...
-> profiles/<name>.json (read by bench.py for roofline.traffic) and a markdown table on stdout.
    python tools/ncu_dram_summary.py gpurun_out/trace_dram_bench.csv profiles/r01_trace_dram.json --launches-per-step 24 --config '{...}'"""
import argparse, csv, json
ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("out")
ap.add_argument("--launches-per-step", type=int, required=True)
ap.add_argument("--skip-steps", type=int, default=1, help="steps (warm-up) to drop from the front of the capture")
ap.add_argument("--config", default="{}")
a = ap.parse_args()
rows = list(csv.reader(open(a.csv, newline="")))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i]; col = {h: j for j, h in enumerate(hdr)}
per = {}
for r in rows[i + 1:]:
    if len(r) < len(hdr):
        continue
    per.setdefault(int(r[col["ID"]]), {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
ids = sorted(per)[a.skip_steps * a.launches_per_step:][:a.launches_per_step]
L = [per[k] for k in ids]
rd = sum(x["dram__bytes_read.sum"] for x in L); wr = sum(x["dram__bytes_write.sum"] for x in L); ns = sum(x["gpu__time_duration.sum"] for x in L)
out = {"kernel": "pt_trace_kernel", "source": a.csv, "config": json.loads(a.config), "launches_per_step": len(L),
       "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr, "traffic_bytes_per_launch": (rd + wr) / len(L),
       "ncu_kernel_ms_per_step": ns / 1e6, "dram_gbs_under_ncu": (rd + wr) / ns,
       "per_launch": [{"ms": x["gpu__time_duration.sum"] / 1e6, "read": x["dram__bytes_read.sum"], "write": x["dram__bytes_write.sum"]} for x in L]}
json.dump(out, open(a.out, "w"), indent=1)
print(f"{len(L)} trace launches of one step: DRAM read {rd/1e9:.1f} GB, write {wr/1e9:.1f} GB, {ns/1e6:.1f} ms under ncu "
      f"({(rd+wr)/ns:.0f} GB/s); traffic per launch {(rd+wr)/len(L)/1e9:.2f} GB")
print("| launch | ms (ncu) | DRAM read GB | DRAM write GB |\n|---|---|---|---|")
for k, x in enumerate(L):
    print(f"| {k} | {x['gpu__time_duration.sum']/1e6:.3f} | {x['dram__bytes_read.sum']/1e9:.2f} | {x['dram__bytes_write.sum']/1e9:.2f} |")
