#!/usr/bin/env python3
"""Key counters of `ncu --set full` reports as one markdown table (one column per report).
    python tools/ncu_raw_summary.py label=path.ncu-rep [label=path.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_static", "static smem / CTA"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("inst_executed", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe FMA heavy"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes.sum.per_second", "DRAM bandwidth"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (of ncu peak)"),
    ("sass__inst_executed_global_loads", "global load instructions"),
    ("sass__inst_executed_local_loads", "local load instructions"),
    ("sass__inst_executed_local_stores", "local store instructions"),
    ("sass__inst_executed_shared_loads", "shared load instructions"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_branch_resolving", "stall samples: branch resolving"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall samples: selected (issuing)"),
    ("smsp__pcsamp_sample_buffer_full", None),
]

cols = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    d = dict(zip(hdr, launches[0]))
    u = dict(zip(hdr, units))
    cols.append((label, d, u, len(launches)))

print("| counter | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
print("| kernel | " + " | ".join(c[1].get("Kernel Name", "?").split("(")[0].replace("void ", "") for c in cols) + " |")
for key, name in KEYS:
    if name is None:
        continue
    vals = []
    for _, d, u, _ in cols:
        v = d.get(key)
        if v in (None, ""):
            vals.append("–")
            continue
        try:
            f = float(v.replace(",", ""))
            v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.2f}"
        except ValueError:
            pass
        vals.append(f"{v} {u.get(key, '')}".strip())
    print(f"| {name} | " + " | ".join(vals) + " |")
