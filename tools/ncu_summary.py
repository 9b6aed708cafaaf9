#!/usr/bin/env python
"""Dev tool: one markdown table row per `ncu --page raw --csv` file (the metrics B200_PROFILING.md names).

usage: tools/ncu_summary.py profiles/r01_ncu_*_raw.csv
"""
import csv
import sys

WANT = [
    ("time us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("DRAM rd MB", "dram__bytes_read.sum"),
    ("DRAM wr MB", "dram__bytes_write.sum"),
    ("DRAM % peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 % peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("warp inst", "smsp__inst_executed.sum"),
]
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}


def table(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return []
    return [row(rows[0], rows[1], v) for v in rows[2:] if len(v) == len(rows[0])]


def row(hdr, units, vals):
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    out = {"kernel": name.split("(")[0][-60:]}
    for label, metric in WANT:
        if metric in hdr:
            i = hdr.index(metric)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                out[label] = vals[i]
                continue
            if units[i] in SCALE and "MB" in label:
                v *= SCALE[units[i]]
            out[label] = f"{v:.3g}" if abs(v) < 1e6 else f"{v:.4g}"
        else:
            out[label] = "-"
    return out


rows = [(p, r) for p in sys.argv[1:] for r in table(p)]
labels = ["kernel"] + [l for l, _ in WANT]
print("| file | " + " | ".join(labels) + " |")
print("|" + "---|" * (len(labels) + 1))
for p, r in rows:
    if r:
        print("| " + p.split("/")[-1] + " | " + " | ".join(str(r[l]) for l in labels) + " |")
