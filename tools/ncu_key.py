#!/usr/bin/env python3
"""Key metrics of every launch in an ncu report. usage: ncu_key.py <report.ncu-rep>"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        vals = [r[i] for r in data]
        try:
            if max(float(v.replace(",", "")) for v in vals) < 0.05 and "stalled" in w:
                continue
        except ValueError:
            pass
        print(f"{w.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} [{units[i]}]: " + "  ".join(vals))
