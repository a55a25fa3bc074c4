#!/usr/bin/env python3
"""Per-source-line view of an ncu report (needs -lineinfo + --import-source on): executed warp instructions, stall samples and
the share of no-instruction stalls per line. usage: ncu_lines.py <report.ncu-rep> [launch index] [top N] [file filter]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
filt = sys.argv[4] if len(sys.argv) > 4 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, agg, nsass = None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ism, ino = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_no_inst")
        continue
    if len(r) > 3 and r[2] == "-" and r[0].isdigit():
        agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[ie]), int(r[ism]), int(r[ino]) if r[ino].isdigit() else 0))
    elif len(r) > 3 and r[2].startswith("0x"):
        nsass += 1
tot = sum(a[3] for a in agg)
ts = sum(a[4] for a in agg)
print(f"warp instructions executed {tot:.4g}, samples {ts}, SASS lines with source {nsass}")
byfile = {}
for a in agg:
    v = byfile.setdefault(a[0], [0, 0])
    v[0] += a[3]
    v[1] += a[4]
print({k: (round(100 * v[0] / tot, 1), round(100 * v[1] / ts, 1)) for k, v in byfile.items() if v[0]})
sel = [a for a in agg if filt in a[0]]
for a in sorted(sel, key=lambda a: -a[4])[:top]:
    print(f"{a[0][:16]:16s} {a[1]:4d} exec {100 * a[3] / tot:5.1f}% smp {100 * a[4] / ts:5.1f}% noinst {100 * a[5] / max(a[4], 1):3.0f}% | {a[2]}")
