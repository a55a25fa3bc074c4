#!/usr/bin/env python3
"""Executed warp instructions / hot SASS per function of validate.cuh (and per other file) for one profiled kernel.
usage: ncu_regions.py <report.ncu-rep>"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, sass = None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ie, ism = r.index("Instructions Executed"), r.index("# Samples")
    elif len(r) > 3 and r[2] == "-" and r[0].isdigit():
        curline = (cur, int(r[0]))
    elif len(r) > 3 and r[2].startswith("0x"):
        g = lambda i: int(r[i]) if r[i].isdigit() else 0
        sass.setdefault(r[2], (curline, g(ie), g(ism)))
top = max(e for _, e, _ in sass.values())
tot = sum(e for _, e, _ in sass.values())
ts = sum(s for _, _, s in sass.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ompl_b200", "csrc", "validate.cuh")).read().split("\n")
marks = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(template|__device__|static __device__|__global__|struct|static __global__)", l)]


def region(f, l):
    if f != "validate.cuh":
        return f
    name = "?"
    for ln, t in marks:
        if ln <= l:
            name = f"{ln}:{(src[ln - 1] if not src[ln - 1].startswith('template') else src[ln])[:80]}"
    return name


reg = {}
for a, ((f, l), e, sm) in sass.items():
    v = reg.setdefault(region(f, l), [0, 0, 0, 0])
    v[0] += 1
    v[1] += 1 if e >= 0.002 * top else 0
    v[2] += e
    v[3] += sm
print(f"warp instructions {tot:.4g}; hot SASS {sum(v[1] for v in reg.values())}")
for k, v in sorted(reg.items(), key=lambda kv: -kv[1][2])[:26]:
    print(f"exec {100 * v[2] / tot:5.1f}%  samples {100 * v[3] / ts:5.1f}%  hot {v[1]:4d} / sass {v[0]:4d}  {k}")
