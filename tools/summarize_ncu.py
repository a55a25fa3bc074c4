#!/usr/bin/env python3
"""Turn ncu CSV logs (--csv --log-file) into the small summaries committed under profiles/.
usage: summarize_ncu.py launches <csv>            -> per-kernel time shares (markdown on stdout)
       summarize_ncu.py flops <csv> <facts.json>  -> profiles/validate_flops.json content on stdout (the all-valid launch = id 2)"""
import csv
import json
import sys
from collections import defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    return name.split("(")[0].replace("b200mp::", "").replace("(anonymous namespace)::", "")[:70]


def launches(path):
    per = defaultdict(lambda: [0, 0.0])
    for r in rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        per[k][0] += 1
        per[k][1] += ns
    tot = sum(v[1] for v in per.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % |")
    print(f"| all | {sum(v[0] for v in per.values())} | {tot / 1e6:.3f} | 100 % |")


def flops(path, facts_path):
    facts = json.load(open(facts_path))
    by_id = defaultdict(dict)
    for r in rows(path):
        by_id[int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = sorted(by_id)
    m = by_id[ids[2]]      # launch #2 of validate_motions = the all-valid roofline launch
    ffma = m["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
    fadd = m["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
    fmul = m["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
    out = dict(facts)
    out.update(flops_per_launch=2 * ffma + fadd + fmul, ffma=ffma, fadd=fadd, fmul=fmul, duration_ns_under_ncu=m.get("gpu__time_duration.sum"),
               dram_bytes=m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
               inst_executed=m.get("smsp__inst_executed.sum"), inst_fma_pipe=m.get("sm__inst_executed_pipe_fma.sum"),
               source="ncu smsp__sass_thread_inst_executed_op_{ffma,fadd,fmul}_pred_on.sum, tools/profile_target.py launch #2")
    if len(ids) > 3:  # launch #3: the same motions through the ordered scan (dense variant)
        d = by_id[ids[3]]
        out["dense"] = dict(flops_per_launch=2 * d["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"] +
                            d["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"] + d["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"],
                            duration_ns_under_ncu=d.get("gpu__time_duration.sum"),
                            dram_bytes=d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0),
                            inst_executed=d.get("smsp__inst_executed.sum"), inst_fma_pipe=d.get("sm__inst_executed_pipe_fma.sum"))
    mixed = by_id[ids[1]]
    out["mixed_launch"] = dict(flops=2 * mixed["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"] +
                               mixed["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"] + mixed["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"],
                               duration_ns_under_ncu=mixed.get("gpu__time_duration.sum"))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        flops(sys.argv[2], sys.argv[3])
