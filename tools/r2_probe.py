#!/usr/bin/env python3
"""Dev helper (GPU box), round 2: (a) kNN phase split (index build vs query) on the roadmap's states, (b) the validation kernel on the
C2 all-valid sub-workload in its three scan variants (candidate grid / sorted ordered scan / unsorted ordered scan)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
scene, q0, q1 = bench.make_inputs(0)
stream = torch.cuda.Stream()


def timed(env, a, b, reps=5):
    d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.zeros(len(a), dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.validate_motions_ptr("panda", env, d0.data_ptr(), d1.data_ptr(), len(a), out.data_ptr(), stream=stream.cuda_stream)
            e1.record(stream)
        stream.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms[2:])), out.cpu().numpy()


if "knn" in sys.argv or len(sys.argv) == 1:
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    n = 1 << 20
    r = ctx.prm_build("panda", env, n, 2, 42, want_weights=False)
    states = torch.from_numpy(r["states"]).cuda()
    idx = torch.empty((n, 16), dtype=torch.int64, device="cuda")
    dst = torch.empty((n, 16), dtype=torch.float64, device="cuda")
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    for rep in range(3):
        nn.clear()
        ctx.synchronize()
        t0 = time.perf_counter()
        nn.add_device_f32(states.data_ptr(), n, 10.0)
        ctx.synchronize()
        t1 = time.perf_counter()
        nn.query_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), self_base=0, exclude_self=True)
        ctx.synchronize()
        t2 = time.perf_counter()
        nn.query_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), self_base=0, exclude_self=True)
        ctx.synchronize()
        t3 = time.perf_counter()
        print(json.dumps({"knn_rep": rep, "add_ms": (t1 - t0) * 1e3, "build_plus_query_ms": (t2 - t1) * 1e3, "query_only_ms": (t3 - t2) * 1e3}), flush=True)
    for nr in (1 << 18, 1 << 20):
        for rep in range(2):
            r = ctx.prm_build("panda", env, nr, 16, 42, want_weights=False)
        print(json.dumps(r["stats"]), flush=True)

if "dense" in sys.argv or "gridonly" in sys.argv or len(sys.argv) == 1:
    variants = (("grid", dict(sort=True)), ("sorted_ordered_scan", dict(sort=True, grid=False)), ("unsorted_ordered_scan", dict(sort=False)))
    for name, kw in variants[:1] if "gridonly" in sys.argv else variants:
        env = synth.fill_env(ompl_b200.Environment(ctx), scene, **kw)
        ms_mixed, v = timed(env, q0, q1, reps=3)
        vi = np.nonzero(v == 1)[0]
        rep = np.resize(vi, 1 << 17)
        ms_av, va = timed(env, q0[rep], q1[rep])
        print(json.dumps({"variant": name, "mixed_ms_2^20": ms_mixed, "valid_fraction": float(v.mean()), "unique_valid": int(len(vi)),
                          "all_valid_ms_2^17": ms_av, "all_valid_states_per_s": len(rep) * 64 / (ms_av * 1e-3), "all_ok": bool(va.all())}), flush=True)
