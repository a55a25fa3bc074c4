#!/usr/bin/env python3
"""Dev helper (GPU box): parity vs the checkers + rough throughput of the validation kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402
from oracle import port  # noqa: E402

try:
    from oracle import refvamp
    HAVE_REF = refvamp.available()
except Exception:
    HAVE_REF = False


def checker_env(scene):
    if HAVE_REF:
        return synth.fill_env(refvamp.RefEnvironment(), scene), "reference"
    return synth.fill_env(port.PortEnvironment(), scene), "port"


def report(name, got, want, clr_fn):
    bad = np.nonzero(got != want)[0]
    msg = f"{name}: n={len(got)} valid={int(want.sum())} mismatches={bad.size}"
    if bad.size:
        clr = clr_fn(bad[:64])
        msg += f" |clearance| of mismatches: max={np.abs(clr).max():.3e} (band 1e-5: {(np.abs(clr) < 1e-5).sum()}/{len(clr)})"
    print(msg, flush=True)


def time_device(ctx, env, robot, q0, q1, reps=3):
    d0 = torch.from_numpy(q0).cuda()
    d1 = torch.from_numpy(q1).cuda()
    out = torch.empty(q0.shape[0], dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        st = ts.cuda_stream
        ctx.validate_motions_ptr(robot, env, d0.data_ptr(), d1.data_ptr(), q0.shape[0], out.data_ptr(), stream=st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ts)
        for _ in range(reps):
            ctx.validate_motions_ptr(robot, env, d0.data_ptr(), d1.data_ptr(), q0.shape[0], out.data_ptr(), stream=st)
        b.record(ts)
        torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out.cpu().numpy()


def main():
    ctx = ompl_b200.Context(0)
    print("device", ctx.device, "SMs", ctx.sm_count, "fp32 peak TFLOP/s", ctx.fp32_peak_tflops(), "checker ref:", HAVE_REF)
    nt = os.cpu_count() or 1
    for scene_name, scene in (("cage", synth.cage_scene()),
                              ("clutter_all", synth.clutter_scene(3, 200, 200, n_capsules=50, n_zcuboids=50, n_zcapsules=50)),
                              ("C2", synth.clutter_scene(1234, 500, 500))):
        env = synth.fill_env(ompl_b200.Environment(ctx), scene)
        env_scan = synth.fill_env(ompl_b200.Environment(ctx), scene, grid=False)
        chk, kind = checker_env(scene)
        pe = synth.fill_env(port.PortEnvironment(), scene)
        q = synth.uniform_states("panda", 200000, 5)
        t = time.time(); got = ctx.validate_states("panda", env, q); tg = time.time() - t
        t = time.time(); want = chk.validate_states("panda", q, nthreads=nt); tc = time.time() - t
        report(f"[{scene_name}] states vs {kind} (gpu {tg*1e3:.1f} ms, cpu {tc*1e3:.1f} ms x{nt}thr)", got, want,
               lambda b: pe.edge_min_clearance("panda", q[b], q[b], resolution=1))
        starts = q[want == 1][:20000]
        q0, q1 = synth.local_edges("panda", starts, 9, 0.05, 2.5)
        t = time.time(); got = ctx.validate_motions("panda", env, q0, q1); tg = time.time() - t
        t = time.time(); want = chk.validate_motions("panda", q0, q1, nthreads=nt); tc = time.time() - t
        report(f"[{scene_name}] local motions vs {kind} (gpu {tg*1e3:.1f} ms, cpu {tc*1e3:.1f} ms x{nt}thr)", got, want,
               lambda b: pe.edge_min_clearance("panda", q0[b], q1[b]))
        q0, q1 = synth.fixed_length_edges("panda", 1 << 18, 1234)
        ms, got = time_device(ctx, env, "panda", q0, q1)
        t = time.time(); want = chk.validate_motions("panda", q0, q1, nthreads=nt); tc = time.time() - t
        report(f"[{scene_name}] 64-step motions vs {kind}: gpu {ms:.2f} ms = {len(q0)/ms*1e3:.3e} edges/s; cpu {tc:.2f}s = {len(q0)/tc:.3e} edges/s",
               got, want, lambda b: pe.edge_min_clearance("panda", q0[b], q1[b]))
        # all-valid subset replicated
        v = np.nonzero(want == 1)[0]
        if v.size:
            rep = np.resize(v, 1 << 16)
            ms, got = time_device(ctx, env, "panda", q0[rep], q1[rep])
            print(f"[{scene_name}] all-valid x{len(rep)} ({v.size} unique): {ms:.2f} ms = {len(rep)/ms*1e3:.3e} edges/s = {len(rep)*64/ms*1e3:.3e} states/s; gpu valid frac {got.mean():.4f}", flush=True)
            ms2, _ = time_device(ctx, env_scan, "panda", q0[rep], q1[rep])
            ms3, _ = time_device(ctx, env_scan, "panda", q0, q1)
            print(f"[{scene_name}] grid off: all-valid {ms2:.2f} ms, mixed 2^18 {ms3:.2f} ms", flush=True)


if __name__ == "__main__":
    main()
