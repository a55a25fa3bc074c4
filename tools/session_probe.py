#!/usr/bin/env python3
"""Dev helper (GPU box): latency of a single checkMotion -- through b200mp_validate_motions (one launch + synchronise per call) and
through a b200mp_session (resident kernel + mailbox) -- on the motions an RRT-Connect run in the C1 cage actually asks, and the
reference's own time for the same calls on one host core."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402
from tests.rrtconnect import solve  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
lo, hi = synth.BOUNDS["panda"]
calls = []
batch = lambda a, b: ctx.validate_motions("panda", env, np.asarray(a, np.float64)[None], np.asarray(b, np.float64)[None])[0]  # noqa: E731


def record(a, b):
    calls.append((np.asarray(a, np.float64).copy(), np.asarray(b, np.float64).copy()))
    return batch(a, b)


for seed in range(1, 6):
    solve(synth.CAGE_START, synth.CAGE_GOAL, lo, hi, record, seed)
q0 = np.stack([c[0] for c in calls])
q1 = np.stack([c[1] for c in calls])
want = ctx.validate_motions("panda", env, q0, q1)
L = ctx._L
out = np.zeros(1, np.uint8)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        got = np.array([fn(i) for i in range(len(q0))], np.uint8)
        best = min(best, (time.perf_counter() - t0) / len(q0))
    assert (got == want).all()
    return best * 1e6


P0, P1, PO, ST = q0.ctypes.data, q1.ctypes.data, out.ctypes.data, q0.strides[0]  # (raw addresses: no numpy work inside the timed loops)
H, E_ = ctx._h, env._h


def via_batch(i):
    L.b200mp_validate_motions_f64(H, 0, E_, P0 + i * ST, P1 + i * ST, 1, 0, PO)
    return out[0]


s = ompl_b200.Session(ctx, "panda", env)


SH = s._h


def via_session(i):
    L.b200mp_session_check_motion_f64(SH, P0 + i * ST, P1 + i * ST, 0, PO)
    return out[0]


def via_nothing(i):  # the loop's own cost (a ctypes call of the same shape that does no GPU work)
    L.b200mp_session_stats(SH, None, None)
    return out[0]


t_loop = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    for i in range(len(q0)):
        via_nothing(i)
    t_loop = min(t_loop, (time.perf_counter() - t0) / len(q0) * 1e6)
line = {"us_python_loop_overhead": t_loop, "calls": int(len(q0)), "valid_fraction": float(want.mean()), "us_per_call_launch_path": timed(via_batch), "us_per_call_session": timed(via_session),
        "session": s.stats()}
try:
    from oracle import refvamp
    if refvamp.available():
        re_ = synth.fill_env(refvamp.RefEnvironment(), synth.cage_scene())
        a32, b32 = q0.astype(np.float32), q1.astype(np.float32)
        re_.validate_motions("panda", a32[:8], b32[:8])
        t0 = time.perf_counter()
        r = re_.validate_motions("panda", a32, b32, nthreads=1)
        line["us_per_call_reference_1core_batched_loop"] = (time.perf_counter() - t0) / len(q0) * 1e6
        line["reference_decisions_differing"] = int((r != want).sum())
except Exception as e:  # noqa: BLE001
    line["reference"] = repr(e)
print(json.dumps(line), flush=True)
s.close()
