#!/usr/bin/env python3
"""GPU-box parity sweep: large random batches of states and motions for every robot and shape mix against the compiled reference
(all host threads). Prints one line per case: decisions compared, differing, differing outside the 1e-5 m clearance band."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402
from oracle import port, refvamp  # noqa: E402

CORES = os.cpu_count() or 1
ctx = ompl_b200.Context(0)
CASES = [
    ("panda", "C2 500 spheres + 500 cuboids", synth.clutter_scene(1234, 500, 500), 1 << 21),
    ("panda", "all five primitive classes", synth.clutter_scene(7, 200, 200, n_capsules=60, n_zcuboids=60, n_zcapsules=60), 1 << 20),
    ("panda", "terrain + clutter", {**synth.clutter_scene(14, 60, 60, n_capsules=10), **synth.terrain(seed=5)}, 1 << 20),
    ("fetch", "clutter", synth.clutter_scene(9, 150, 150, n_capsules=30, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6), 1 << 20),
    ("ur5", "clutter", synth.clutter_scene(12, 300, 300), 1 << 20),
    ("baxter", "sparse clutter", synth.clutter_scene(5, 15, 15, n_capsules=4), 1 << 19),
]
total = bad_total = 0
for robot, name, scene, m in CASES:
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    re_ = synth.fill_env(refvamp.RefEnvironment(), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states(robot, m, 101)
    t0 = time.perf_counter()
    want = re_.validate_states(robot, q, nthreads=CORES)
    t_cpu = time.perf_counter() - t0
    got = ctx.validate_states(robot, env, q)
    bad = np.nonzero(got != want)[0]
    out = 0
    if bad.size:
        out = int((np.abs(pe.edge_min_clearance(robot, q[bad], q[bad], resolution=1)) >= 1e-5).sum())
    starts = q[want == 1][: m // 4]
    q0, q1 = synth.local_edges(robot, starts, 103, 0.05, 2.5)
    wm = re_.validate_motions(robot, q0, q1, nthreads=CORES)
    gm = ctx.validate_motions(robot, env, q0, q1)
    badm = np.nonzero(gm != wm)[0]
    outm = 0
    if badm.size:
        outm = int((np.abs(pe.edge_min_clearance(robot, q0[badm], q1[badm])) >= 1e-5).sum())
    total += m + len(q0)
    bad_total += out + outm
    print(f"{robot:7s} {name:32s} states {m} (valid {want.mean():.3f}) differ {bad.size} outside-band {out}; "
          f"motions {len(q0)} (valid {wm.mean():.3f}) differ {badm.size} outside-band {outm}; reference states/s {m / t_cpu:.3e} x{CORES} threads", flush=True)
print(f"TOTAL decisions {total}, outside the 1e-5 m band: {bad_total}")
