#!/usr/bin/env python3
"""Dev helper (GPU box): kNN timing, exhaustive vs cell index, on roadmap-like data (valid Panda states of the C2 scene)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
r = ctx.prm_build("panda", env, n, 2, 42, want_weights=False)
states = torch.from_numpy(r["states"]).cuda()
idx = torch.empty((n, 16), dtype=torch.int64, device="cuda")
dst = torch.empty((n, 16), dtype=torch.float64, device="cuda")
for use_index in (True, False):
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.set_index(use_index)
    nn.add_device_f32(states.data_ptr(), n, 10.0)
    for rep in range(2):
        torch.cuda.synchronize(); ctx.synchronize()
        t = time.perf_counter()
        nn.query_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), self_base=0, exclude_self=True)
        ctx.synchronize()
        print(f"n={n} index={use_index} rep={rep}: {(time.perf_counter() - t) * 1e3:.1f} ms rows_visited={nn.rows_visited():.3e}", flush=True)
    if use_index:
        keep = (idx.clone(), dst.clone())
    else:
        print("identical:", bool((keep[0] == idx).all()), bool((keep[1] == dst).all()))
