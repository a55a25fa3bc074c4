#!/usr/bin/env python3
"""Dev helper (GPU box): where the time of one rank's share of the roadmap kNN goes (parts = 1, 2, 4, 8 simulated on one GPU)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
n = 1 << 20
r = ctx.prm_build("panda", env, n, 2, 42, want_weights=False)
states = torch.from_numpy(r["states"]).cuda()
idx = torch.full((n, 16), -1, dtype=torch.int64, device="cuda")
dst = torch.full((n, 16), float("inf"), dtype=torch.float64, device="cuda")


def timed(fn):
    torch.cuda.synchronize(); ctx.synchronize()
    t = time.perf_counter()
    fn()
    ctx.synchronize()
    return (time.perf_counter() - t) * 1e3


for parts in ([int(sys.argv[1])] if len(sys.argv) > 1 else (1, 2, 4, 8)):
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    t_add = timed(lambda: nn.add_device_f32(states.data_ptr(), n, 10.0))
    t_first = timed(lambda: nn.query_part_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), 0, parts, exclude_self=True))
    t_second = timed(lambda: nn.query_part_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), 0, parts, exclude_self=True))
    t_other = timed(lambda: nn.query_part_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), parts - 1, parts, exclude_self=True))
    print(f"parts={parts}: add {t_add:.1f} ms, first query (index build + scratch alloc) {t_first:.1f} ms, repeat {t_second:.1f} ms, last part {t_other:.1f} ms, "
          f"rows_visited {nn.rows_visited():.3e}", flush=True)
    nn.close()
