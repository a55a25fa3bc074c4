#!/usr/bin/env python3
"""ncu target (GPU box): C4 -- Fetch motions against the filtered bookshelf cloud. Launch #0 warm-up, #1 2^18 mixed motions."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
o = np.array([0, 0, 0.4], np.float32)
pts = synth.shelf_cloud(1 << 20, seed=7)
kept = ctx.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
env = ompl_b200.Environment(ctx)
env.add_pointcloud(kept, 0.012, 0.24, 0.0025)
env.commit()
q0, q1 = synth.fixed_length_edges("fetch", 1 << 18, 4321)
d0, d1 = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
out = torch.zeros(len(q0), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
for m in (1 << 12, len(q0)):
    ctx.validate_motions_ptr("fetch", env, d0.data_ptr(), d1.data_ptr(), m, out.data_ptr(), stream=stream.cuda_stream)
    stream.synchronize()
print(float(out.float().mean()))
