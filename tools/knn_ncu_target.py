#!/usr/bin/env python3
"""ncu target (GPU box): one kNN batch query (2^20 x 2^20, k = 16) on the roadmap's states; states cached in /tmp by knn_variants_probe.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402

ctx = ompl_b200.Context(0)
st = np.load("/tmp/knn_probe_states.npy")
n = len(st)
states = torch.from_numpy(st).cuda()
idx = torch.empty((n, 16), dtype=torch.int64, device="cuda")
dst = torch.empty((n, 16), dtype=torch.float64, device="cuda")
nn = ompl_b200.NearestNeighbors(ctx, 7)
nn.add_device_f32(states.data_ptr(), n, 10.0)
for rep in range(2):
    nn.query_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), self_base=0, exclude_self=True)
ctx.synchronize()
