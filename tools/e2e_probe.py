#!/usr/bin/env python3
"""Dev helper (GPU box): end-to-end time of b200mp_validate_motions with pinned host buffers on the C2 batch (fp32 and fp64)."""
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
env = synth.fill_env(ompl_b200.Environment(ctx), scene)
M = len(q0)
for name, a, b in (("f32", q0, q1), ("f64", q0.astype(np.float64), q1.astype(np.float64))):
    h0, h1 = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    hv = torch.zeros(M, dtype=torch.uint8).pin_memory()
    L = ctx._L
    fn = L.b200mp_validate_motions if name == "f32" else L.b200mp_validate_motions_f64
    for _ in range(2):
        fn(ctx._h, 0, env._h, h0.data_ptr(), h1.data_ptr(), M, 0, hv.data_ptr())
    t0 = time.perf_counter()
    for _ in range(5):
        fn(ctx._h, 0, env._h, h0.data_ptr(), h1.data_ptr(), M, 0, hv.data_ptr())
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt * 1e3:.2f} ms per 2^20 motions = {M / dt:.3e} edges/s, valid {int(hv.sum())}", flush=True)
