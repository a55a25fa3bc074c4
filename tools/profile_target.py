#!/usr/bin/env python3
"""ncu target (GPU box): the bench's own workloads as a short, fixed launch sequence.
  validate_motions launches: #0 warm-up (2^14 mixed), #1 C2 mixed 2^20 (the bench step), #2 C2 all-valid 2^17 (the roofline launch),
                             #3 the same all-valid motions through the ordered scan (candidate grid off: the dense variant)
  then one single-GPU roadmap build (n = 2^18, k = 16): sample / validate_states / compaction / kNN index + cells / edges / validate
Prints the facts profiles/validate_flops.json is keyed on (motions, unique_valid)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
stream = torch.cuda.Stream()


def run(env, q0, q1):
    d0, d1 = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    out = torch.zeros(len(q0), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.validate_motions_ptr("panda", env, d0.data_ptr(), d1.data_ptr(), len(q0), out.data_ptr(), stream=stream.cuda_stream)
    stream.synchronize()
    return out.cpu().numpy()


scene, q0, q1 = bench.make_inputs(0)
env = synth.fill_env(ompl_b200.Environment(ctx), scene)
run(env, q0[:1 << 14], q1[:1 << 14])
v = run(env, q0, q1)
valid_idx = np.nonzero(v == 1)[0]
rep = np.resize(valid_idx, 1 << 17)
va = run(env, q0[rep], q1[rep])
assert va.all()
env_dense = synth.fill_env(ompl_b200.Environment(ctx), scene, sort=True, grid=False)
vd = run(env_dense, q0[rep], q1[rep])
assert vd.all()
if "--no-roadmap" not in sys.argv:
    r = ctx.prm_build("panda", env, 1 << 18, 16, 42, want_weights=False)
    print(json.dumps(r["stats"]))
print(json.dumps({"motions": int(len(rep)), "unique_valid": int(len(valid_idx)), "mixed_valid_fraction": float(v.mean()),
                  "kernel_sha256": bench.kernel_hash()}))
