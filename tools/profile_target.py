#!/usr/bin/env python3
"""ncu target (GPU box): a handful of validation launches of known composition.
launch 0: warm-up (C2 mixed, 2^14 motions)   launch 1: C2 mixed 2^18   launch 2: C2 all-valid 2^15
launch 3: cage all-valid 2^16               launch 4: C2 states 2^20 uniform random"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


c2 = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
cage = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
q0, q1 = synth.fixed_length_edges("panda", 1 << 18, 1234)
run(c2, q0[:1 << 14], q1[:1 << 14])
v = run(c2, q0, q1)
idx = np.resize(np.nonzero(v == 1)[0], 1 << 15)
run(c2, q0[idx], q1[idx])
vc = ctx.validate_motions("panda", cage, q0, q1)   # host path: launch not on our stream but still profiled (launch 3)
idc = np.resize(np.nonzero(vc == 1)[0], 1 << 16)
run(cage, q0[idc], q1[idc])
q = synth.uniform_states("panda", 1 << 20, 3)
ctx.validate_states("panda", c2, q)
print("valid C2", v.mean(), "valid cage", vc.mean())
