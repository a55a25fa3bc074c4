#!/usr/bin/env python3
"""Dev helper (GPU box): the state kernel on 2^23 uniform Panda candidates of the C2 scene (what the roadmap sampler validates), device timed."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
m = 1 << 23
q = torch.from_numpy(synth.uniform_states("panda", m, 5)).cuda()
out = torch.zeros(m, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        ctx.validate_states_ptr("panda", env, q.data_ptr(), m, out.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
    stream.synchronize()
    ms.append(e0.elapsed_time(e1))
print(json.dumps({"sift": os.environ.get("B200MP_SIFT", "1"), "states": m, "ms": float(np.median(ms[1:])), "states_per_s": m / (np.median(ms[1:]) * 1e-3),
                  "valid_fraction": float(out.float().mean())}))
