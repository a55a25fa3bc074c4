#!/usr/bin/env python3
"""Dev helper (GPU box): phase timings of the single-GPU roadmap build at growing sizes (C3 scene = C2 obstacles)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
sizes = [int(a) for a in sys.argv[1:]] or [1 << 14, 1 << 16, 1 << 18, 1 << 20]
for n in sizes:
    for rep in range(2):
        r = ctx.prm_build("panda", env, n, 16, 42, want_weights=False)
    print(json.dumps(r["stats"]), flush=True)
