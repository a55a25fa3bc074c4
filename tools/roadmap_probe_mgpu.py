#!/usr/bin/env python3
"""Dev helper (GPU box, under torchrun): repeated multi-GPU roadmap builds with per-phase timings on every rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402
from ompl_b200.roadmap import EngineBackend, RoadmapBuilder  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ompl_b200.Context(local)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
b = RoadmapBuilder(EngineBackend(ctx, env, "panda"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for rep in range(3):
    if world > 1:
        dist.barrier()
    rm = b.build(n, 16, 42)
    s = rm.stats
    print(rank, rep, json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in s.items()}), flush=True)
    del rm
if world > 1:
    dist.destroy_process_group()
