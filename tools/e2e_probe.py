import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench, ompl_b200
from ompl_b200 import synth
ctx = ompl_b200.Context(0)
scene, q0, q1 = bench.make_inputs(0)
env = synth.fill_env(ompl_b200.Environment(ctx), scene)
h0, h1 = torch.from_numpy(q0).pin_memory(), torch.from_numpy(q1).pin_memory()
hv = torch.zeros(len(q0), dtype=torch.uint8).pin_memory()
for _ in range(3):
    ctx.validate_motions_ptr("panda", env, h0.data_ptr(), h1.data_ptr(), len(q0), hv.data_ptr(), device=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ctx.validate_motions_ptr("panda", env, h0.data_ptr(), h1.data_ptr(), len(q0), hv.data_ptr(), device=False)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"e2e {dt*1e3:.3f} ms/step = {len(q0)/dt:.4e} edges/s")
