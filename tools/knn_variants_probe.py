#!/usr/bin/env python3
"""Dev helper (GPU box): kNN query time of the warp-centric cell search on the roadmap's states under the development switches
(B200MP_KNN_OLD / _QPT / _CELL_ROWS / _TAU0 / _TAUG), each variant in a fresh process (the switches are read at index build)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
import ompl_b200
from ompl_b200 import synth
ctx = ompl_b200.Context(0)
n = int(os.environ.get("PROBE_N", 1 << 20))
st = np.load("/tmp/knn_probe_states.npy")[:n]
states = torch.from_numpy(st).cuda()
idx = torch.empty((n, 16), dtype=torch.int64, device="cuda"); dst = torch.empty((n, 16), dtype=torch.float64, device="cuda")
nn = ompl_b200.NearestNeighbors(ctx, 7)
nn.add_device_f32(states.data_ptr(), n, 10.0)
ts = []
for rep in range(3):
    ctx.synchronize(); t = time.perf_counter()
    nn.query_device_f32(states.data_ptr(), n, 16, idx.data_ptr(), dst.data_ptr(), self_base=0, exclude_self=True)
    ctx.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
rv = nn.rows_visited()
stt = nn.search_stats()
ref = "/tmp/knn_probe_ref.npz"
ok = None
if os.path.exists(ref):
    z = np.load(ref); ok = bool((z["idx"] == idx.cpu().numpy()).all() and (z["dst"] == dst.cpu().numpy()).all())
else:
    np.savez(ref, idx=idx.cpu().numpy(), dst=dst.cpu().numpy())
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B200MP_KNN")}, "ms": [round(t, 2) for t in ts], "rows_visited": rv,
                  "frac_of_N2_per_3_runs": rv / 3 / n / n if os.environ.get("B200MP_KNN_STAT_PAIRS") else None,
                  "per_query": {k: round(v / 3 / n, 2) for k, v in stt.items()}, "equals_first_variant": ok}), flush=True)
''' % ROOT

sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ompl_b200  # noqa: E402
from ompl_b200 import synth  # noqa: E402

ctx = ompl_b200.Context(0)
env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 500, 500))
r = ctx.prm_build("panda", env, 1 << 20, 2, 42, want_weights=False)
np.save("/tmp/knn_probe_states.npy", r["states"])
del ctx
variants = [dict(B200MP_KNN_OLD="1")] + [json.loads(a) for a in sys.argv[1:]]
for v in variants:
    e = dict(os.environ)
    e.update({k: str(x) for k, x in v.items()})
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-2000:], flush=True)
