#!/usr/bin/env python3
"""Generate tests/golden/filter.npz from the UNMODIFIED reference filter compiled in oracle/_ref
(vamp::collision::filter_pointcloud, collision/filter.hh:176-274; run where /root/reference exists).
Cases are sparse enough that no two entries share a Morton code, so the reference's unstable sort has a unique answer;
the inputs are stored with the answers."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refvamp  # noqa: E402

rng = np.random.default_rng(20261017)
out = {}
cases = {
    # name: (points, min_dist, max_range, origin, cull)
    "ball": (rng.uniform(-1.7, 1.7, (6000, 3)) + [0, 0, 0.4], 0.012, 1.5, [0.0, 0.0, 0.4], True),
    "nocull": (rng.uniform(-1.2, 1.2, (3000, 3)), 0.05, 1.4, [0.0, 0.0, 0.0], False),
    "coarse": (rng.uniform(-1.0, 1.0, (4000, 3)) + [0.3, -0.2, 0.5], 0.15, 1.2, [0.3, -0.2, 0.5], True),
    "first_point_culled": (np.concatenate([[[9.0, 9.0, 9.0]], rng.uniform(-1.0, 1.0, (2000, 3))]), 0.03, 1.4, [0.0, 0.0, 0.0], True),
}
for name, (pts, md, mr, origin, cull) in cases.items():
    pts = pts.astype(np.float32)
    origin = np.asarray(origin, np.float32)
    kept = refvamp.filter_pointcloud(pts, md, mr, origin, origin - mr, origin + mr, cull)
    out[name + "_points"] = pts
    out[name + "_params"] = np.array([md, mr, *origin, float(cull)], np.float32)
    out[name + "_kept"] = kept
    print(name, len(pts), "->", len(kept))
np.savez_compressed(os.path.join(HERE, "filter.npz"), **out)
