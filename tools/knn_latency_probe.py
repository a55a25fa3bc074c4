#!/usr/bin/env python3
"""GPU-box probe: latency of the one-at-a-time NearestNeighbors calls an RRT makes (add one state, nearest of one query)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ompl_b200
ctx = ompl_b200.Context(0)
rng = np.random.default_rng(1)
for n0 in (1000, 100000):
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.add(rng.uniform(-3, 3, (n0, 7)))
    nn.nearest_k(rng.uniform(-3, 3, (1, 7)), 1)
    t0 = time.perf_counter()
    for i in range(500):
        nn.nearest_k(rng.uniform(-3, 3, (1, 7)), 1)
    t1 = time.perf_counter()
    for i in range(500):
        nn.add(rng.uniform(-3, 3, (1, 7)))
        nn.nearest_k(rng.uniform(-3, 3, (1, 7)), 1)
    t2 = time.perf_counter()
    print(f"n={n0}: nearest {1e6 * (t1 - t0) / 500:.1f} us/call; add+nearest {1e6 * (t2 - t1) / 500:.1f} us/pair", flush=True)
