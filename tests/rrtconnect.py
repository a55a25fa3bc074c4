"""TEST INFRASTRUCTURE: a small single-query bidirectional planner with the control flow of OMPL's RRTConnect
(src/ompl/geometric/planners/rrt/src/RRTConnect.cpp:112-300: growTree = nearest, clip to maxDistance, one checkMotion;
extend tree A towards a sample, then greedily connect tree B to the new state; swap). It exists so that BASELINE config 1
(Panda in the 14-sphere cage, demos/vamp/VAMPPlanning.cpp, tests/vamp/vamp_test.cpp:29-88) can be run against ANY
MotionValidator-like callable: the engine behind its C ABI and the CPU checker must then produce the same trees, because
every decision they hand back is the same. Linear nearest neighbour (the reference's ground truth), seeded numpy RNG."""
from __future__ import annotations

import numpy as np


class Tree:
    def __init__(self, root):
        self.q = [np.asarray(root, np.float64)]
        self.parent = [-1]

    def nearest(self, x):
        d = np.linalg.norm(np.asarray(self.q) - x, axis=1)
        return int(np.argmin(d))

    def add(self, x, parent):
        self.q.append(x)
        self.parent.append(parent)
        return len(self.q) - 1

    def path_to_root(self, i):
        out = []
        while i >= 0:
            out.append(self.q[i])
            i = self.parent[i]
        return out


TRAPPED, ADVANCED, REACHED = 0, 1, 2


def grow(tree, target, max_distance, check_motion, log):
    i = tree.nearest(target)
    near = tree.q[i]
    d = float(np.linalg.norm(target - near))
    reach = d <= max_distance
    new = target if reach else near + (target - near) * (max_distance / d)   # RealVectorStateSpace::interpolate
    ok = bool(check_motion(near, new))
    log.append(ok)
    if not ok:
        return TRAPPED, -1
    return (REACHED if reach else ADVANCED), tree.add(new, i)


def solve(start, goal, lo, hi, check_motion, seed, max_iterations=20000, range_fraction=0.2):
    """-> (path [n][dim] or None, decisions log, iterations)"""
    rng = np.random.default_rng(seed)
    max_distance = range_fraction * float(np.linalg.norm(np.asarray(hi, np.float64) - np.asarray(lo, np.float64)))  # SelfConfig::configurePlannerRange
    ta, tb = Tree(start), Tree(goal)
    a_is_start = True
    log = []
    for it in range(max_iterations):
        sample = rng.uniform(lo, hi)
        st, ia = grow(ta, sample, max_distance, check_motion, log)
        if st != TRAPPED:
            target = ta.q[ia]
            while True:
                st2, ib = grow(tb, target, max_distance, check_motion, log)
                if st2 != ADVANCED:
                    break
            if st2 == REACHED:
                pa, pb = ta.path_to_root(ia)[::-1], tb.path_to_root(ib)
                path = pa + pb[1:]
                if not a_is_start:
                    path = path[::-1]
                return np.asarray(path), log, it + 1
        ta, tb = tb, ta
        a_is_start = not a_is_start
    return None, log, max_iterations
