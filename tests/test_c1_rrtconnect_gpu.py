"""BASELINE config 1 on the GPU box: a single-query RRT-Connect run in the 14-sphere Panda cage of the reference's demo and
test (demos/vamp/VAMPPlanning.cpp:43-47,87-88; tests/vamp/vamp_test.cpp:29-88 expects an exact solution), once with every
checkMotion answered by the engine through the C ABI (one motion per call, fp64 states as OMPL hands them over) and once by
the CPU checker. The two runs must take the same decisions in the same order, hence build the same trees and path."""
import numpy as np
import pytest

import ompl_b200
from ompl_b200 import synth
from oracle import port
from tests.conftest import have_reference_build
from tests.rrtconnect import solve

pytestmark = pytest.mark.gpu


def checker():
    if have_reference_build():
        from oracle import refvamp
        return synth.fill_env(refvamp.RefEnvironment(), synth.cage_scene())
    return synth.fill_env(port.PortEnvironment(), synth.cage_scene())


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_same_planner_run_with_engine_and_checker(ctx, seed):
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    ce = checker()
    lo, hi = synth.BOUNDS["panda"]
    gpu = lambda a, b: ctx.validate_motions("panda", env, np.asarray(a, np.float64)[None], np.asarray(b, np.float64)[None])[0]  # noqa: E731
    cpu = lambda a, b: ce.validate_motions("panda", a[None].astype(np.float32), b[None].astype(np.float32))[0]  # noqa: E731
    p1, log1, it1 = solve(synth.CAGE_START, synth.CAGE_GOAL, lo, hi, gpu, seed)
    p2, log2, it2 = solve(synth.CAGE_START, synth.CAGE_GOAL, lo, hi, cpu, seed)
    assert p1 is not None and p2 is not None, "RRT-Connect must find an exact solution in the cage (vamp_test.cpp:87)"
    assert log1 == log2 and it1 == it2
    np.testing.assert_array_equal(p1, p2)
    np.testing.assert_array_equal(p1[0], synth.CAGE_START.astype(np.float64))
    np.testing.assert_array_equal(p1[-1], synth.CAGE_GOAL.astype(np.float64))
    # every edge of the solution is valid for the checker, and the straight line is not (known answer of the demo)
    assert ce.validate_motions("panda", p1[:-1].astype(np.float32), p1[1:].astype(np.float32)).all()
    assert not cpu(synth.CAGE_START.astype(np.float64), synth.CAGE_GOAL.astype(np.float64))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_same_planner_run_through_the_resident_kernel(ctx, seed):
    """The same run with every checkMotion answered by a b200mp_session (no launch per call): identical decisions again."""
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    ce = checker()
    lo, hi = synth.BOUNDS["panda"]
    s = ompl_b200.Session(ctx, "panda", env)
    gpu = lambda a, b: s.check_motion(np.asarray(a, np.float64), np.asarray(b, np.float64))  # noqa: E731
    cpu = lambda a, b: ce.validate_motions("panda", a[None].astype(np.float32), b[None].astype(np.float32))[0]  # noqa: E731
    p1, log1, it1 = solve(synth.CAGE_START, synth.CAGE_GOAL, lo, hi, gpu, seed)
    p2, log2, it2 = solve(synth.CAGE_START, synth.CAGE_GOAL, lo, hi, cpu, seed)
    assert p1 is not None and log1 == log2 and it1 == it2
    np.testing.assert_array_equal(p1, p2)
    s.close()
