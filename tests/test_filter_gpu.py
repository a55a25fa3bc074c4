"""GPU parity tests of b200mp_filter_pointcloud against the golden vectors of the unmodified reference filter
(collision/filter.hh:176-274) and against the CPU restatement (same stable tie order, so outputs must be bit-identical,
including on dense clouds where curve codes collide)."""
import os

import numpy as np
import pytest

from ompl_b200 import synth
from oracle import port
from tests.test_filter_cpu import CASES, GOLDEN, golden_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_golden(ctx, name):
    z = np.load(GOLDEN)
    kept = ctx.filter_pointcloud(*golden_case(z, name))
    np.testing.assert_array_equal(kept, z[name + "_kept"])


@pytest.mark.parametrize("n,min_dist,cull", [(1, 0.01, True), (2, 0.5, True), (777, 0.02, False), (5000, 0.0, True),
                                             (200000, 0.012, True), (1 << 20, 0.012, True)])
def test_dense_cloud_equals_restatement(ctx, n, min_dist, cull):
    pts = synth.shelf_cloud(n, seed=11)
    o = np.array([0, 0, 0.4], np.float32)
    want = port.filter_pointcloud(pts, min_dist, 1.5, o, o - 1.5, o + 1.5, cull)
    got = ctx.filter_pointcloud(pts, min_dist, 1.5, o, o - 1.5, o + 1.5, cull)
    np.testing.assert_array_equal(got, want)


def test_dense_cloud_equals_compiled_reference_with_stable_ties(ctx):
    """The unmodified reference, its absent pdqsort replaced by std::stable_sort (oracle/shim/pdqsort.h), on the C4 cloud: point for point
    the engine's output. With std::sort in its place the outputs differ in the order of equal curve codes only (test_filter_cpu.py)."""
    from tests.conftest import have_reference_build
    if not have_reference_build():
        pytest.skip("oracle/_ref not built")
    from oracle import refvamp
    pts = synth.shelf_cloud(1 << 20, seed=7)
    o = np.array([0, 0, 0.4], np.float32)
    got = ctx.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    want = refvamp.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True, stable_ties=True)
    np.testing.assert_array_equal(got, want)


def test_everything_culled_duplicates_and_empty(ctx):
    o = np.zeros(3, np.float32)
    far = np.full((100, 3), 5.0, np.float32)
    np.testing.assert_array_equal(ctx.filter_pointcloud(far, 0.01, 1.0, o, o - 1, o + 1, True), port.filter_pointcloud(far, 0.01, 1.0, o, o - 1, o + 1, True))
    dup = np.tile(np.array([[0.1, 0.2, 0.3]], np.float32), (3000, 1))
    got = ctx.filter_pointcloud(dup, 0.01, 1.0, o, o - 1, o + 1, True)
    np.testing.assert_array_equal(got, dup[:1])
    assert ctx.filter_pointcloud(np.zeros((0, 3), np.float32), 0.01, 1.0, o, o - 1, o + 1, True).shape == (0, 3)


def test_filtered_cloud_feeds_the_tree_and_validation(ctx):
    """C4 pipeline at reduced size: filter -> tree build -> Fetch validation, decisions equal to the CPU checker's."""
    import ompl_b200
    pts = synth.shelf_cloud(60000, seed=7)
    o = np.array([0, 0, 0.4], np.float32)
    f = ctx.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    env = ompl_b200.Environment(ctx)
    env.add_pointcloud(f, 0.012, 0.24, 0.0025)
    env.commit()
    pe = port.PortEnvironment()
    pe.add_capt(ompl_b200.capt_build(f, 0.012, 0.24, 0.0025))
    q = synth.uniform_states("fetch", 20000, 3)
    np.testing.assert_array_equal(ctx.validate_states("fetch", env, q), pe.validate_states("fetch", q, nthreads=8))
