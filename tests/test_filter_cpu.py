"""CPU tests of the point-cloud filter restatement (oracle/oracle_filter.cpp) against golden vectors generated from the
unmodified reference (vamp::collision::filter_pointcloud, collision/filter.hh:176-274) and, where oracle/_ref exists,
against the compiled reference live."""
import os

import numpy as np
import pytest

from ompl_b200 import synth
from oracle import port
from tests.conftest import have_reference_build

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "filter.npz")
CASES = ["ball", "nocull", "coarse", "first_point_culled"]


def golden_case(z, name):
    md, mr, ox, oy, oz, cull = (float(v) for v in z[name + "_params"])
    o = np.array([ox, oy, oz], np.float32)
    return z[name + "_points"], md, mr, o, o - np.float32(mr), o + np.float32(mr), bool(cull)


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_golden(name):
    z = np.load(GOLDEN)
    kept = port.filter_pointcloud(*golden_case(z, name))
    np.testing.assert_array_equal(kept, z[name + "_kept"])


def test_reference_quirk_first_point_reappears():
    """The reference never shrinks its work list after culling (filter.hh:198-219), so point 0 is emitted even when it
    lies outside the range ball; the golden vector shows it and the restatement reproduces it."""
    z = np.load(GOLDEN)
    kept = z["first_point_culled_kept"]
    assert (kept == np.float32(9.0)).all(axis=1).sum() == 1


def test_contract_properties_on_a_dense_cloud():
    pts = synth.shelf_cloud(100000, seed=7)
    o = np.array([0, 0, 0.4], np.float32)
    kept = port.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    assert 0 < len(kept) < len(pts) // 3
    # output is a subset of the input, without consecutive near-duplicates in the final curve order
    src = set(map(tuple, pts))
    assert all(tuple(p) in src for p in kept)
    d = np.linalg.norm(kept[1:].astype(np.float64) - kept[:-1].astype(np.float64), axis=1)
    assert (d > 0.012 * (1 - 1e-6)).all()
    # idempotent up to the curve's resolution: a second pass removes almost nothing
    again = port.filter_pointcloud(kept, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    assert len(again) > 0.9 * len(kept)


def test_empty_and_single():
    o = np.zeros(3, np.float32)
    assert port.filter_pointcloud(np.zeros((0, 3), np.float32), 0.01, 1.0, o, o - 1, o + 1, True).shape == (0, 3)
    one = np.array([[0.1, 0.2, 0.3]], np.float32)
    np.testing.assert_array_equal(port.filter_pointcloud(one, 0.01, 1.0, o, o - 1, o + 1, True), one)


@pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref not built")
def test_restatement_matches_compiled_reference_live():
    from oracle import refvamp
    rng = np.random.default_rng(5)
    o = np.array([0.1, -0.1, 0.4], np.float32)
    for n, md, cull in [(1000, 0.02, True), (20000, 0.012, True), (20000, 0.012, False), (50000, 0.03, True)]:
        pts = (rng.uniform(-1.6, 1.6, (n, 3)) + o).astype(np.float32)
        a = refvamp.filter_pointcloud(pts, md, 1.5, o, o - 1.5, o + 1.5, cull)
        b = port.filter_pointcloud(pts, md, 1.5, o, o - 1.5, o + 1.5, cull)
        np.testing.assert_array_equal(a, b)
    # dense surface cloud: equal codes occur, the reference's tie order is unspecified -> same size within 0.5 %, > 85 % common points
    pts = synth.shelf_cloud(200000, seed=7)
    a = refvamp.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    b = port.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    assert abs(len(a) - len(b)) < 0.005 * len(a)
    assert len(set(map(tuple, a)) & set(map(tuple, b))) > 0.85 * len(a)
    # ... and tie order is ALL they differ by: the same reference code with its (absent) pdqsort replaced by a stable sort instead of
    # std::sort gives the restatement's output point for point
    c = refvamp.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True, stable_ties=True)
    np.testing.assert_array_equal(c, b)
