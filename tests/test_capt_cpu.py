"""CPU tests of the point-cloud tree: the engine's host build (b200mp_capt_build) must produce the arrays the
unmodified reference builds (vamp::collision::CAPT, collision/capt.hh:104-365) -- golden fixture generated from the
reference, plus live comparison where oracle/_ref exists -- and the CPU checker's tree query on those arrays must
reproduce the reference's validity decisions (pins the oracle's CAPT path)."""
import os

import numpy as np
import pytest

import ompl_b200
from ompl_b200 import synth
from oracle import port
from tests.conftest import have_reference_build
from tests.util import assert_validity_parity

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fetch_cloud.npz")


def test_build_matches_golden_arrays():
    z = np.load(GOLDEN)
    t = ompl_b200.capt_build(z["points"], *[float(v) for v in z["radii"]])
    assert t["nlog2"] == int(z["capt_nlog2"])
    np.testing.assert_array_equal(t["tests"], z["capt_tests"])
    np.testing.assert_array_equal(t["aff_starts"], z["capt_aff_starts"])
    np.testing.assert_array_equal(t["aabbs"], z["capt_aabbs"])
    np.testing.assert_array_equal(t["top"], z["capt_top"])
    np.testing.assert_array_equal(np.where(np.isfinite(t["aff"]), t["aff"], 0).sum(axis=1, dtype=np.float64), z["capt_aff_sum"])


def test_checker_query_on_engine_built_tree_matches_golden_validity():
    z = np.load(GOLDEN)
    pe = port.PortEnvironment()
    pe.add_capt(ompl_b200.capt_build(z["points"], *[float(v) for v in z["radii"]]))
    synth.fill_env(pe, {k[6:]: z[k] for k in z.files if k.startswith("scene_")})
    sv = pe.validate_states("fetch", z["states"], nthreads=4)
    assert_validity_parity(sv, z["states_valid"], lambda b: pe.edge_min_clearance("fetch", z["states"][b], z["states"][b], resolution=1),
                           "cloud states")
    mv = pe.validate_motions("fetch", z["q0"], z["q1"], nthreads=4)
    assert_validity_parity(mv, z["motions_valid"], lambda b: pe.edge_min_clearance("fetch", z["q0"][b], z["q1"][b]), "cloud motions")


def test_tree_structure_invariants():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 17, 64, 1000):
        pts = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        t = ompl_b200.capt_build(pts, 0.02, 0.1, 0.005)
        leaves = 1 << t["nlog2"]
        assert leaves >= max(n, 1) and (leaves == 1 or leaves // 2 < n)
        assert len(t["tests"]) == leaves - 1 and t["aabbs"].shape == (leaves, 6) and len(t["aff_starts"]) == leaves + 1
        assert (np.diff(t["aff_starts"].astype(np.int64)) >= 0).all() and t["aff_starts"][-1] * 8 == t["aff"].shape[1]
        fin = np.isfinite(t["aff"][0])
        # every input point is some leaf's representative: it appears in the affordance buffer
        have = {tuple(p) for p in t["aff"][:, fin].T.tolist()}
        assert all(tuple(p) in have for p in pts.tolist())


@pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref not built")
def test_build_matches_reference_live():
    from oracle import refvamp
    rng = np.random.default_rng(11)
    for n, radii in ((2, (0.012, 0.24, 0.0025)), (77, (0.02, 0.08, 0.01)), (4096, (0.012, 0.24, 0.0025)), (5000, (0.02, 0.08, 0.01))):
        pts = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        if n > 100:
            pts[:50] = pts[50:100]    # duplicate points
        mine = ompl_b200.capt_build(pts, *radii)
        e = refvamp.RefEnvironment()
        e.add_pointcloud(pts, *radii)
        ref = e.capt_export()
        assert mine["nlog2"] == ref["nlog2"]
        for k in ("tests", "aabbs", "aff_starts", "aff", "top"):
            np.testing.assert_array_equal(mine[k], ref[k])
        # scalar query of the reference on random spheres == the checker's query on the engine-built arrays
        pe = port.PortEnvironment()
        pe.add_capt(mine)
