"""CPU tests: pin the oracle (CPU restatement) against the golden vectors generated from the UNMODIFIED
reference (tests/golden/make_golden.py) and, where oracle/_ref was built, against the reference itself."""
import os

import numpy as np
import pytest

from ompl_b200 import synth
from oracle import port
from tests.util import assert_validity_parity, load_golden
from tests.conftest import have_reference_build

CASES = ["panda_cage", "panda_clutter", "panda_clutter_unsorted", "fetch_clutter", "ur5_clutter", "baxter_clutter", "panda_terrain", "fetch_terrain_clutter"]


@pytest.mark.parametrize("name", CASES)
def test_port_matches_golden_validity(name):
    z, scene, robot, is_sorted = load_golden(name)
    env = synth.fill_env(port.PortEnvironment(), scene, sort=is_sorted)
    for kind in range(5):
        np.testing.assert_allclose(env.min_distances(kind), z[f"min_d_{kind}"], rtol=0, atol=2.5e-7)
    sv = env.validate_states(robot, z["states"], nthreads=4)
    assert_validity_parity(sv, z["states_valid"],
                           lambda b: env.edge_min_clearance(robot, z["states"][b], z["states"][b], resolution=1), name + " states")
    mv = env.validate_motions(robot, z["q0"], z["q1"], nthreads=4)
    assert_validity_parity(mv, z["motions_valid"], lambda b: env.edge_min_clearance(robot, z["q0"][b], z["q1"][b]), name + " motions")


ATTACH_CASES = ["panda", "panda_unsorted", "fetch", "ur5", "baxter"]


def attach_scene(tag):
    from tests.golden.make_golden import ATTACH_SCENES
    return synth.clutter_scene(**ATTACH_SCENES[tag]), tag.split("_")[0], not tag.endswith("unsorted")


@pytest.mark.parametrize("tag", ATTACH_CASES)
def test_port_matches_golden_attachments(tag):
    """Robot::fkcc_attach (collision/attachments.hh:13-56, robots/<robot>.hh): golden answers of the unmodified reference."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "attach.npz"))
    scene, robot, is_sorted = attach_scene(tag)
    env = synth.fill_env(port.PortEnvironment(), scene, sort=is_sorted)
    q, q0, q1 = z[tag + "_states"], z[tag + "_q0"], z[tag + "_q1"]
    clr_s = lambda b: env.edge_min_clearance(robot, q[b], q[b], resolution=1)
    assert_validity_parity(env.validate_states(robot, q, nthreads=4), z[tag + "_states_valid_bare"], clr_s, tag + " bare states")
    env.set_attachment(z[tag + "_tf"], z[tag + "_spheres"])
    assert_validity_parity(env.validate_states(robot, q, nthreads=4), z[tag + "_states_valid"], clr_s, tag + " states")
    assert (z[tag + "_states_valid"] <= z[tag + "_states_valid_bare"]).all() and (z[tag + "_states_valid"] < z[tag + "_states_valid_bare"]).any()
    assert_validity_parity(env.validate_motions(robot, q0, q1, nthreads=4), z[tag + "_motions_valid"],
                           lambda b: env.edge_min_clearance(robot, q0[b], q1[b]), tag + " motions")
    # the end-effector frame the attachment is posed in
    # (Robot::eefk is the reference's SCALAR trace: libm sin / cos, where fkcc_attach and the restatement use the 4.2e-6 polynomials)
    assert np.abs(port.eefk(robot, q[:64]) - z[tag + "_eefk"]).max() < 3e-5
    env.clear_attachment()
    assert_validity_parity(env.validate_states(robot, q, nthreads=4), z[tag + "_states_valid_bare"], clr_s, tag + " cleared")


def test_demo_known_answers():
    # demos/vamp/VAMPPlanning.cpp:43-47,87-88: start and goal are valid in the 14-sphere cage, the straight edge is not
    env = synth.fill_env(port.PortEnvironment(), synth.cage_scene())
    assert list(env.validate_states("panda", [synth.CAGE_START, synth.CAGE_GOAL])) == [1, 1]
    assert list(env.validate_motions("panda", [synth.CAGE_START], [synth.CAGE_GOAL])) == [0]


@pytest.mark.parametrize("robot", ["panda", "fetch", "ur5", "baxter"])
def test_port_sphere_fk_matches_reference_trace(robot):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{robot}_sphere_fk.npz"))
    got = port.sphere_fk(robot, z["q"])
    assert np.abs(got - z["spheres"]).max() < 1e-6  # fp32 rounding of two different operation orders
    np.testing.assert_array_equal(got[..., 3], z["spheres"][..., 3])


def test_trig_error_profile():
    # the reference's vector sin/cos are NOT libm-accurate: rounded octant index (avx.hh:484) -> ~4.2e-6
    x = np.linspace(-3.2, 3.2, 200001).astype(np.float32)
    s, c = port.trig(x)
    es = np.abs(s - np.sin(x.astype(np.float64))).max()
    ec = np.abs(c - np.cos(x.astype(np.float64))).max()
    assert 3e-6 < es < 5e-6 and 3e-6 < ec < 5e-6


def test_bounds_match_bridge():
    for robot in ("panda", "fetch"):
        lo, hi = port.bounds(robot)
        np.testing.assert_array_equal(lo, synth.BOUNDS[robot][0])
        np.testing.assert_array_equal(hi, synth.BOUNDS[robot][1])


def test_edge_cases():
    empty = port.PortEnvironment()
    q = synth.uniform_states("panda", 64, 1)
    # no obstacles: only self-collision decides; zero-length motion == state check
    sv = empty.validate_states("panda", q)
    mv = empty.validate_motions("panda", q, q)
    np.testing.assert_array_equal(sv, mv)
    assert empty.validate_motions("panda", q[:0], q[:0]).shape == (0,)


def test_w_ref_counter_monotone():
    env = synth.fill_env(port.PortEnvironment(), synth.clutter_scene(5, 50, 50))
    q0, q1 = synth.fixed_length_edges("panda", 64, 3)
    v, fl = env.validate_motions("panda", q0, q1, want_flops=True)
    assert (fl > 0).all()
    # a valid 64-state edge runs all 8 rakes: at least 8 x (FK + 21 coarse self tests) x 8 lanes
    if v.any():
        assert fl[v == 1].min() >= 8 * 8 * (1595 + 21 * 11)


# ---------------------------------------------------------------- kNN
def test_port_knn_matches_reference_linear():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "knn_linear.npz"))
    for tag in ("d7", "d8", "d20", "d3"):
        db, qs, k = z[tag + "_db"], z[tag + "_q"], int(z[tag + "_k"])
        idx, dist = port.knn(db, qs, k)
        cnt = z[tag + "_cnt"]
        assert ((idx >= 0).sum(axis=1) == cnt).all()  # k > n returns all n
        # tests/datastructures/nearestneighbors.cpp:149-158 compares the i-th DISTANCE (eps 1e-6), not identities
        np.testing.assert_allclose(dist[np.isfinite(dist)], z[tag + "_dist"][np.isfinite(z[tag + "_dist"])], rtol=0, atol=1e-12)
        same = (idx == z[tag + "_idx"]) | (np.abs(dist - z[tag + "_dist"]) < 1e-12)
        assert same.all()
        # nearestR(2.0): same counts
        got_cnt = np.array([len(port.knn_radius(db, q, 2.0)[0]) for q in qs])
        np.testing.assert_array_equal(got_cnt, z[tag + "_rcnt"])


@pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_compiled_reference_bulk():
    from oracle import refvamp
    scene = synth.clutter_scene(21, 150, 150, n_capsules=30, n_zcuboids=30, n_zcapsules=30)
    r = synth.fill_env(refvamp.RefEnvironment(), scene)
    p = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states("panda", 20000, 4)
    a, b = r.validate_states("panda", q, nthreads=4), p.validate_states("panda", q, nthreads=4)
    assert_validity_parity(b, a, lambda i: p.edge_min_clearance("panda", q[i], q[i], resolution=1), "bulk states")
    q0, q1 = synth.local_edges("panda", q[a == 1][:3000], 6, 0.05, 2.0)
    a, b = r.validate_motions("panda", q0, q1, nthreads=4), p.validate_motions("panda", q0, q1, nthreads=4)
    assert_validity_parity(b, a, lambda i: p.edge_min_clearance("panda", q0[i], q1[i]), "bulk motions")
