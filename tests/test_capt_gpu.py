"""GPU parity tests of point-cloud (CAPT) collision checking: Fetch / Panda against a synthetic cloud, the engine's own
tree build + device query versus the unmodified reference (which builds its own tree from the same points) and versus
the CPU checker fed the engine-built arrays."""
import numpy as np
import pytest

import ompl_b200
from ompl_b200 import synth
from oracle import port
from tests.conftest import have_reference_build
from tests.util import assert_validity_parity

pytestmark = pytest.mark.gpu

FETCH_R = (0.012, 0.24, 0.0025)    # robots/fetch.hh:17 min/max sphere radius, point radius of the demos
PANDA_R = (0.012, 0.18, 0.0025)


def _envs(ctx, pts, radii, scene=None):
    env = ompl_b200.Environment(ctx)
    env.add_pointcloud(pts, *radii)
    pe = port.PortEnvironment()
    pe.add_capt(ompl_b200.capt_build(pts, *radii))
    if scene:
        synth.fill_env(env, scene)
        synth.fill_env(pe, scene)
    else:
        env.commit()
    return env, pe


@pytest.mark.parametrize("robot,radii,n_pts", [("fetch", FETCH_R, 6000), ("panda", PANDA_R, 4000)])
def test_pointcloud_validity_vs_checker_and_reference(ctx, robot, radii, n_pts):
    pts = synth.shelf_cloud(n_pts, seed=7, robot=robot)
    env, pe = _envs(ctx, pts, radii)
    q = synth.uniform_states(robot, 20000, 5)
    want = pe.validate_states(robot, q, nthreads=8)
    got = ctx.validate_states(robot, env, q)
    assert 0.02 < want.mean() < 0.98
    assert_validity_parity(got, want, lambda b: pe.edge_min_clearance(robot, q[b], q[b], resolution=1), "cloud states")
    starts = q[want == 1][:3000]
    q0, q1 = synth.local_edges(robot, starts, 21, 0.05, 2.0)
    wantm = pe.validate_motions(robot, q0, q1, nthreads=8)
    gotm = ctx.validate_motions(robot, env, q0, q1)
    assert 0.02 < wantm.mean() < 0.98
    assert_validity_parity(gotm, wantm, lambda b: pe.edge_min_clearance(robot, q0[b], q1[b]), "cloud motions")
    if have_reference_build():
        from oracle import refvamp
        re_ = refvamp.RefEnvironment()
        re_.add_pointcloud(pts, *radii)
        assert_validity_parity(got, re_.validate_states(robot, q, nthreads=8),
                               lambda b: pe.edge_min_clearance(robot, q[b], q[b], resolution=1), "cloud states vs reference")
        assert_validity_parity(gotm, re_.validate_motions(robot, q0, q1, nthreads=8),
                               lambda b: pe.edge_min_clearance(robot, q0[b], q1[b]), "cloud motions vs reference")


def test_pointcloud_with_primitives_and_golden(ctx):
    z = np.load(synth_golden("fetch_cloud"))
    pts, radii = z["points"], tuple(float(v) for v in z["radii"])
    scene = {k[6:]: z[k] for k in z.files if k.startswith("scene_")}
    env, pe = _envs(ctx, pts, radii, scene)
    got = ctx.validate_states("fetch", env, z["states"])
    assert_validity_parity(got, z["states_valid"], lambda b: pe.edge_min_clearance("fetch", z["states"][b], z["states"][b], resolution=1),
                           "golden cloud states")
    gotm = ctx.validate_motions("fetch", env, z["q0"], z["q1"])
    assert_validity_parity(gotm, z["motions_valid"], lambda b: pe.edge_min_clearance("fetch", z["q0"][b], z["q1"][b]), "golden cloud motions")


def synth_golden(name):
    import os
    return os.path.join(os.path.dirname(__file__), "golden", name + ".npz")


def test_empty_and_single_point_cloud(ctx):
    for pts in (np.zeros((0, 3), np.float32), np.array([[0.4, 0.0, 0.5]], np.float32), np.array([[0.4, 0.0, 0.5]] * 3, np.float32)):
        env, pe = _envs(ctx, pts, PANDA_R)
        q = synth.uniform_states("panda", 3000, 2)
        np.testing.assert_array_equal(ctx.validate_states("panda", env, q), pe.validate_states("panda", q, nthreads=4))


# The device-built tree lists, for every leaf, EVERY point within r_max + r_point of the leaf's cell, so it answers "is any point
# within r + r_point of c" exactly for r <= r_max. The reference's build hands candidate points down its recursion through slab
# tests with an early break (capt.hh:123-285; reproduced by the host build, whose arrays equal the reference's) and can drop a
# point a leaf needs, so the reference occasionally misses a collision the exact tree reports. Hence: decisions of the device
# tree == brute force over the raw points; and device-valid implies host-valid, with a handful of extra collisions.
FETCH_EXACT = (0.012, 0.35, 0.0025)   # covers Fetch's largest bounding sphere (0.34), not just its largest fine sphere (0.24)


def _brute_force_valid(robot, q, pts, r_point):
    """state valid <=> no fine sphere within r + r_point of any point (float64)."""
    sph = port.sphere_fk(robot, q).astype(np.float64)          # [m][n_spheres][4]
    out = np.ones(len(q), np.uint8)
    p = pts.astype(np.float64)
    for i in range(len(q)):
        d2 = ((sph[i, :, None, :3] - p[None, :, :]) ** 2).sum(axis=2)
        rr = (sph[i, :, 3] + r_point)[:, None] ** 2
        if (d2 <= rr).any():
            out[i] = 0
    return out


@pytest.mark.parametrize("robot,radii,n_pts", [("fetch", FETCH_EXACT, 6000), ("panda", PANDA_R, 4000), ("fetch", FETCH_EXACT, 1), ("fetch", FETCH_EXACT, 33)])
def test_device_built_tree_is_exact_and_at_least_as_strict_as_the_reference_build(ctx, robot, radii, n_pts):
    pts = synth.shelf_cloud(n_pts, seed=9, robot=robot)
    env_host, pe = _envs(ctx, pts, radii)
    env_dev = ompl_b200.Environment(ctx)
    env_dev.add_pointcloud_device(pts, *radii)
    env_dev.commit()
    q = synth.uniform_states(robot, 30000, 15)
    host, dev = ctx.validate_states(robot, env_host, q), ctx.validate_states(robot, env_dev, q)
    assert not (dev & (1 - host)).any(), "a state the exact tree accepts must be accepted by the reference-identical tree"
    assert (host != dev).mean() < 2e-3
    # exactness on a sample that contains every disagreement: brute force over the raw points (self-collision filtered out by
    # taking states that are valid in an empty environment)
    empty = ompl_b200.Environment(ctx)
    empty.commit()
    free = ctx.validate_states(robot, empty, q) == 1
    pick = np.unique(np.concatenate([np.nonzero((host != dev) & free)[0], np.nonzero(free)[0][:300]]))
    want = _brute_force_valid(robot, q[pick], pts, radii[2])
    got = dev[pick]
    bad = np.nonzero(got != want)[0]
    if bad.size:   # float32 kernel vs float64 brute force: only contacts within 1e-5 m may flip
        clr = pe.edge_min_clearance(robot, q[pick][bad], q[pick][bad], resolution=1)
        assert (np.abs(clr) < 1e-5).all()
    starts = q[host == 1][:4000]
    if len(starts):
        q0, q1 = synth.local_edges(robot, starts, 25, 0.05, 2.0)
        mh, md = ctx.validate_motions(robot, env_host, q0, q1), ctx.validate_motions(robot, env_dev, q0, q1)
        assert not (md & (1 - mh)).any() and (mh != md).mean() < 5e-3


def test_device_built_tree_on_the_filtered_c4_cloud(ctx):
    pts = synth.shelf_cloud(300000, seed=7)
    o = np.array([0, 0, 0.4], np.float32)
    kept = ctx.filter_pointcloud(pts, 0.012, 1.5, o, o - 1.5, o + 1.5, True)
    eh, ed = ompl_b200.Environment(ctx), ompl_b200.Environment(ctx)
    eh.add_pointcloud(kept, *FETCH_R)
    ed.add_pointcloud_device(kept, *FETCH_R)
    eh.commit()
    ed.commit()
    q0, q1 = synth.fixed_length_edges("fetch", 1 << 15, 77)
    mh, md = ctx.validate_motions("fetch", eh, q0, q1), ctx.validate_motions("fetch", ed, q0, q1)
    assert (mh != md).mean() < 5e-3 and 0.05 < mh.mean() < 0.95
