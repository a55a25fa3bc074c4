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
