"""GPU tests of the resident single-call kernel (b200mp_session_*): same decisions as the batch entry points and as the oracle, on the
candidate-grid, sorted and unsorted scene variants; survives its own idle exit; the RRT-Connect demo (BASELINE config 1) run through it."""
import time

import numpy as np
import pytest

import ompl_b200
from ompl_b200 import synth
from oracle import port
from tests.util import assert_validity_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot,scene_kw,sort,grid", [
    ("panda", dict(seed=31, n_spheres=120, n_cuboids=120, n_capsules=20, n_zcuboids=20, n_zcapsules=20), True, True),
    ("panda", dict(seed=32, n_spheres=40, n_cuboids=40, n_capsules=8), True, False),
    ("panda", dict(seed=33, n_spheres=40, n_cuboids=40, n_capsules=8), False, None),
    ("fetch", dict(seed=34, n_spheres=60, n_cuboids=60, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6), True, True),
    ("baxter", dict(seed=5, n_spheres=6, n_cuboids=6, n_capsules=4), True, None),
])
def test_session_equals_batch_and_oracle(ctx, robot, scene_kw, sort, grid):
    scene = synth.clutter_scene(**scene_kw)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene, sort=sort, grid=grid)
    pe = synth.fill_env(port.PortEnvironment(), scene, sort=sort)
    q = synth.uniform_states(robot, 600, 17)
    sv = ctx.validate_states(robot, env, q)
    starts = q[sv == 1][:300]
    q0, q1 = synth.local_edges(robot, starts, 19, 0.02, 3.0)   # 1 .. ~12 rakes: one pass and several passes of the four warps
    mv = ctx.validate_motions(robot, env, q0, q1)
    s = ompl_b200.Session(ctx, robot, env)
    got_s = np.array([s.is_valid(x) for x in q], np.uint8)
    got_m = np.array([s.check_motion(a, b) for a, b in zip(q0, q1)], np.uint8)
    np.testing.assert_array_equal(got_s, sv)
    np.testing.assert_array_equal(got_m, mv)
    assert_validity_parity(got_s, pe.validate_states(robot, q, nthreads=4), lambda b: pe.edge_min_clearance(robot, q[b], q[b], resolution=1), "states")
    assert_validity_parity(got_m, pe.validate_motions(robot, q0, q1, nthreads=4), lambda b: pe.edge_min_clearance(robot, q0[b], q1[b]), "motions")
    # doubles (what OMPL stores) go through the same cast as the bridge's ompl_to_vamp
    assert s.check_motion(q0[0].astype(np.float64), q1[0].astype(np.float64)) == bool(mv[0])
    st = s.stats()
    assert st["calls"] == len(q) + len(q0) + 1 and st["launches"] >= 1
    s.close()


def test_session_restarts_after_idle_exit(ctx, monkeypatch):
    monkeypatch.setenv("B200MP_SESSION_IDLE_MS", "5")
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    s = ompl_b200.Session(ctx, "panda", env)
    assert s.is_valid(np.asarray(synth.CAGE_START, np.float32)) and s.is_valid(np.asarray(synth.CAGE_GOAL, np.float32))
    first = s.stats()["launches"]
    for _ in range(3):
        time.sleep(0.05)  # ten idle periods: the kernel has left
        assert not s.check_motion(np.asarray(synth.CAGE_START, np.float32), np.asarray(synth.CAGE_GOAL, np.float32))
    assert s.stats()["launches"] >= first + 1
    # the context stays usable beside a live session (other kernels run while the resident block polls)
    q = synth.uniform_states("panda", 4096, 3)
    assert ctx.validate_states("panda", env, q).shape == (4096,)
    s.close()


def test_two_sessions_and_environment_changes(ctx):
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    a = ompl_b200.Session(ctx, "panda", env)
    a2 = ompl_b200.Session(ctx, "panda", env)   # several sessions on one environment (one per planner thread)
    start = np.asarray(synth.CAGE_START, np.float32)
    assert a.is_valid(start) and a2.is_valid(start)
    tf, ball = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.05]], np.float32), [[0, 0, 0.1, 0.3]]
    # a resident kernel reads the environment's arrays: changing them under it is refused, not undefined
    for change in (lambda: env.set_attachment(tf, ball), env.commit):
        with pytest.raises(ompl_b200.B200MPError) as e:
            change()
        assert e.value.code == 5  # B200MP_ERR_STATE
    a.close()
    a2.close()
    env.set_attachment(tf, ball)
    b = ompl_b200.Session(ctx, "panda", env)   # opened after the grasp: sees the attachment
    assert not b.is_valid(start)
    b.close()
    env.clear_attachment()
    c = ompl_b200.Session(ctx, "panda", env)
    assert c.is_valid(start)
    c.close()
