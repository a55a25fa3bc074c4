"""GPU parity tests: the CUDA engine, called through the C ABI, against the golden vectors of the unmodified
reference and against the oracle on seeded inputs. Bit-exact except decisions whose min clearance is < 1e-5 m."""
import numpy as np
import pytest

import ompl_b200
from ompl_b200 import synth
from oracle import port
from tests.conftest import have_reference_build
from tests.util import assert_validity_parity, load_golden

pytestmark = pytest.mark.gpu

CASES = ["panda_cage", "panda_clutter", "panda_clutter_unsorted", "fetch_clutter", "ur5_clutter", "baxter_clutter", "panda_terrain", "fetch_terrain_clutter"]


def clr_states(pe, robot, q):
    return lambda b: pe.edge_min_clearance(robot, q[b], q[b], resolution=1)


def clr_motions(pe, robot, q0, q1, res=0):
    return lambda b: pe.edge_min_clearance(robot, q0[b], q1[b], resolution=res)


@pytest.mark.parametrize("name", CASES)
def test_golden_validity(ctx, name):
    z, scene, robot, is_sorted = load_golden(name)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene, sort=is_sorted)
    pe = synth.fill_env(port.PortEnvironment(), scene, sort=is_sorted)
    for kind in range(5):
        np.testing.assert_allclose(env.min_distances(kind), z[f"min_d_{kind}"], rtol=0, atol=2.5e-7)
    sv = ctx.validate_states(robot, env, z["states"])
    assert_validity_parity(sv, z["states_valid"], clr_states(pe, robot, z["states"]), name + " states")
    mv = ctx.validate_motions(robot, env, z["q0"], z["q1"])
    assert_validity_parity(mv, z["motions_valid"], clr_motions(pe, robot, z["q0"], z["q1"]), name + " motions")


@pytest.mark.parametrize("tag", ["panda", "panda_unsorted", "fetch", "ur5", "baxter"])
@pytest.mark.parametrize("grid", [True, False])
def test_golden_attachments(ctx, tag, grid):
    """Robot::fkcc_attach: golden answers of the unmodified reference (tests/golden/attach.npz), set / replace / clear after commit."""
    import os
    from tests.golden.make_golden import ATTACH_SCENES
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "attach.npz"))
    robot, is_sorted = tag.split("_")[0], not tag.endswith("unsorted")
    if not is_sorted and grid:
        pytest.skip("unsorted scenes have no grid")
    scene = synth.clutter_scene(**ATTACH_SCENES[tag])
    env = ompl_b200.Environment(ctx)
    env.set_grid(grid)
    synth.fill_env(env, scene, sort=is_sorted)
    pe = synth.fill_env(port.PortEnvironment(), scene, sort=is_sorted)
    q, q0, q1 = z[tag + "_states"], z[tag + "_q0"], z[tag + "_q1"]
    assert_validity_parity(ctx.validate_states(robot, env, q), z[tag + "_states_valid_bare"], clr_states(pe, robot, q), tag + " bare")
    env.set_attachment(z[tag + "_tf"], z[tag + "_spheres"])
    pe.set_attachment(z[tag + "_tf"], z[tag + "_spheres"])
    assert_validity_parity(ctx.validate_states(robot, env, q), z[tag + "_states_valid"], clr_states(pe, robot, q), tag + " states")
    assert_validity_parity(ctx.validate_motions(robot, env, q0, q1), z[tag + "_motions_valid"], clr_motions(pe, robot, q0, q1), tag + " motions")
    # the bucketed path (packed rakes, first look) sees the attachment as well: a batch large enough to take it
    reps = 4096 // len(q0) + 1
    big0, big1 = np.tile(q0, (reps, 1)), np.tile(q1, (reps, 1))
    got = ctx.validate_motions(robot, env, big0, big1)
    assert_validity_parity(got, np.tile(z[tag + "_motions_valid"], reps), clr_motions(pe, robot, big0, big1), tag + " bucketed motions")
    env.clear_attachment()
    pe.clear_attachment()
    assert_validity_parity(ctx.validate_states(robot, env, q), z[tag + "_states_valid_bare"], clr_states(pe, robot, q), tag + " cleared")


def test_attachment_seeded_vs_oracle_with_cloud_and_terrain(ctx):
    """An attachment in a scene with a point cloud and a height field as well (every obstacle kind the attachment is tested against)."""
    scene = synth.clutter_scene(14, 20, 20, n_capsules=6, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6)
    scene.update(synth.terrain(seed=5))
    pts = synth.shelf_cloud(800, seed=3, robot="fetch")
    env = ompl_b200.Environment(ctx)
    pe = port.PortEnvironment()
    env.add_pointcloud(pts, 0.012, 0.24, 0.0025)
    pe.add_capt(ompl_b200.capt_build(pts, 0.012, 0.24, 0.0025))
    synth.fill_env(env, scene)
    synth.fill_env(pe, scene)
    tf = np.array([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 0.1]], np.float32)
    sph = np.array([[0, 0, 0, 0.06], [0, 0, 0.12, 0.05], [0.05, 0, 0.2, 0.03]], np.float32)
    env.set_attachment(tf, sph)
    pe.set_attachment(tf, sph)
    q = synth.uniform_states("fetch", 8000, 23)
    want = pe.validate_states("fetch", q, nthreads=8)
    assert_validity_parity(ctx.validate_states("fetch", env, q), want, clr_states(pe, "fetch", q), "states")
    q0, q1 = synth.local_edges("fetch", q[want == 1][:1500], 29, 0.05, 2.0)
    want = pe.validate_motions("fetch", q0, q1, nthreads=8)
    assert_validity_parity(ctx.validate_motions("fetch", env, q0, q1), want, clr_motions(pe, "fetch", q0, q1), "motions")


def test_large_state_batches_are_sifted_with_the_same_decisions(ctx):
    """Batches of >= 2^16 states on a candidate-grid scene go through the sifting pass + packed full pass (launch_states): decisions
    equal the oracle's, and equal the unsifted path's on the same states in smaller batches."""
    scene = synth.clutter_scene(seed=31, n_spheres=300, n_cuboids=300, n_capsules=60, n_zcuboids=60, n_zcapsules=60)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states("panda", 100000, 23)
    got = ctx.validate_states("panda", env, q)
    small = np.concatenate([ctx.validate_states("panda", env, q[i:i + 20000]) for i in range(0, len(q), 20000)])
    np.testing.assert_array_equal(got, small)
    assert 0.02 < got.mean() < 0.98
    want = pe.validate_states("panda", q, nthreads=8)
    assert_validity_parity(got, want, clr_states(pe, "panda", q), "sifted states")


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_first_look_and_queue_takes_do_not_change_decisions(ctx, robot):
    """A large mixed batch (short roadmap-like edges, long motions between random end points, the same edges backwards): the long
    bucket gets a first look that goes on (two states per motion, passes that give up at a few undecided lanes) and the motion kernel
    then takes its queue entries several at a time. Decisions equal those of the same motions in small batches (other sample
    verdicts, other takes) and the oracle's."""
    kw = dict(seed=31, n_spheres=300, n_cuboids=300, n_capsules=60, n_zcuboids=60, n_zcapsules=60)
    if robot == "fetch":
        kw = dict(seed=32, n_spheres=80, n_cuboids=80, n_capsules=20, n_zcuboids=20, n_zcapsules=20, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6)
    scene = synth.clutter_scene(**kw)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states(robot, 400000, 57)
    ok = ctx.validate_states(robot, env, q)
    starts = q[ok == 1][:40000]
    a0, a1 = synth.local_edges(robot, starts, 61, 0.02, 3.0)            # short and long, many valid
    b0, b1 = q[:30000], q[30000:60000]                                   # random end points: long, nearly all invalid
    q0, q1 = np.concatenate([a0, b0, a1]), np.concatenate([a1, b1, a0])                # ... and the same edges backwards
    assert len(q0) >= (1 << 16)
    got = ctx.validate_motions(robot, env, q0, q1)
    small = np.concatenate([ctx.validate_motions(robot, env, q0[i:i + 30000], q1[i:i + 30000]) for i in range(0, len(q0), 30000)])
    np.testing.assert_array_equal(got, small)
    assert 0.05 < got.mean() < 0.95
    sub = np.random.default_rng(5).choice(len(q0), 12000, replace=False)
    want = pe.validate_motions(robot, q0[sub], q1[sub], nthreads=8)
    assert_validity_parity(got[sub], want, clr_motions(pe, robot, q0[sub], q1[sub]), "first look + takes")


def test_demo_known_answers(ctx):
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    assert list(ctx.validate_states("panda", env, [synth.CAGE_START, synth.CAGE_GOAL])) == [1, 1]
    assert list(ctx.validate_motions("panda", env, [synth.CAGE_START], [synth.CAGE_GOAL])) == [0]


@pytest.mark.parametrize("robot,scene_kw", [
    ("panda", dict(seed=31, n_spheres=300, n_cuboids=300, n_capsules=60, n_zcuboids=60, n_zcapsules=60)),
    ("fetch", dict(seed=32, n_spheres=80, n_cuboids=80, n_capsules=20, n_zcuboids=20, n_zcapsules=20, rho_lo=0.6, rho_hi=1.4,
                   z_lo=0.2, z_hi=1.6)),
    ("ur5", dict(seed=33, n_spheres=200, n_cuboids=200, n_capsules=30, n_zcuboids=30, n_zcapsules=30)),
    ("baxter", dict(seed=5, n_spheres=6, n_cuboids=6, n_capsules=4)),
])
def test_seeded_parity_vs_oracle(ctx, robot, scene_kw):
    scene = synth.clutter_scene(**scene_kw)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states(robot, 30000, 17)
    want = pe.validate_states(robot, q, nthreads=8)
    got = ctx.validate_states(robot, env, q)
    assert_validity_parity(got, want, clr_states(pe, robot, q), "states")
    starts = q[want == 1][:4000]
    q0, q1 = synth.local_edges(robot, starts, 19, 0.02, 3.0)
    want = pe.validate_motions(robot, q0, q1, nthreads=8)
    got = ctx.validate_motions(robot, env, q0, q1)
    assert 0.05 < want.mean() < 0.95
    assert_validity_parity(got, want, clr_motions(pe, robot, q0, q1), "motions")


@pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref not built")
def test_parity_vs_compiled_reference_c2_sample(ctx):
    """A sample of BASELINE config 2 (500 spheres + 500 cuboids, 64-state motions) against the reference itself."""
    from oracle import refvamp
    scene = synth.clutter_scene(1234, 500, 500)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    re_ = synth.fill_env(refvamp.RefEnvironment(), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q0, q1 = synth.fixed_length_edges("panda", 1 << 16, 1234)
    want = re_.validate_motions("panda", q0, q1, nthreads=8)
    got = ctx.validate_motions("panda", env, q0, q1)
    assert_validity_parity(got, want, clr_motions(pe, "panda", q0, q1), "C2 sample")


def test_ragged_and_empty_batches(ctx):
    scene = synth.clutter_scene(3, 40, 40)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states("panda", 200, 2)
    for m in (0, 1, 7, 31, 32, 33, 65, 200):
        np.testing.assert_array_equal(ctx.validate_states("panda", env, q[:m]), pe.validate_states("panda", q[:m]))
    q0, q1 = synth.local_edges("panda", q, 4, 0.0, 3.5)
    for m in (0, 1, 3, 9, 200):
        np.testing.assert_array_equal(ctx.validate_motions("panda", env, q0[:m], q1[:m]), pe.validate_motions("panda", q0[:m], q1[:m]))


def test_empty_environment_and_zero_length_motion(ctx):
    env = ompl_b200.Environment(ctx)
    env.commit()
    pe = port.PortEnvironment()
    q = synth.uniform_states("panda", 2048, 8)
    sv = ctx.validate_states("panda", env, q)
    np.testing.assert_array_equal(sv, pe.validate_states("panda", q, nthreads=4))  # self-collision only
    assert 0 < sv.sum() < len(sv)
    np.testing.assert_array_equal(ctx.validate_motions("panda", env, q, q), sv)  # validate_motion(q, q): n = 1, 8 equal lanes


def test_long_motions_many_rakes(ctx):
    # corner-to-corner motions: |dq| ~ 14 rad -> n = 56 rakes = 448 states, several passes per warp
    scene = synth.clutter_scene(12, 30, 30)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    lo, hi = synth.BOUNDS["panda"]
    rng = np.random.default_rng(5)
    q0 = np.where(rng.random((300, 7)) < 0.5, lo, hi).astype(np.float32)
    q1 = np.where(rng.random((300, 7)) < 0.5, lo, hi).astype(np.float32)
    want = pe.validate_motions("panda", q0, q1, nthreads=8)
    got = ctx.validate_motions("panda", env, q0, q1)
    assert_validity_parity(got, want, clr_motions(pe, "panda", q0, q1), "long motions")


def test_resolution_argument(ctx):
    scene = synth.clutter_scene(14, 60, 60)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states("panda", 4000, 3)
    q0, q1 = synth.local_edges("panda", q, 5, 0.1, 2.0)
    for res in (1, 8, 32, 64):
        want = pe.validate_motions("panda", q0, q1, resolution=res, nthreads=8)
        got = ctx.validate_motions("panda", env, q0, q1, resolution=res)
        assert_validity_parity(got, want, clr_motions(pe, "panda", q0, q1, res), f"resolution {res}")


def test_double_precision_entry_points(ctx):
    # OMPL states are double; the bridge casts each coordinate to float (src/ompl/vamp/Utils.h:26)
    scene = synth.clutter_scene(15, 60, 60)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    rng = np.random.default_rng(1)
    lo, hi = synth.BOUNDS["panda"]
    q = rng.uniform(lo, hi, (3000, 7))  # float64
    np.testing.assert_array_equal(ctx.validate_states("panda", env, q), ctx.validate_states("panda", env, q.astype(np.float32)))
    q1 = np.clip(q + rng.normal(scale=0.3, size=q.shape), lo, hi)
    np.testing.assert_array_equal(ctx.validate_motions("panda", env, q, q1),
                                  ctx.validate_motions("panda", env, q.astype(np.float32), q1.astype(np.float32)))


def test_double_precision_large_batch_is_chunked(ctx):
    """fp64 batches of >= 2^17 motions go through the chunked pipeline (host cast into a pinned ring overlapping copies and kernels):
    same answers as the fp32 entry point, ragged last chunk included."""
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.clutter_scene(1234, 120, 120))
    m = (1 << 17) + (1 << 16) + 77
    q0, q1 = synth.fixed_length_edges("panda", m, 99, 0.3, 1.2)
    want = ctx.validate_motions("panda", env, q0, q1)
    got = ctx.validate_motions("panda", env, q0.astype(np.float64), q1.astype(np.float64))
    np.testing.assert_array_equal(got, want)
    assert 0 < want.sum() < m


def test_device_pointer_entry_points(ctx):
    import torch
    scene = synth.clutter_scene(16, 60, 60)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    q0, q1 = synth.fixed_length_edges("panda", 5000, 3)
    want = ctx.validate_motions("panda", env, q0, q1)
    d0, d1 = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    out = torch.zeros(len(q0), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    before = ctx.launch_count
    ctx.validate_motions_ptr("panda", env, d0.data_ptr(), d1.data_ptr(), len(q0), out.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    assert ctx.launch_count - before in (1, 5, 6)  # one kernel, or the bucketed sequence (count, plan, scatter, rakes, [first look,] motions)
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_api_errors(ctx):
    env = ompl_b200.Environment(ctx)
    env.add_spheres([[1, 0, 0, 0.1]])
    q = synth.uniform_states("panda", 4, 1)
    with pytest.raises(ompl_b200.B200MPError) as e:   # not committed
        ctx.validate_states("panda", env, q)
    assert e.value.code == 5
    env.commit()
    ctx.validate_states("panda", env, q)


def test_linearity_properties_full_size(ctx):
    """Size-independent properties at BASELINE scale (2^20 motions): (i) a motion is valid only if both of the
    reference's end samples are (goal state is tested, validate.hh:36-39); (ii) batch order does not matter."""
    scene = synth.clutter_scene(1234, 500, 500)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)
    q0, q1 = synth.fixed_length_edges("panda", 1 << 20, 99)
    v = ctx.validate_motions("panda", env, q0, q1)
    goal_ok = ctx.validate_states("panda", env, q1)
    assert not (v & (1 - goal_ok)).any()
    perm = np.random.default_rng(0).permutation(len(q0))
    v2 = ctx.validate_motions("panda", env, q0[perm], q1[perm])
    np.testing.assert_array_equal(v2, v[perm])
    assert 0.001 < v.mean() < 0.2


@pytest.mark.parametrize("robot,scene_kw", [
    ("panda", dict(seed=41, n_spheres=300, n_cuboids=300, n_capsules=60, n_zcuboids=60, n_zcapsules=60)),
    ("panda", dict(seed=1234, n_spheres=500, n_cuboids=500)),
    ("fetch", dict(seed=42, n_spheres=150, n_cuboids=150, n_capsules=30, n_zcuboids=30, n_zcapsules=30, rho_lo=0.5, rho_hi=1.6,
                   z_lo=0.0, z_hi=1.8)),
    ("ur5", dict(seed=43, n_spheres=300, n_cuboids=300)),
    ("baxter", dict(seed=5, n_spheres=15, n_cuboids=15, n_capsules=4)),
])
def test_grid_culling_makes_the_same_decisions_as_the_ordered_scan(ctx, robot, scene_kw):
    """Sorted scenes are validated through candidate grids; the plain ordered scan (grid off) is the reference semantics."""
    scene = synth.clutter_scene(**scene_kw)
    eg = synth.fill_env(ompl_b200.Environment(ctx), scene, grid=True)
    es = synth.fill_env(ompl_b200.Environment(ctx), scene, grid=False)
    pe = synth.fill_env(port.PortEnvironment(), scene)
    q = synth.uniform_states(robot, 100000, 23)
    a, b = ctx.validate_states(robot, eg, q), ctx.validate_states(robot, es, q)
    assert_validity_parity(a, b, clr_states(pe, robot, q), "grid vs scan, states", band=1e-6)
    starts = q[b == 1][:20000]
    q0, q1 = synth.local_edges(robot, starts, 29, 0.02, 3.0)
    a, b = ctx.validate_motions(robot, eg, q0, q1), ctx.validate_motions(robot, es, q0, q1)
    assert 0.05 < b.mean() < 0.95
    assert_validity_parity(a, b, clr_motions(pe, robot, q0, q1), "grid vs scan, motions", band=1e-6)
    # obstacles far outside the grid's core region and huge ones still collide correctly (border cells are unbounded)
    far = dict(spheres=np.concatenate([scene["spheres"], [[0.3, 0.0, 0.6, 0.25], [5.0, 5.0, 5.0, 0.1], [0.0, -0.4, 0.3, 0.02]]]).astype(np.float32),
               cuboids=scene["cuboids"])
    eg2 = synth.fill_env(ompl_b200.Environment(ctx), far, grid=True)
    es2 = synth.fill_env(ompl_b200.Environment(ctx), far, grid=False)
    np.testing.assert_array_equal(ctx.validate_states(robot, eg2, q[:30000]), ctx.validate_states(robot, es2, q[:30000]))
