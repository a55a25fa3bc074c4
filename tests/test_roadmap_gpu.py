"""GPU parity tests of the batched PRM roadmap builder (ompl_b200/csrc/roadmap.cu, ompl_b200/roadmap.py) against the
CPU checker: same sample stream, identical kNN index sets, the same edge list, edge validity bit-exact outside the
1e-5 m clearance band."""
import os

import numpy as np
import pytest
import torch

import ompl_b200
from ompl_b200 import synth
from ompl_b200.roadmap import EngineBackend, RoadmapBuilder
from oracle import port
from tests.checker_backend import CheckerBackend
from tests.util import assert_validity_parity

pytestmark = pytest.mark.gpu


def _scene(name):
    return synth.cage_scene() if name == "cage" else synth.clutter_scene(1234, 500, 500)


@pytest.mark.parametrize("scene,n,k,seed", [("cage", 3000, 8, 1), ("c2", 1500, 16, 42), ("cage", 5, 16, 3)])
def test_prm_build_matches_checker(ctx, scene, n, k, seed):
    sc = _scene(scene)
    env = synth.fill_env(ompl_b200.Environment(ctx), sc)
    pe = synth.fill_env(port.PortEnvironment(), sc)
    got = ctx.prm_build("panda", env, n, k, seed)
    want = RoadmapBuilder(CheckerBackend(pe, "panda", nthreads=8)).build(n, k, seed)
    np.testing.assert_array_equal(got["states"], want.states.numpy())      # same Philox stream, same validity decisions
    np.testing.assert_array_equal(got["edges"], want.edges.numpy())         # identical kNN sets -> identical edge list
    np.testing.assert_array_equal(got["weights"], want.weights.numpy())     # fp64 distances bit-identical
    s, e = got["states"], got["edges"]
    assert_validity_parity(got["valid"], want.valid.numpy(), lambda b: pe.edge_min_clearance("panda", s[e[b, 0]], s[e[b, 1]]), "roadmap edges")
    st = got["stats"]
    assert st["n_states"] == n and st["n_edges"] == len(e) and st["n_valid_edges"] == int(got["valid"].sum())
    assert st["n_candidates"] >= n


@pytest.mark.parametrize("robot,scene_kw,n,k", [
    ("ur5", dict(seed=51, n_spheres=100, n_cuboids=100), 1200, 8),        # 6 joints
    ("fetch", dict(seed=52, n_spheres=40, n_cuboids=40, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6), 1000, 8),   # 8 joints: both Philox blocks full
    ("baxter", dict(seed=5, n_spheres=6, n_cuboids=6, n_capsules=4), 600, 6),   # 14 joints: four Philox blocks per candidate
])
def test_prm_build_other_robots(ctx, robot, scene_kw, n, k):
    sc = synth.clutter_scene(**scene_kw)
    env = synth.fill_env(ompl_b200.Environment(ctx), sc)
    pe = synth.fill_env(port.PortEnvironment(), sc)
    got = ctx.prm_build(robot, env, n, k, 77)
    want = RoadmapBuilder(CheckerBackend(pe, robot, nthreads=8)).build(n, k, 77)
    np.testing.assert_array_equal(got["states"], want.states.numpy())
    np.testing.assert_array_equal(got["edges"], want.edges.numpy())
    np.testing.assert_array_equal(got["weights"], want.weights.numpy())
    s, e = got["states"], got["edges"]
    assert_validity_parity(got["valid"], want.valid.numpy(), lambda b: pe.edge_min_clearance(robot, s[e[b, 0]], s[e[b, 1]]), "roadmap edges")


def test_builder_with_engine_backend_equals_prm_build(ctx):
    sc = _scene("cage")
    env = synth.fill_env(ompl_b200.Environment(ctx), sc)
    a = ctx.prm_build("panda", env, 20000, 16, 9)
    rm = RoadmapBuilder(EngineBackend(ctx, env, "panda")).build(20000, 16, 9, candidates_per_round=30000)
    np.testing.assert_array_equal(a["states"], rm.states.cpu().numpy())
    np.testing.assert_array_equal(a["edges"], rm.edges.cpu().numpy())
    np.testing.assert_array_equal(a["valid"], rm.valid.cpu().numpy())
    np.testing.assert_array_equal(a["weights"], rm.weights.cpu().numpy())
    pd = rm.to_planner_data()
    assert pd["vertices"].shape == (20000, 7) and len(pd["edges"]) == int(a["valid"].sum())


def test_roadmap_properties_at_scale(ctx):
    """Size-independent properties at 2^17 states (the full 2^20 build runs in bench.py): every state valid, every vertex
    has <= k kept + incoming edges consistent with the rule, no duplicate edges, weights = end-point distance."""
    env = synth.fill_env(ompl_b200.Environment(ctx), _scene("c2"))
    n, k = 1 << 17, 16
    r = ctx.prm_build("panda", env, n, k, 42)
    s, e, w = r["states"], r["edges"], r["weights"]
    assert (ctx.validate_states("panda", env, s) == 1).all()
    assert (e[:, 0] < e[:, 1]).all() and e.max() < n
    key = e[:, 0].astype(np.int64) * n + e[:, 1]
    assert len(np.unique(key)) == len(key)
    assert n * k / 2 <= len(e) <= n * k
    d = np.sqrt(((s[e[:, 0]].astype(np.float64) - s[e[:, 1]].astype(np.float64)) ** 2).sum(axis=1))
    np.testing.assert_allclose(w, d, rtol=1e-12)
    # edge validity is what the motion entry point says for the same end points
    sub = np.random.default_rng(0).choice(len(e), 50000, replace=False)
    np.testing.assert_array_equal(r["valid"][sub], ctx.validate_motions("panda", env, s[e[sub, 0]], s[e[sub, 1]]))


def _nccl_worker(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    c = ompl_b200.Context(rank)
    env = synth.fill_env(ompl_b200.Environment(c), synth.cage_scene())
    rm = RoadmapBuilder(EngineBackend(c, env, "panda")).build(70001, 16, 9)
    q.put((rank, rm.states.cpu().numpy(), rm.edges.cpu().numpy(), rm.valid.cpu().numpy(), rm.weights.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_roadmap_equals_single_gpu(ctx):
    import torch.multiprocessing as mp
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    one = ctx.prm_build("panda", env, 70001, 16, 9)
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for o in outs:
        np.testing.assert_array_equal(o[1], one["states"])
        np.testing.assert_array_equal(o[2], one["edges"])
        np.testing.assert_array_equal(o[3], one["valid"])
        np.testing.assert_array_equal(o[4], one["weights"])


# ---- the C library's own collective build (b200mp_prm_build_multi): NCCL inside the library ------------------------------------
def _cabi_worker(rank, world, uid_q, out_q, n, k, seed):
    c = ompl_b200.Context(rank)
    env = synth.fill_env(ompl_b200.Environment(c), synth.cage_scene())
    uid = uid_q.get(timeout=120)
    comm = ompl_b200.Comm(c, uid, rank, world)
    r = comm.prm_build("panda", env, n, k, seed, want_neighbors=True)
    out_q.put((rank, r["states"], r["edges"], r["valid"], r["weights"], r["neighbors"], r["stats"]))
    comm.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_c_abi_roadmap_equals_single_gpu(ctx):
    """One process per GPU, the NCCL unique id handed over a queue: the roadmap equals b200mp_prm_build's on every rank."""
    import torch.multiprocessing as mp
    n, k, seed = 70001, 16, 9
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    one = ctx.prm_build("panda", env, n, k, seed)
    mpc = mp.get_context("spawn")
    uid_q, out_q = mpc.Queue(), mpc.Queue()
    uid = ompl_b200.comm_unique_id()
    procs = [mpc.Process(target=_cabi_worker, args=(r, 2, uid_q, out_q, n, k, seed)) for r in range(2)]
    for p in procs:
        p.start()
        uid_q.put(uid)
    outs = [out_q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for o in outs:
        np.testing.assert_array_equal(o[1], one["states"])
        np.testing.assert_array_equal(o[2], one["edges"])
        np.testing.assert_array_equal(o[3], one["valid"])
        np.testing.assert_array_equal(o[4], one["weights"])
        assert o[6]["n_local_edges"] < o[6]["n_edges"]
    np.testing.assert_array_equal(outs[0][5], outs[1][5])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_single_process_threads(ctx):
    """One process, one thread per GPU (b200mp_comm_create_all = ncclCommInitAll, environment cloned across contexts)."""
    import ctypes as C
    import threading
    from ompl_b200 import engine
    L = engine.load_library()
    n, k, seed = 30011, 12, 5
    env0 = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    one = ctx.prm_build("panda", env0, n, k, seed)
    c1 = ompl_b200.Context(1)
    h_env1 = C.c_void_p()
    engine._check(L.b200mp_env_clone(env0._h, c1._h, C.byref(h_env1)))
    ctxs = (C.c_void_p * 2)(ctx._h, c1._h)
    comms = (C.c_void_p * 2)()
    engine._check(L.b200mp_comm_create_all(ctxs, 2, comms))
    envs = [env0._h, h_env1]
    res = [None, None]

    def run(r):
        cap = n * k
        st = engine.RoadmapStats()
        states = np.empty((n, 7), np.float32)
        edges = np.empty((cap, 2), np.int32)
        valid = np.empty(cap, np.uint8)
        w = np.empty(cap, np.float64)
        rc = L.b200mp_prm_build_multi(comms[r], 0, envs[r], seed, n, k, 0, 0, states.ctypes.data, edges.ctypes.data, valid.ctypes.data, w.ctypes.data,
                                      None, cap, C.byref(st))
        res[r] = (rc, states, edges[:st.n_edges], valid[:st.n_edges], w[:st.n_edges])

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    for r in range(2):
        assert res[r] is not None and res[r][0] == 0
        np.testing.assert_array_equal(res[r][1], one["states"])
        np.testing.assert_array_equal(res[r][2], one["edges"])
        np.testing.assert_array_equal(res[r][3], one["valid"])
        np.testing.assert_array_equal(res[r][4], one["weights"])
    for r in range(2):
        L.b200mp_comm_destroy(comms[r])
    L.b200mp_env_destroy(h_env1)
    c1.close()


def test_comm_of_one_needs_no_nccl(ctx):
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    comm = ompl_b200.Comm(ctx, None, 0, 1)
    a = comm.prm_build("panda", env, 5000, 8, 3, want_neighbors=True)
    b = ctx.prm_build("panda", env, 5000, 8, 3)
    for key in ("states", "edges", "valid", "weights"):
        np.testing.assert_array_equal(a[key], b[key])
    # the neighbour table behind the edges: ids ascending by distance, self excluded, equal to the container's own answer
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.add(a["states"].astype(np.float64))
    idx, _, _ = nn.nearest_k(a["states"][:256].astype(np.float64), 8, self_ids=np.arange(256))
    np.testing.assert_array_equal(a["neighbors"][:256], idx.astype(np.int32))
    comm.close()


def test_incremental_mode_is_stock_prm_graph(ctx):
    """B200MP_PRM_INCREMENTAL: state i is connected to its k nearest among 0 .. i - 1 (PRM::addMilestone with KStrategy queried before
    nn_->add, PRM.cpp:583-601): edge set equal to a brute-force restatement of that loop, validity equal to the motion entry point."""
    env = synth.fill_env(ompl_b200.Environment(ctx), synth.cage_scene())
    n, k = 3000, 10
    comm = ompl_b200.Comm(ctx, None, 0, 1)
    r = comm.prm_build("panda", env, n, k, 7, want_neighbors=True, incremental=True)
    base = ctx.prm_build("panda", env, n, k, 7)
    np.testing.assert_array_equal(r["states"], base["states"])          # same milestones, only the connection rule differs
    s = r["states"].astype(np.float64)
    want_edges = []
    for i in range(1, n):
        d2 = np.zeros(i)
        for j in range(7):
            diff = s[:i, j] - s[i, j]
            d2 += diff * diff
        order = np.lexsort((np.arange(i), d2))[:k]
        np.testing.assert_array_equal(r["neighbors"][i, :len(order)], order)
        assert (r["neighbors"][i, len(order):] == -1).all()
        want_edges += [(int(j), i) for j in order]
    assert (r["neighbors"][0] == -1).all()
    np.testing.assert_array_equal(r["edges"], np.array(want_edges, np.int32))
    e = r["edges"]
    np.testing.assert_array_equal(r["valid"], ctx.validate_motions("panda", env, r["states"][e[:, 0]], r["states"][e[:, 1]]))
    comm.close()
