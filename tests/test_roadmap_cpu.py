"""CPU tests of the roadmap host logic: the sample stream's generator against the published Philox known answers, the
edge rule, and the N > 1 path of RoadmapBuilder under gloo with world_size 2 (sharding, ragged all-gathers, assembly)
using the checker-backed phases. The same builder code runs on NCCL with the CUDA engine in the gpu tests."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ompl_b200 import synth
from ompl_b200.roadmap import RoadmapBuilder
from oracle import port
from tests.checker_backend import CheckerBackend


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = port.philox4x32_10(np.array([ctr], np.uint32), np.array(key, np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_candidates_in_bounds_and_stream_is_position_independent():
    lo, hi = port.bounds("panda")
    a = port.prm_candidates("panda", 7, 0, 5000)
    assert a.dtype == np.float32 and (a >= lo).all() and (a <= hi).all()
    np.testing.assert_array_equal(port.prm_candidates("panda", 7, 1234, 100), a[1234:1334])
    assert not np.array_equal(port.prm_candidates("panda", 8, 0, 100), a[:100])
    # roughly uniform
    assert np.abs((a - lo) / (hi - lo) - 0.5).mean(axis=0).max() < 0.27


def test_edge_rule():
    nbr = np.array([[1, 2], [0, 2], [3, 1], [2, 0]], np.int64)   # 0<->1 mutual, 0->2 one-sided, 1<->2 mutual, 2<->3 mutual, 3->0 one-sided
    d = np.arange(8, dtype=np.float64).reshape(4, 2)
    e, w = port.prm_edge_list(nbr, d)
    # mutual pairs: {0,1} odd -> kept by 1; {1,2} odd -> kept by 2; {2,3} odd -> kept by 3; one-sided: 0->2 by 0, 3->0 by 3
    assert e.tolist() == [[0, 2], [0, 1], [1, 2], [2, 3], [0, 3]]
    assert w.tolist() == [1.0, 2.0, 5.0, 6.0, 7.0]
    nbr2 = np.array([[2, 1], [0, 2], [0, 1]], np.int64)            # {0,2} even -> kept by 0; {0,1} odd -> by 1; {1,2} odd -> by 2
    e2, _ = port.prm_edge_list(nbr2, np.zeros((3, 2)))
    assert e2.tolist() == [[0, 2], [0, 1], [1, 2]]


def _build(rank, world, port_file, n, k, seed, q):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_file))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    env = synth.fill_env(port.PortEnvironment(), synth.cage_scene())
    rm = RoadmapBuilder(CheckerBackend(env, "panda", nthreads=2)).build(n, k, seed, candidates_per_round=700)
    q.put((rank, rm.states.numpy(), rm.edges.numpy(), rm.valid.numpy(), rm.weights.numpy(), rm.stats))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, n, k, seed):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_build, args=(r, world, 29500 + (os.getpid() % 2000), n, k, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_builder_world2_gloo_matches_world1():
    n, k, seed = 301, 6, 5          # odd n: ragged query blocks
    one = _run(1, n, k, seed)[0]
    two = _run(2, n, k, seed)
    for r in two:                    # every rank holds the whole roadmap, identical to the single-process one
        for a, b in zip(one[1:5], r[1:5]):
            np.testing.assert_array_equal(a, b)
    assert two[0][5]["world"] == 2 and two[0][5]["local_edges"] + two[1][5]["local_edges"] == one[5]["n_edges"]
    states, edges, valid, weights = one[1:5]
    # properties of the result itself
    env = synth.fill_env(port.PortEnvironment(), synth.cage_scene())
    assert (env.validate_states("panda", states) == 1).all()
    cand = port.prm_candidates("panda", seed, 0, one[5]["n_candidates"])
    ok = env.validate_states("panda", cand) == 1
    np.testing.assert_array_equal(states, cand[ok][:n])          # the first n valid candidates, in stream order
    assert (edges[:, 0] < edges[:, 1]).all()
    assert len({tuple(e) for e in edges.tolist()}) == len(edges)  # undirected de-duplication
    d = np.sqrt(((states[edges[:, 0]].astype(np.float64) - states[edges[:, 1]].astype(np.float64)) ** 2).sum(axis=1))
    np.testing.assert_allclose(weights, d, rtol=1e-12)
    assert 0 < valid.sum() < len(valid)
