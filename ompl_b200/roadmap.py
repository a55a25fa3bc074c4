"""Batched PRM roadmap builder, sharded over the GPUs of one node (one process per GPU).

Replaces the per-milestone loop of ompl::geometric::PRM (src/ompl/geometric/planners/prm/src/PRM.cpp:240-270,570-604)
by whole-batch phases (ompl_b200/csrc/roadmap.cu). Rank g of G owns the query block [g*B, (g+1)*B), B = ceil(n / G),
and the edges its queries keep; the exchange steps are the ones SURVEY.md 8(e) lists:

  1. sample     every rank validates a contiguous slice of the candidate stream; all-gather of the per-rank counts and
                of the valid states (padded to the largest count) -> every rank holds all n states, in stream order.
                The states are the first n valid candidates of the stream, independent of G.
  2. kNN        each rank answers a share of ALL queries (whole cells of the kNN index, so its blocks stay full) against the
                full set, self excluded, into [n][k] tables pre-filled with -1 / +inf
  3. all-reduce (max on ids, min on distances) merges the tables: every rank holds the neighbour lists of all states
  4. edges      each rank keeps, for the states i of its block, (i, j) when j does not list i or i is the pair's keeper
                (lower index if i + j even, higher if odd: balanced counts); validates them
  5. all-gather of counts, edges, weights and validity -> every rank holds the whole roadmap, same order for any G

All compute goes through the C ABI of the CUDA engine (`EngineBackend`); torch is used for device buffers and the NCCL
collectives only. Any object with the four phase methods can stand in for the backend (the CPU gloo tests use a
checker-backed one); there is no CPU fallback in this package.
"""
from __future__ import annotations

import contextlib
import time
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from .engine import Comm, Context, Environment, NearestNeighbors, comm_unique_id, robot_info


@dataclass
class Roadmap:
    states: torch.Tensor    # [n][dim] float32
    edges: torch.Tensor     # [E][2] int32, (older milestone, newer milestone)
    valid: torch.Tensor     # [E] uint8
    weights: torch.Tensor   # [E] float64 = edge length (PRM's motionCost for path-length objectives)
    stats: dict = field(default_factory=dict)
    neighbors: torch.Tensor | None = None       # [n][k] int64: the kNN table the edges were derived from (-1 = none)
    neighbor_dist: torch.Tensor | None = None   # [n][k] float64, ascending

    def to_planner_data(self):
        """Vertices and valid edges in the form ompl::base::PlannerData takes (addVertex / addEdge with weight), ready for
        ompl::geometric::PRM(const PlannerData &) (PRM.cpp:96-156)."""
        keep = self.valid.bool()
        return dict(vertices=self.states.double().cpu().numpy(), edges=self.edges[keep].cpu().numpy(),
                    weights=self.weights[keep].cpu().numpy())


def make_comm(ctx: Context, group=None) -> Comm:
    """This process's rank of the GPU group behind b200mp_prm_build_multi: rank 0 makes the NCCL unique id, torch.distributed (any
    backend) carries its 128 bytes to the other processes, every process joins with b200mp_comm_create. Without an initialised
    process group the group has one member and no NCCL is involved."""
    rank, world = _world(group)
    if world == 1:
        return Comm(ctx, None, 0, 1)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return Comm(ctx, box[0], rank, world)


def build_roadmap(comm: Comm, env: Environment, robot: str, n: int, k: int, seed: int, resolution: int = 0, want_arrays: bool = True,
                  want_neighbors: bool = False) -> "Roadmap":
    """The whole build inside the C library (b200mp_prm_build_multi): collective over the ranks of `comm`, NCCL exchanges included.
    Returns host tensors (every rank holds the whole roadmap) or, with want_arrays = False, only the statistics."""
    r = comm.prm_build(robot, env, n, k, seed, resolution, want_arrays=want_arrays, want_neighbors=want_neighbors)
    st = dict(r["stats"], world=comm.world)
    if not want_arrays:
        return Roadmap(None, None, None, None, st)
    nb = torch.from_numpy(r["neighbors"]) if r.get("neighbors") is not None else None
    return Roadmap(torch.from_numpy(r["states"]), torch.from_numpy(r["edges"]), torch.from_numpy(r["valid"]), torch.from_numpy(r["weights"]), st, nb)


class EngineBackend:
    """The four roadmap phases on one GPU through the C ABI (the torch.distributed front-end of RoadmapBuilder; the C library's own
    collective build is build_roadmap above)."""

    def __init__(self, ctx: Context, env: Environment, robot: str):
        self.ctx, self.env, self.robot = ctx, env, robot
        self.dim = robot_info(robot)["dim"]
        self.device = torch.device("cuda", ctx.device)
        info = robot_info(robot)
        self.max_norm = float(np.linalg.norm(np.maximum(np.abs(info["lo"]), np.abs(info["hi"])).astype(np.float64)))
        self._knn = None
        # One explicit (non-default) stream carries the engine kernels, the torch copies and the NCCL collectives of a
        # build, so they are ordered with respect to each other. (Handing the C ABI the default stream's handle, 0, would
        # mean "the context's own stream" there and break that ordering.)
        self.stream = torch.cuda.Stream(device=self.device)

    def stream_context(self):
        return torch.cuda.stream(self.stream)

    def _stream(self) -> int:
        return self.stream.cuda_stream

    def sample(self, seed: int, first: int, count: int, out: torch.Tensor) -> int:
        return self.ctx.prm_sample_ptr(self.robot, self.env, seed, first, count, out.data_ptr(), out.shape[0], self._stream())

    def neighbors_part(self, states: torch.Tensor, k: int, part: int, parts: int, idx: torch.Tensor, dst: torch.Tensor):
        """Fill the rows of this rank's share of ALL queries (every state against the whole set, self excluded) into the
        [n][k] tables idx / dst; rows of other shares stay untouched (b200mp_knn_query_part_device_f32)."""
        n = states.shape[0]
        if self._knn is None:
            self._knn = NearestNeighbors(self.ctx, self.dim)   # kept across builds: its device buffers are grow-only
        else:
            self._knn.clear()
        self._knn.add_device_f32(states.data_ptr(), n, self.max_norm, self._stream())
        self._knn.query_part_device_f32(states.data_ptr(), n, k, idx.data_ptr(), dst.data_ptr(), part, parts, self_base=0, exclude_self=True,
                                        stream=self._stream())

    def edges(self, nbr: torch.Tensor, dist_local: torch.Tensor, q_first: int, q_count: int):
        n, k = nbr.shape
        cap = max(q_count * k, 1)
        e = torch.empty((cap, 2), dtype=torch.int32, device=self.device)
        w = torch.empty(cap, dtype=torch.float64, device=self.device)
        ne = self.ctx.prm_edges_ptr(nbr.data_ptr(), dist_local.data_ptr(), n, k, q_first, q_count, e.data_ptr(), w.data_ptr(), cap,
                                    self._stream())
        return e[:ne], w[:ne]

    def validate_edges(self, states: torch.Tensor, edges: torch.Tensor, resolution: int) -> torch.Tensor:
        v = torch.empty(edges.shape[0], dtype=torch.uint8, device=self.device)
        if edges.shape[0]:
            self.ctx.validate_edges_ptr(self.robot, self.env, states.data_ptr(), edges.data_ptr(), edges.shape[0], v.data_ptr(),
                                        resolution, self._stream())
        return v


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _all_gather_ragged(local: torch.Tensor, group, world: int):
    """Concatenate per-rank tensors of different leading sizes, in rank order (counts all-gather + padded all-gather)."""
    if world == 1:
        return local
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    cnts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    cnts = cnts.tolist()
    m = max(max(cnts), 1)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(cnts)], dim=0)


class RoadmapBuilder:
    def __init__(self, backend, group=None):
        self.backend = backend
        self.group = group
        self.rank, self.world = _world(group)

    def _sync(self):
        if self.backend.device.type == "cuda":
            torch.cuda.synchronize(self.backend.device)

    def build(self, n: int, k: int, seed: int, resolution: int = 0, candidates_per_round: int | None = None) -> Roadmap:
        ctx = self.backend.stream_context() if hasattr(self.backend, "stream_context") else contextlib.nullcontext()
        with ctx:
            rm = self._build(n, k, seed, resolution, candidates_per_round)
        self._sync()
        return rm

    def _build(self, n, k, seed, resolution, candidates_per_round) -> Roadmap:
        be, G, g = self.backend, self.world, self.rank
        dev, dim = be.device, be.dim
        t = [time.perf_counter()]

        def mark():
            self._sync()
            t.append(time.perf_counter())

        # ---- 1. sample: rounds over the candidate stream, contiguous slice per rank, until n valid states exist
        states = torch.empty((n, dim), dtype=torch.float32, device=dev)
        have, drawn, rounds = 0, 0, 0
        per_round = candidates_per_round or max(n, 1 << 16)
        while have < n:
            rounds += 1
            if rounds > 4096:
                raise RuntimeError(f"no valid states after {drawn} candidates")
            per_rank = -(-per_round // G)
            buf = torch.empty((per_rank, dim), dtype=torch.float32, device=dev)
            got = be.sample(seed, drawn + g * per_rank, per_rank, buf)
            block = _all_gather_ragged(buf[:got], self.group, G)
            take = min(block.shape[0], n - have)
            states[have:have + take] = block[:take]
            have += take
            drawn += per_rank * G
            if have < n:
                rate = max(1e-4, have / drawn)
                per_round = max(1 << 16, int(1.05 * (n - have) / rate))
        mark()
        # ---- 2. kNN: every rank answers its share of the queries (whole index cells), tables pre-filled with -1 / +inf
        B = -(-n // G)
        q_first = min(g * B, n)
        q_count = min(B, n - q_first)
        kk = min(k, n - 1)
        if kk < 1:
            empty = torch.empty((0, 2), dtype=torch.int32, device=dev)
            return Roadmap(states, empty, torch.empty(0, dtype=torch.uint8, device=dev), torch.empty(0, dtype=torch.float64, device=dev),
                           dict(n_states=n, n_candidates=drawn, n_edges=0, n_valid_edges=0, world=G))
        nbr = torch.full((n, kk), -1, dtype=torch.int64, device=dev)
        dst = torch.full((n, kk), float("inf"), dtype=torch.float64, device=dev)
        be.neighbors_part(states, kk, g, G, nbr, dst)
        mark()
        # ---- 3. neighbour table of every state: element-wise max / min over the ranks' partially filled tables
        if G > 1:
            dist.all_reduce(nbr, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(dst, op=dist.ReduceOp.MIN, group=self.group)
        # ---- 4. edge list of the local index block and its validity
        e_loc, w_loc = be.edges(nbr, dst[q_first:q_first + q_count], q_first, q_count)
        mark()
        v_loc = be.validate_edges(states, e_loc, resolution)
        mark()
        # ---- 5. whole roadmap on every rank
        edges = _all_gather_ragged(e_loc, self.group, G)
        weights = _all_gather_ragged(w_loc, self.group, G)
        valid = _all_gather_ragged(v_loc, self.group, G)
        mark()
        ms = [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]
        stats = dict(n_states=n, n_candidates=drawn, n_edges=int(edges.shape[0]), n_valid_edges=int(valid.sum().item()), world=G,
                     sample_rounds=rounds, ms_sample=ms[0], ms_knn=ms[1], ms_edges=ms[2], ms_validate=ms[3], ms_gather=ms[4],
                     ms_total=sum(ms), local_edges=int(e_loc.shape[0]))
        return Roadmap(states, edges, valid, weights, stats, nbr, dst)
