#!/usr/bin/env python3
"""Headline benchmark: Panda edge checks/s at 64 interpolated states per edge (BASELINE.json config 2).

Workload: scene seed 1234 = 500 spheres + 500 cuboids, sorted; 2^20 random Panda motions whose
length puts the reference at n = 8 rakes = 64 states. One "step" = one pass of the validation kernel over the batch.

  value     edges/s, whole job, inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       edges/s through the C ABI host entry point (b200mp_validate_motions) with pinned HOST buffers:
            H2D of both state arrays + kernel + D2H of the validity bytes inside the timed region
  roofline  FP32-FMA roofline of the validation kernel on the all-valid sub-workload (no early exit hides work):
            achieved = W_ref (flops the reference's sequential semantics execute, counted by the CPU oracle) / kernel time;
            peak = FP32 FMA throughput measured on this GPU by b200mp_fp32_peak (MEASURED_PEAKS.json has no FP32 entry)
  cpu_baseline   the reference's own CPU path (oracle/_ref, unmodified VAMP) or the port, all host threads, bounded sample

  roadmap   second half of BASELINE's metric: wall time to build a 2^20-sample, k = 16 Panda PRM roadmap in the same scene on
            the N GPUs of the run (sample -> kNN -> edge list -> edge validity; NCCL all-gathers between phases), max over ranks

`--impl reference` times the reference CPU implementation instead (rank 0 only).
Multi-GPU: motion validation shards by motions with no data-path collective -> weak scaling, every rank validates its own 2^20
motions; the roadmap build is a fixed-size job split over the ranks (strong scaling, reported inside `roadmap`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "panda_edge_checks_per_sec_at_64_steps"
UNIT = "edges/s"
M_EDGES = 1 << 20
SCENE_SEED = 1234
EDGE_SEED = 1234


def workload_config(n_gpus):
    return {"workload": "C2: 2^20 random Panda motions x 64 interpolated states vs 500 spheres + 500 cuboids (seed 1234, sorted)",
            "motions_per_gpu": M_EDGES, "states_per_motion": 64, "obstacles": 1000, "robot": "panda",
            "cache": "L2 flushed (256 MiB write) before every timed step", "parallelism": f"motions sharded x{n_gpus}, no collective"}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.max_mhz = None
        self.reasons = set()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_inputs(rank):
    from ompl_b200 import synth
    scene = synth.clutter_scene(SCENE_SEED, 500, 500)
    q0, q1 = synth.fixed_length_edges("panda", M_EDGES, EDGE_SEED + 1000 * rank)
    return scene, q0, q1


def checker_env(scene):
    """(env, kind): the compiled unmodified reference when oracle/_ref exists, else the CPU restatement."""
    from ompl_b200 import synth
    from oracle import port
    try:
        from oracle import refvamp
        if refvamp.available():
            return synth.fill_env(refvamp.RefEnvironment(), scene), "reference"
    except Exception:
        pass
    return synth.fill_env(port.PortEnvironment(), scene), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, q0, q1 = make_inputs(0)
    env, kind = checker_env(scene)
    cores = os.cpu_count() or 1
    sample = 1 << 17  # bounded sample of the same motions per step
    a, b = q0[:sample], q1[:sample]
    for _ in range(max(args.warmup, 1)):
        env.validate_motions("panda", a[:8192], b[:8192], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env.validate_motions("panda", a, b, nthreads=cores)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"first {sample} of the 2^20 motions per step, all host threads (dynamic chunks)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=claim_stdout(), flush=True)


def roadmap_parity(states, edges, valid, nbr, scene, cores, knn_queries=4096, max_edges=0):
    """North-star parity of a built roadmap against the compiled reference (oracle/_ref), on the host cores:
    every edge through the unmodified vamp::planning::validate_motion<Panda, 8, 32> (validate.hh:69-77) and a spread sample of
    queries through the unmodified NearestNeighborsLinear::nearestK (NearestNeighborsLinear.h:119-132). Differing edge decisions are
    classified by the restatement's minimum clearance (the reference cannot report one): inside / outside the 1e-5 m band."""
    from ompl_b200 import synth
    from oracle import port, refnn, refvamp
    out = {"checker": "oracle/_ref (unmodified reference), %d host threads" % cores}
    e = edges if not max_edges else edges[:max_edges]
    v = valid[:len(e)]
    env = synth.fill_env(refvamp.RefEnvironment(), scene)
    q0, q1 = np.ascontiguousarray(states[e[:, 0]]), np.ascontiguousarray(states[e[:, 1]])
    t0 = time.perf_counter()
    want = env.validate_motions("panda", q0, q1, nthreads=cores)
    dt = time.perf_counter() - t0
    bad = np.nonzero(want != v)[0]
    in_band = 0
    if bad.size:
        pe = synth.fill_env(port.PortEnvironment(), scene)
        clr = pe.edge_min_clearance("panda", q0[bad], q1[bad], nthreads=cores)
        in_band = int((np.abs(clr) < 1e-5).sum())
    out.update(edges_compared=int(len(e)), differ=int(bad.size), differ_in_1e5_band=in_band, differ_outside=int(bad.size - in_band),
               reference_valid_edges=int(want.sum()), reference_edges_per_s=len(e) / dt)
    del q0, q1
    # kNN sample
    n, k = nbr.shape
    ids = np.unique(np.linspace(0, n - 1, min(knn_queries, n)).astype(np.int64))
    db = states.astype(np.float64)
    t0 = time.perf_counter()
    ri, rd = refnn.nearest_k(db, db[ids], k, nthreads=cores, exclude_ids=ids)
    dtk = time.perf_counter() - t0
    gi = nbr[ids].astype(np.int64)
    gd2 = np.zeros(gi.shape)  # distances of the reported neighbours, accumulated in the reference's order (RealVectorStateSpace.cpp:230-242)
    for j in range(db.shape[1]):
        diff = db[gi, j] - db[ids, j][:, None]
        gd2 += diff * diff
    gd = np.sqrt(gd2)
    same_order = (gi == ri).all(axis=1)
    same_set = np.array([set(a) == set(b) for a, b in zip(gi, ri)])
    rel = np.abs(gd - rd) / np.maximum(rd, 1e-300)
    # a row whose id sets differ counts as tie-only when its sorted distances still agree to 1e-6 relative (equidistant alternatives)
    tie_only = (~same_set) & (rel.max(axis=1) < 1e-6)
    out.update(knn_queries=int(len(ids)), identical_sets=int(same_set.sum()), identical_order=int(same_order.sum()),
               tie_only=int(tie_only.sum()), differ_sets_not_tie=int((~same_set & ~tie_only).sum()), max_rel_dist_err=float(rel.max()),
               reference_knn_queries_per_s=len(ids) / dtk)
    return out


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version at init), so the real
    stdout is set aside for the result line and file descriptor 1 is pointed at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--roadmap-n", type=int, default=1 << 20, help="samples of the PRM roadmap build (0 = skip)")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact config 4 / config 5 blocks")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference check of the built roadmap (all edges + kNN sample, ~1 min of CPU)")
    ap.add_argument("--parity-edges", type=int, default=0, help="check only the first N roadmap edges against the reference (0 = all)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ompl_b200
    from ompl_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = ompl_b200.Context(local)
    scene, q0, q1 = make_inputs(rank)
    env = synth.fill_env(ompl_b200.Environment(ctx), scene)

    d0, d1 = torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev)
    dv = torch.zeros(M_EDGES, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_pass(n_steps, a0, a1, out, m):
        """returns summed kernel-only milliseconds over n_steps (L2 flushed before each)"""
        total = 0.0
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(n_steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.validate_motions_ptr("panda", env, a0.data_ptr(), a1.data_ptr(), m, out.data_ptr(), stream=stream.cuda_stream)
                e1.record(stream)
                evs.append((e0, e1))
            stream.synchronize()
        for e0, e1 in evs:
            total += e0.elapsed_time(e1)
        return total

    # ---- device-resident throughput (value)
    device_pass(args.warmup, d0, d1, dv, M_EDGES)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = ctx.launch_count
    ms = device_pass(args.steps, d0, d1, dv, M_EDGES)
    barrier()
    launches = ctx.launch_count - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * M_EDGES * args.steps / (ms_max * 1e-3)
    valid_frac = float(dv.float().mean().item())

    # ---- end to end through the host C ABI with pinned host buffers
    h0, h1 = torch.from_numpy(q0).pin_memory(), torch.from_numpy(q1).pin_memory()
    hv = torch.zeros(M_EDGES, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        ctx.validate_motions_ptr("panda", env, h0.data_ptr(), h1.data_ptr(), M_EDGES, hv.data_ptr(), device=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.validate_motions_ptr("panda", env, h0.data_ptr(), h1.data_ptr(), M_EDGES, hv.data_ptr(), device=False)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * M_EDGES * args.steps / float(t.item())
    assert np.array_equal(hv.numpy(), dv.cpu().numpy()), "host and device entry points disagree"
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roadmap build (all ranks): BASELINE config 3, same scene, through the C library's collective entry point
    # (b200mp_prm_build_multi: NCCL exchanges inside the library; torch.distributed only carries the NCCL unique id)
    roadmap, rm_host = None, None
    if args.roadmap_n > 0:
        from ompl_b200.roadmap import build_roadmap, make_comm
        comm = make_comm(ctx)
        build_roadmap(comm, env, "panda", args.roadmap_n, 16, 42, want_arrays=False)  # warm-up = the timed build itself (grow-only buffers, memory pool, NCCL channels)
        barrier()
        l0 = ctx.launch_count
        rm = build_roadmap(comm, env, "panda", args.roadmap_n, 16, 42, want_arrays=False)
        barrier()
        st = dict(rm.stats)
        launches_rm = int(ctx.launch_count - l0)
        keys = ["ms_sample", "ms_knn", "ms_edges", "ms_validate", "ms_gather", "ms_total"]
        t = torch.tensor([st[k] for k in keys], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for k, v in zip(keys, t.tolist()):
            st[k] = round(v, 3)
        # the same build once more, end to end: wall clock around the C call, the whole roadmap copied to rank 0's host memory
        barrier()
        t0 = time.perf_counter()
        rm = build_roadmap(comm, env, "panda", args.roadmap_n, 16, 42, want_arrays=(rank == 0), want_neighbors=(rank == 0))
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert rm.stats["n_edges"] == st["n_edges"] and rm.stats["n_valid_edges"] == st["n_valid_edges"], "roadmap builds disagree"
        roadmap = {"workload": f"{args.roadmap_n} valid Panda samples, k=16, C2 scene, b200mp_prm_build_multi, 1 warm-up build", "n_gpus": world,
                   "scaling": "strong", "build_ms": st["ms_total"], "build_e2e_ms": round(e2e_ms, 3), "gpu_launches": launches_rm}
        roadmap.update({k: st[k] for k in ("n_states", "n_candidates", "n_edges", "n_valid_edges", "ms_sample", "ms_knn", "ms_edges",
                                           "ms_validate", "ms_gather")})
        if rank == 0 and not args.no_parity:
            rm_host = tuple(x.numpy() for x in (rm.states, rm.edges, rm.valid, rm.neighbors))
        del rm
        comm.close()

    if world > 1:
        dist.destroy_process_group()  # nothing below is collective: rank 0 alone finishes the report
    if rank != 0:
        return

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gpu_launches": int(launches), "valid_fraction": valid_frac,
            "states_per_sec_upper": value * 64, "roadmap_build_ms": roadmap["build_ms"] if roadmap else None,
            "config": workload_config(world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(q0.nbytes + q1.nbytes), "d2h_bytes_per_step": int(M_EDGES)},
            "clocks": sampler.summary()}
    cores = os.cpu_count() or 1
    if world == 1:
        # ---- roofline of the validation kernel on the all-valid sub-workload (no early exit hides work) + CPU baseline
        from oracle import port
        pe = synth.fill_env(port.PortEnvironment(), scene)
        valid_idx = np.nonzero(dv.cpu().numpy() == 1)[0]
        peak = ctx.fp32_peak_tflops(4096)
        samp = valid_idx[:512]
        _, fl = pe.validate_motions("panda", q0[samp], q1[samp], nthreads=cores, want_flops=True)
        w_ref_edge = float(fl.mean())     # flops the reference's sequential semantics execute per valid motion (oracle walker)
        rep = np.resize(valid_idx, 1 << 17)
        a0, a1 = torch.from_numpy(q0[rep]).to(dev), torch.from_numpy(q1[rep]).to(dev)
        av = torch.zeros(len(rep), dtype=torch.uint8, device=dev)
        device_pass(3, a0, a1, av, len(rep))
        l1 = ctx.launch_count
        ms_av = device_pass(args.steps, a0, a1, av, len(rep)) / args.steps
        assert ctx.launch_count - l1 >= args.steps
        assert bool(av.all()), "all-valid sub-workload must validate as valid"
        w_ref = w_ref_edge * len(rep) / (ms_av * 1e-3) / 1e12
        # executed flops per launch of exactly this workload, counted by ncu (profiles/validate_flops.json)
        rec = flops_record(len(rep), int(len(valid_idx)))
        w_exec_launch, src, traffic = (float(rec["flops_per_launch"]), rec.get("source"), rec.get("dram_bytes")) if rec else (None, None, None)
        achieved = (w_exec_launch / (ms_av * 1e-3) / 1e12) if w_exec_launch else w_ref
        # the dense variant SURVEY 8(d) asks for beside it: the same motions through the ordered scan (sorted scene, candidate grid off),
        # i.e. every bounding sphere against every obstacle the reference's scan reaches
        env_dense = synth.fill_env(ompl_b200.Environment(ctx), scene, sort=True, grid=False)
        ctx.set_option(1, 0)  # the ordered-scan kernel alone: no first look (its sample passes are dear in a 1000-obstacle scan)
        dense_ms = []
        for _ in range(3):
            evd = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            with torch.cuda.stream(stream):
                flush.fill_(1)
                evd[0].record(stream)
                ctx.validate_motions_ptr("panda", env_dense, a0.data_ptr(), a1.data_ptr(), len(rep), av.data_ptr(), stream=stream.cuda_stream)
                evd[1].record(stream)
            stream.synchronize()
            dense_ms.append(evd[0].elapsed_time(evd[1]))
        assert bool(av.all()), "dense variant must agree on the all-valid sub-workload"
        ctx.set_option(1, 1)
        ms_dense = float(np.median(dense_ms[1:]))
        dflops = float(rec["dense"]["flops_per_launch"]) if rec and rec.get("dense") else None
        dense_t = (dflops if dflops else w_ref_edge * len(rep)) / (ms_dense * 1e-3) / 1e12
        del env_dense
        line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                            "kernel": "validate_motions_kernel<Panda, sorted, grid>",
                            "achieved_is": ("executed FP32 flops/launch (ncu FFMA x2 + FADD + FMUL, " + str(src) + ")")
                            if w_exec_launch else "W_ref (no ncu flop count matching this build found)",
                            "w_ref_tflops": w_ref, "w_ref_flops_per_edge": w_ref_edge,
                            "workload": f"all-valid: {len(valid_idx)} valid motions of the batch tiled to 2^17",
                            "ms_per_launch": ms_av, "states_per_sec": len(rep) * 64 / (ms_av * 1e-3),
                            "dense": {"kernel": "validate_motions_kernel<Panda, sorted, ordered scan>", "ms_per_launch": ms_dense,
                                      "achieved": dense_t, "frac": dense_t / peak, "states_per_sec": len(rep) * 64 / (ms_dense * 1e-3),
                                      "achieved_is": "executed FP32 flops (ncu)" if dflops else "W_ref (oracle-counted reference flops; no ncu count for this build)"},
                            "peak_source": "b200mp_fp32_peak FFMA microbenchmark on this GPU (nominal 148*128*2*1.965 GHz = 74.4)"}
        # what actually bounds the grid variant: instruction issue. Executed warp instructions per launch (ncu smsp__inst_executed.sum, same
        # record) / time against the issue slots of the chip (SMs x 4 schedulers x SM clock).
        if rec and rec.get("inst_executed"):
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
            slots = sm_count * 4 * 1.965e9
            for name, tgt, ms_k, r_ in (("issue", line["roofline"], ms_av, rec), ("issue", line["roofline"]["dense"], ms_dense, rec.get("dense") or {})):
                if r_.get("inst_executed"):
                    rate = float(r_["inst_executed"]) / (ms_k * 1e-3)
                    tgt[name] = {"warp_inst_per_launch": float(r_["inst_executed"]), "achieved_ginst_s": rate / 1e9, "peak_ginst_s": slots / 1e9,
                                 "frac": rate / slots, "peak_is": f"{sm_count} SMs x 4 schedulers x 1.965 GHz"}
        if not args.no_cpu_baseline:
            chk, kind = checker_env(scene)
            sample = 1 << 18
            chk.validate_motions("panda", q0[:8192], q1[:8192], nthreads=cores)
            t0 = time.perf_counter()
            want = chk.validate_motions("panda", q0[:sample], q1[:sample], nthreads=cores)
            dt = time.perf_counter() - t0
            got = dv.cpu().numpy()[:sample]
            mism = int((got != want).sum())
            # the same reference on the all-valid sub-workload the roofline is quoted on (no early exit on either arm)
            nv = min(len(valid_idx), 4096)
            t0 = time.perf_counter()
            wv = chk.validate_motions("panda", q0[valid_idx[:nv]], q1[valid_idx[:nv]], nthreads=cores)
            dtv = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first 2^18 of the 2^20 motions, all host threads; {mism} decisions differ from the GPU",
                                    "all_valid_edges_per_s": nv / dtv, "all_valid_sample": f"{nv} valid motions, {int((wv != 1).sum())} differ"}
    if world == 1 and not args.no_configs:
        # ---- BASELINE configs 4 and 5, compact (tools/configs_probe.py holds the full probes): throughput, CPU checker beside it, mismatches
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import configs_probe
        c4 = configs_probe.c4(ctx, 1 << 20)
        line["c4"] = {"workload": "Fetch, 2^20 motions x 64 states vs 2^20-point shelf cloud (filtered -> point-cloud tree)",
                      "edges_per_s": c4["edges_per_s"], "all_valid_states_per_s": c4.get("all_valid_states_per_s"), "ms_filter_gpu": c4["ms_filter_gpu"],
                      "ms_filter_reference_1thread": c4["ms_filter_reference_1thread"], "filtered_points": c4["filtered_points"],
                      "filtered_points_reference": c4["filtered_points_reference"], "cpu_baseline": c4["cpu_baseline"],
                      "reference_on_own_filter": c4.get("reference_pipeline_on_reference_filtered_cloud"),
                      "reference_filter_stable_ties": c4.get("reference_filter_with_stable_ties")}
        c5 = configs_probe.c5(ctx, 1 << 16)
        line["c5"] = {"workload": "20-link planar chain, horn environment: 2^20 states, chain-metric kNN k=16 over 2^16 valid states, its edges",
                      "states_per_s": c5["states_per_s"], "cpu_states": c5["cpu_states"], "knn_queries_per_s": c5["knn_queries_per_s"],
                      "cpu_knn": c5["cpu_knn"], "edges_per_s": c5["edges_per_s"], "cpu_motions": c5["cpu_motions"]}
    if roadmap is not None:
        if rm_host is not None:
            from oracle import refvamp, refnn
            if refvamp.available() and refnn.available():
                roadmap["parity"] = roadmap_parity(*rm_host, scene, cores, max_edges=args.parity_edges)
                line["roadmap_parity_differ_outside"] = roadmap["parity"]["differ_outside"]
                roadmap["cpu_knn"] = {"value": roadmap["parity"].pop("reference_knn_queries_per_s"), "unit": "queries/s at 2^20 states", "cores": cores,
                                      "kind": "reference", "sample": f"{roadmap['parity']['knn_queries']} queries, NearestNeighborsLinear via std::function"}
                roadmap["cpu_edges"] = {"value": roadmap["parity"].pop("reference_edges_per_s"), "unit": "edges/s", "cores": cores, "kind": "reference",
                                        "sample": "all roadmap edges"}
            else:
                roadmap["parity"] = {"skipped": "oracle/_ref not built"}
        line["roadmap"] = roadmap  # last: the driver keeps the tail of the line
    print(json.dumps(line), file=claim_stdout(), flush=True)


def flops_record(motions, unique_valid):
    """The ncu flop counts of the all-valid launches (profiles/validate_flops.json) -- only if they were taken on the kernel code of
    THIS build (sha256 of the validation kernel sources recorded next to them) and on this workload, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "validate_flops.json")) as f:
            pj = json.load(f)
        if pj.get("motions") == motions and pj.get("unique_valid") == unique_valid and pj.get("kernel_sha256") == kernel_hash():
            return pj
    except Exception:
        pass
    return None


def kernel_hash():
    """sha256 over the sources the validation kernels are compiled from (a changed kernel invalidates recorded ncu counts)."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "ompl_b200", "csrc")
    for rel in ("validate.cuh", "robot_common.cuh", "robots/panda_gen.cuh"):
        with open(os.path.join(csrc, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


if __name__ == "__main__":
    main()
