#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref (run where /root/reference
exists: `make -C oracle ref && python tests/golden/make_golden.py`). Each file stores the inputs (scene arrays,
states, motions) together with the reference's answers, so the fixtures do not depend on numpy's RNG staying stable.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ompl_b200 import synth  # noqa: E402
from oracle import refvamp  # noqa: E402


def ref_nn():
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_nn.so"))
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    L.refnn_create.restype = C.c_void_p
    L.refnn_create.argtypes = [f64p, C.c_int64, C.c_int]
    L.refnn_destroy.argtypes = [C.c_void_p]
    L.refnn_nearest_k.restype = C.c_int64
    L.refnn_nearest_k.argtypes = [C.c_void_p, f64p, C.c_int, i64p, f64p]
    L.refnn_nearest_r.restype = C.c_int64
    L.refnn_nearest_r.argtypes = [C.c_void_p, f64p, C.c_double, i64p, f64p, C.c_int64]
    return L


def validation_case(name, robot, scene, sort, n_states, n_motions, seed, len_hi=2.5):
    env = synth.fill_env(refvamp.RefEnvironment(), scene, sort=sort)
    q = synth.uniform_states(robot, n_states, seed)
    sv = env.validate_states(robot, q)
    starts = q[sv == 1]
    if len(starts) < n_motions // 2:
        starts = np.concatenate([starts, q[: n_motions // 2]])
    starts = np.resize(starts, (n_motions // 2, q.shape[1]))
    a0, a1 = synth.local_edges(robot, starts, seed + 1, 0.05, len_hi)
    b0 = synth.uniform_states(robot, n_motions - len(a0), seed + 2)
    b1 = synth.uniform_states(robot, n_motions - len(a0), seed + 3)
    q0 = np.concatenate([a0, b0]).astype(np.float32)
    q1 = np.concatenate([a1, b1]).astype(np.float32)
    mv = env.validate_motions(robot, q0, q1)
    out = dict(robot=np.array(robot), sorted=np.array(sort), states=q, states_valid=sv, q0=q0, q1=q1, motions_valid=mv)
    for k, v in scene.items():
        out["scene_" + k] = v
    for kind in range(5):
        out[f"min_d_{kind}"] = env.min_distances(kind)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: states valid {sv.mean():.3f}, motions valid {mv.mean():.3f}")


def cloud_case(name, robot, radii, n_pts, scene, n_states, n_motions, seed):
    """Point cloud (the reference builds its own CAPT from the raw points) + primitives."""
    pts = synth.shelf_cloud(n_pts, seed=seed, robot=robot)
    env = refvamp.RefEnvironment()
    env.add_pointcloud(pts, *radii)
    synth.fill_env(env, scene)
    q = synth.uniform_states(robot, n_states, seed)
    sv = env.validate_states(robot, q)
    starts = np.resize(q[sv == 1], (n_motions, q.shape[1]))
    q0, q1 = synth.local_edges(robot, starts, seed + 1, 0.05, 2.0)
    mv = env.validate_motions(robot, q0, q1)
    t = env.capt_export()
    out = dict(robot=np.array(robot), points=pts, radii=np.array(radii, np.float32), states=q, states_valid=sv, q0=q0, q1=q1,
               motions_valid=mv, capt_nlog2=np.array(t["nlog2"]), capt_tests=t["tests"], capt_aff_starts=t["aff_starts"],
               capt_top=t["top"], capt_aabbs=t["aabbs"], capt_aff_sum=np.where(np.isfinite(t["aff"]), t["aff"], 0).sum(axis=1, dtype=np.float64))
    for k, v in scene.items():
        out["scene_" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: states valid {sv.mean():.3f}, motions valid {mv.mean():.3f}, affordance records {t['aff'].shape[1] // 8}")


def chain_case():
    """Planar 20-link chain in the horn environment (KinematicChainBenchmark.cpp:48-49): answers of the UNMODIFIED
    demos/KinematicChain.h + DiscreteMotionValidator.cpp (oracle/_ref/libref_chain.so)."""
    from oracle import chain
    n = 20
    env = chain.RefChain.horn(n, np.log(n) / n)
    R = chain.RefChain(n, 1.0 / n, env)
    rng = np.random.default_rng(606)
    q = rng.uniform(-np.pi, np.pi, (6000, n))
    q[:3000] *= rng.uniform(0.02, 0.35, (3000, 1))      # gently curled chains: mostly valid
    # chains that follow the horn (each wall segment turns by pi / n), with growing noise
    q[3000:5000] = np.pi / n * rng.uniform(0.5, 1.5, (2000, 1)) + rng.normal(scale=1.0, size=(2000, n)) * rng.uniform(0.0, 0.12, (2000, 1))
    q[3000:5000, 0] = rng.normal(scale=0.05, size=2000)
    sv = R.validate_states(q)
    ok = q[sv == 1]
    a = ok[rng.integers(0, len(ok), 1500)]
    b = np.concatenate([ok[rng.integers(0, len(ok), 1000)], q[rng.integers(0, len(q), 500)]])
    b[:300] = a[:300] + rng.normal(scale=0.02, size=(300, n))   # short motions
    v0, _, nd = R.check_motions(a, b, 0)
    v1, last, _ = R.check_motions(a, b, 1)
    t = rng.uniform(0, 1, len(a))
    np.savez_compressed(os.path.join(HERE, "chain.npz"), n=np.array(n), env=env, lvs=np.array(R.longest_valid_segment), states=q, states_valid=sv,
                        a=a, b=b, motions_valid=v0, motions_valid_form1=v1, last_t=last, nd=nd, dist=R.distance(a, b), t=t,
                        interp=R.interpolate(a, b, t))
    print(f"chain: states valid {sv.mean():.3f}, motions valid {v0.mean():.3f}, nd mean {nd.mean():.1f} max {nd.max()}")


ATTACH_SCENES = dict(
    panda=dict(seed=41, n_spheres=40, n_cuboids=40, n_capsules=8, n_zcuboids=8, n_zcapsules=8),
    panda_unsorted=dict(seed=42, n_spheres=30, n_cuboids=30, n_capsules=6, n_zcuboids=6, n_zcapsules=6),
    fetch=dict(seed=43, n_spheres=30, n_cuboids=30, n_capsules=6, n_zcuboids=6, n_zcapsules=6, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6),
    ur5=dict(seed=44, n_spheres=40, n_cuboids=40, n_capsules=8),
    baxter=dict(seed=45, n_spheres=6, n_cuboids=6, n_capsules=3),
)


def attachment(seed):
    """A grasped object: a rotated, offset pose relative to the end effector and a few spheres of mixed size."""
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(-0.6, 0.6, 2)
    rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    tf = np.concatenate([rz @ rx, rng.uniform(-0.03, 0.03, (3, 1)) + np.array([[0], [0], [0.06]])], axis=1).astype(np.float32)
    n = int(rng.integers(3, 9))
    sph = np.concatenate([rng.uniform(-0.08, 0.08, (n, 2)), rng.uniform(0.0, 0.3, (n, 1)), rng.uniform(0.015, 0.07, (n, 1))], axis=1).astype(np.float32)
    return tf, sph


def attachment_cases():
    """Robot::fkcc_attach of the unmodified reference (collision/attachments.hh, robots/<robot>.hh) through oracle/_ref, whose
    Eigen stand-in (oracle/shim/Eigen) executes Transform * Transform and Transform * Vector3 with plain float arithmetic."""
    out = {}
    for tag, kw in ATTACH_SCENES.items():
        robot = tag.split("_")[0]
        scene = synth.clutter_scene(**kw)
        is_sorted = not tag.endswith("unsorted")
        env = synth.fill_env(refvamp.RefEnvironment(), scene, sort=is_sorted)
        tf, sph = attachment(kw["seed"])
        q = synth.uniform_states(robot, 3000, kw["seed"] + 1)
        bare = env.validate_states(robot, q)
        env.set_attachment(tf, sph)
        sv = env.validate_states(robot, q)
        starts = np.resize(q[sv == 1], (1200, q.shape[1]))
        q0, q1 = synth.local_edges(robot, starts, kw["seed"] + 2, 0.05, 2.5)
        mv = env.validate_motions(robot, q0, q1)
        out.update({f"{tag}_tf": tf, f"{tag}_spheres": sph, f"{tag}_states": q, f"{tag}_states_valid_bare": bare, f"{tag}_states_valid": sv,
                    f"{tag}_q0": q0, f"{tag}_q1": q1, f"{tag}_motions_valid": mv, f"{tag}_eefk": refvamp.eefk(robot, q[:64])})
        print(f"attach {tag}: states valid {bare.mean():.3f} -> {sv.mean():.3f} with the attachment, motions valid {mv.mean():.3f}")
    np.savez_compressed(os.path.join(HERE, "attach.npz"), **out)


def main():
    if "--attachments" in sys.argv:
        attachment_cases()
        return
    if "--chain" in sys.argv:
        chain_case()
        return
    if "--cloud" in sys.argv:
        cloud_case("fetch_cloud", "fetch", (0.012, 0.24, 0.0025), 1500,
                   synth.clutter_scene(11, 20, 20, rho_lo=0.6, rho_hi=1.4, z_lo=0.2, z_hi=1.6), 1500, 800, 505)
        return
    if "--heightfield" in sys.argv:  # terrain under Panda, alone and together with primitives (validity.hh:132-138)
        validation_case("panda_terrain", "panda", synth.terrain(), True, 3000, 1500, 808)
        sc = synth.clutter_scene(14, 40, 40, n_capsules=10)
        sc.update(synth.terrain(seed=5))
        validation_case("fetch_terrain_clutter", "fetch", sc, True, 2000, 1000, 909)
        return
    if "--more-robots" in sys.argv:  # UR5 (6-DOF chain) and Baxter (two 7-DOF arms, resolution 64); written without touching the older files
        validation_case("ur5_clutter", "ur5", synth.clutter_scene(12, 60, 60, n_capsules=20, n_zcuboids=20, n_zcapsules=20), True, 2000, 1000, 606)
        validation_case("baxter_clutter", "baxter", synth.clutter_scene(13, 10, 10, n_capsules=4, n_zcuboids=4, n_zcapsules=4), True, 2000, 1000, 707)
        for robot in ("ur5", "baxter"):
            q = synth.uniform_states(robot, 256, 55)
            np.savez_compressed(os.path.join(HERE, f"{robot}_sphere_fk.npz"), q=q, spheres=refvamp.sphere_fk(robot, q))
        return
    validation_case("panda_cage", "panda", synth.cage_scene(), True, 3000, 1500, 101)
    validation_case("panda_clutter", "panda",
                    synth.clutter_scene(7, 120, 120, n_capsules=40, n_zcuboids=40, n_zcapsules=40), True, 3000, 1500, 202)
    validation_case("panda_clutter_unsorted", "panda",
                    synth.clutter_scene(8, 80, 80, n_capsules=20, n_zcuboids=20, n_zcapsules=20), False, 2000, 1000, 303)
    validation_case("fetch_clutter", "fetch",
                    synth.clutter_scene(9, 60, 60, n_capsules=20, n_zcuboids=20, n_zcapsules=20, rho_lo=0.6, rho_hi=1.4,
                                        z_lo=0.2, z_hi=1.6), True, 2000, 1000, 404)
    # known answers of the demo (demos/vamp/VAMPPlanning.cpp:43-47,87-88)
    env = synth.fill_env(refvamp.RefEnvironment(), synth.cage_scene())
    assert list(env.validate_states("panda", [synth.CAGE_START, synth.CAGE_GOAL])) == [1, 1]
    assert list(env.validate_motions("panda", [synth.CAGE_START], [synth.CAGE_GOAL])) == [0]
    # FK of the fine spheres
    for robot in ("panda", "fetch"):
        q = synth.uniform_states(robot, 256, 55)
        np.savez_compressed(os.path.join(HERE, f"{robot}_sphere_fk.npz"), q=q, spheres=refvamp.sphere_fk(robot, q))
    # nearest neighbours: NearestNeighborsLinear<unsigned long> over RealVectorStateSpace::distance
    L = ref_nn()
    rng = np.random.default_rng(77)
    out = {}
    for dim, n, k in ((7, 500, 10), (8, 300, 16), (20, 200, 30), (3, 12, 40)):
        db = rng.uniform(-3, 3, (n, dim))
        qs = rng.uniform(-3, 3, (40, dim))
        qs[:5] = db[:5]  # queries that are database members (distance-0 self match)
        h = L.refnn_create(np.ascontiguousarray(db), n, dim)
        idx = np.full((len(qs), k), -1, np.int64)
        dist = np.full((len(qs), k), np.inf)
        cnt = np.zeros(len(qs), np.int64)
        ridx, rcnt = [], []
        for i, q in enumerate(qs):
            ii = np.zeros(k, np.int64)
            dd = np.zeros(k)
            c = L.refnn_nearest_k(h, np.ascontiguousarray(q), k, ii, dd)
            cnt[i] = c
            idx[i, :c] = ii[:c]
            dist[i, :c] = dd[:c]
            ii = np.zeros(n, np.int64)
            dd = np.zeros(n)
            c = L.refnn_nearest_r(h, np.ascontiguousarray(q), 2.0, ii, dd, n)
            ridx.append(ii[:c].copy())
            rcnt.append(c)
        L.refnn_destroy(h)
        tag = f"d{dim}"
        out[tag + "_db"], out[tag + "_q"], out[tag + "_k"] = db, qs, np.array(k)
        out[tag + "_idx"], out[tag + "_dist"], out[tag + "_cnt"] = idx, dist, cnt
        out[tag + "_rcnt"] = np.array(rcnt)
        out[tag + "_ridx"] = np.concatenate(ridx) if ridx else np.zeros(0, np.int64)
    np.savez_compressed(os.path.join(HERE, "knn_linear.npz"), **out)
    print("knn golden written")


if __name__ == "__main__":
    main()
