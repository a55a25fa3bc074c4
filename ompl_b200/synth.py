"""Seeded synthetic scenes and motion sets (numpy only; the same arrays feed the engine and the checkers).

Recipes follow SURVEY.md 8(d): C1 = the 14-sphere cage of demos/vamp/VAMPPlanning.cpp:43-47,87-88;
C2 = 500 spheres + 500 cuboids around the robot and edges whose length puts the reference at exactly
n = 8 rakes = 64 interpolated states (validate.hh:41: n = ceil(|dq| * 32 / 8)).
"""
from __future__ import annotations

import numpy as np

PANDA_LO = np.array([-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671], np.float32)
PANDA_HI = (PANDA_LO + np.array([5.9342, 3.6652, 5.9342, 3.2289, 5.9342, 3.9096, 5.9342], np.float32)).astype(np.float32)
FETCH_LO = np.array([0.0, -1.6056, -1.221, -3.14159, -2.251, -3.14159, -2.16, -3.14159], np.float32)
FETCH_HI = (FETCH_LO + np.array([0.38615, 3.2112, 2.739, 6.28318, 4.502, 6.28318, 4.32, 6.28318], np.float32)).astype(np.float32)
BOUNDS = {"panda": (PANDA_LO, PANDA_HI), "fetch": (FETCH_LO, FETCH_HI)}


def _bounds_from_model(robot: str):
    """lo = s_a, hi = s_a + s_m in float32, from the robot model the CUDA tables are generated from (ompl_b200/robots/*.json)."""
    import json
    import os
    m = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "robots", robot + ".json")))
    lo = np.asarray(m["s_a"], np.float32)
    return lo, (lo + np.asarray(m["s_m"], np.float32)).astype(np.float32)


for _r in ("ur5", "baxter"):
    BOUNDS[_r] = _bounds_from_model(_r)

CAGE_CENTRES = np.array([
    (0.55, 0, 0.25), (0.35, 0.35, 0.25), (0, 0.55, 0.25), (-0.55, 0, 0.25), (-0.35, -0.35, 0.25),
    (0, -0.55, 0.25), (0.35, -0.35, 0.25), (0.35, 0.35, 0.8), (0, 0.55, 0.8), (-0.35, 0.35, 0.8),
    (-0.55, 0, 0.8), (-0.35, -0.35, 0.8), (0, -0.55, 0.8), (0.35, -0.35, 0.8)], np.float32)
CAGE_START = np.array([0., -0.785, 0., -2.356, 0., 1.571, 0.785], np.float32)
CAGE_GOAL = np.array([2.35, 1., 0., -0.8, 0., 2.5, 0.785], np.float32)


def cage_scene(radius: float = 0.2) -> dict:
    xyzr = np.concatenate([CAGE_CENTRES, np.full((len(CAGE_CENTRES), 1), radius, np.float32)], axis=1)
    return dict(spheres=xyzr.astype(np.float32))


def _euler_xyz_axes(rho, theta, phi):
    """Rz(phi) Ry(theta) Rx(rho) columns = cuboid axes (factory.hh:26-62), float32."""
    cr, sr = np.cos(rho), np.sin(rho)
    cp, sp = np.cos(theta), np.sin(theta)
    cy, sy = np.cos(phi), np.sin(phi)
    rot = np.empty(rho.shape + (3, 3), np.float64)
    rot[..., 0, 0] = cy * cp
    rot[..., 0, 1] = cy * sp * sr - sy * cr
    rot[..., 0, 2] = cy * sp * cr + sy * sr
    rot[..., 1, 0] = sy * cp
    rot[..., 1, 1] = sy * sp * sr + cy * cr
    rot[..., 1, 2] = sy * sp * cr - cy * sr
    rot[..., 2, 0] = -sp
    rot[..., 2, 1] = cp * sr
    rot[..., 2, 2] = cp * cr
    return rot


def clutter_scene(seed: int = 1234, n_spheres: int = 500, n_cuboids: int = 500, n_capsules: int = 0,
                  n_zcuboids: int = 0, n_zcapsules: int = 0, r_lo: float = 0.01, r_hi: float = 0.05,
                  rho_lo: float = 0.35, rho_hi: float = 1.30, z_lo: float = -0.2, z_hi: float = 1.2) -> dict:
    """C2 scene: obstacle centres cylindrical rho in U[0.35,1.30], theta in U[0,2pi), z in U[-0.2,1.2]."""
    rng = np.random.default_rng(seed)

    def centres(n):
        rho = rng.uniform(rho_lo, rho_hi, n)
        th = rng.uniform(0, 2 * np.pi, n)
        z = rng.uniform(z_lo, z_hi, n)
        return np.stack([rho * np.cos(th), rho * np.sin(th), z], axis=1)

    out = {}
    c = centres(n_spheres)
    out["spheres"] = np.concatenate([c, rng.uniform(r_lo, r_hi, (n_spheres, 1))], axis=1).astype(np.float32)

    def cuboids(n, z_aligned):
        c = centres(n)
        e = rng.uniform(-np.pi, np.pi, (n, 3))
        if z_aligned:
            e[:, 0] = 0
            e[:, 1] = 0
        rot = _euler_xyz_axes(e[:, 0], e[:, 1], e[:, 2])
        h = rng.uniform(r_lo, r_hi, (n, 3))
        # centre, axis1 (column 0), axis2, axis3, half extents
        return np.concatenate([c, rot[:, :, 0], rot[:, :, 1], rot[:, :, 2], h], axis=1).astype(np.float32)

    def capsules(n, z_aligned):
        c = centres(n)
        v = rng.normal(size=(n, 3))
        if z_aligned:
            v[:, 0] = 0
            v[:, 1] = 0
        v = v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0.05, 0.3, (n, 1))
        return np.concatenate([c - v / 2, v, rng.uniform(r_lo, r_hi, (n, 1))], axis=1).astype(np.float32)

    out["cuboids"] = cuboids(n_cuboids, False)
    if n_capsules:
        out["capsules"] = capsules(n_capsules, False)
    if n_zcuboids:
        out["z_cuboids"] = cuboids(n_zcuboids, True)
    if n_zcapsules:
        out["z_capsules"] = capsules(n_zcapsules, True)
    return out


def fill_env(env, scene: dict, sort: bool = True, grid: bool | None = None):
    """Push a scene dict into any environment object exposing add_spheres/add_cuboids/add_capsules/sort."""
    if grid is not None:
        env.set_grid(grid)
    if "spheres" in scene and len(scene["spheres"]):
        env.add_spheres(scene["spheres"])
    if "capsules" in scene and len(scene["capsules"]):
        env.add_capsules(scene["capsules"], z_aligned=False)
    if "z_capsules" in scene and len(scene["z_capsules"]):
        env.add_capsules(scene["z_capsules"], z_aligned=True)
    if "cuboids" in scene and len(scene["cuboids"]):
        env.add_cuboids(scene["cuboids"], z_aligned=False)
    if "z_cuboids" in scene and len(scene["z_cuboids"]):
        env.add_cuboids(scene["z_cuboids"], z_aligned=True)
    if "heightfield" in scene:
        env.add_heightfield(scene["heightfield_centre"], scene["heightfield_scale"], scene["heightfield"])
    if sort:
        env.sort()
    if hasattr(env, "finalize"):
        env.finalize()
    return env


def uniform_states(robot: str, n: int, seed: int) -> np.ndarray:
    lo, hi = BOUNDS[robot]
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (n, len(lo))).astype(np.float32)


def fixed_length_edges(robot: str, m: int, seed: int, len_lo: float = 1.75, len_hi: float = 2.0):
    """q0 uniform in bounds, direction uniform on the sphere, |q1-q0| in (len_lo, len_hi] (fp32 length,
    so the reference's n = ceil(len*4) is 8 for the defaults), resampled until q1 is inside the bounds."""
    lo, hi = BOUNDS[robot]
    d = len(lo)
    rng = np.random.default_rng(seed)
    q0 = np.empty((m, d), np.float32)
    q1 = np.empty((m, d), np.float32)
    todo = np.arange(m)
    while todo.size:
        k = todo.size
        a = rng.uniform(lo, hi, (k, d))
        v = rng.normal(size=(k, d))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        ln = len_hi - rng.uniform(0, len_hi - len_lo, (k, 1))
        b = a + v * ln
        a32 = a.astype(np.float32)
        b32 = b.astype(np.float32)
        l32 = np.sqrt(((b32 - a32).astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
        ok = np.all((b32 >= lo) & (b32 <= hi), axis=1) & (l32 > np.float32(len_lo) + 1e-4) & (l32 <= np.float32(len_hi) - 1e-4)
        q0[todo[ok]] = a32[ok]
        q1[todo[ok]] = b32[ok]
        todo = todo[~ok]
    return q0, q1


def local_edges(robot: str, starts: np.ndarray, seed: int, len_lo: float = 0.05, len_hi: float = 2.5):
    """Edges leaving given start states in random directions with random lengths (clipped to bounds)."""
    lo, hi = BOUNDS[robot]
    rng = np.random.default_rng(seed)
    m, d = starts.shape
    v = rng.normal(size=(m, d))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ln = rng.uniform(len_lo, len_hi, (m, 1))
    q1 = np.clip(starts + v * ln, lo, hi).astype(np.float32)
    return starts.astype(np.float32), q1


def shelf_boxes(robot: str = "fetch"):
    """Bookshelf-style scene (after the MotionBenchMaker `bookshelf_tall` problems): two side panels, a back panel and
    four shelves in front of the robot. Returns [(centre, half extents)] of axis-aligned boxes."""
    x0, w, z0 = (0.95, 0.45, 0.35) if robot == "fetch" else (0.62, 0.40, 0.05)
    depth, height, t = 0.18, 1.2, 0.015
    boxes = [((x0, -w, z0 + height / 2), (depth, t, height / 2)), ((x0, w, z0 + height / 2), (depth, t, height / 2)),
             ((x0 + depth, 0.0, z0 + height / 2), (t, w, height / 2))]
    for i in range(4):
        boxes.append(((x0, 0.0, z0 + 0.05 + i * (height - 0.1) / 3), (depth, w, t)))
    return [(np.array(c, np.float64), np.array(h, np.float64)) for c, h in boxes]


def shelf_cloud(n: int, seed: int = 7, robot: str = "fetch") -> np.ndarray:
    """n points sampled uniformly (area weighted) on the surfaces of shelf_boxes(robot), float32 [n][3]."""
    rng = np.random.default_rng(seed)
    faces = []  # (centre, half, fixed axis, sign)
    for c, h in shelf_boxes(robot):
        for ax in range(3):
            o = [a for a in range(3) if a != ax]
            area = 4 * h[o[0]] * h[o[1]]
            for sg in (-1.0, 1.0):
                faces.append((c, h, ax, sg, area))
    areas = np.array([f[4] for f in faces])
    which = rng.choice(len(faces), size=n, p=areas / areas.sum())
    u = rng.uniform(-1, 1, (n, 3))
    pts = np.empty((n, 3))
    for i, f in enumerate(which):
        c, h, ax, sg, _ = faces[f]
        p = c + u[i] * h
        p[ax] = c[ax] + sg * h[ax]
        pts[i] = p
    return pts.astype(np.float32)


def terrain(seed: int = 4, nx: int = 64, ny: int = 48, size: float = 4.0, z0: float = -0.6) -> dict:
    """A rolling height field (heights 0 .. 0.55 m above z0) covering size x size metres around the robot, as scene-dict
    entries for fill_env (factory::heightfield::flat arguments: centre, cell scale, heights[ny][nx])."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    h = (0.25 + 0.25 * np.sin(xx / 5.0) * np.cos(yy / 4.0) + 0.05 * rng.uniform(size=(ny, nx))).astype(np.float32)
    return {"heightfield": h, "heightfield_centre": np.array([0.1, -0.05, z0], np.float32),
            "heightfield_scale": np.array([size / nx, size / ny, 1.0], np.float32)}
