"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_ref/libref_vamp.so (the UNMODIFIED reference).

The library is the reference's own VAMP implementation (validate_motion / fkcc / CAPT, pinned
aa4f4b53) compiled from /root/reference by oracle/Makefile. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module; the product package
(ompl_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_vamp.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(LIB_PATH)
    L.refv_env_create.restype = C.c_void_p
    L.refv_env_destroy.argtypes = [C.c_void_p]
    L.refv_env_add_sphere.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    L.refv_env_add_cuboid.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.refv_env_add_capsule.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.refv_env_add_heightfield.argtypes = [C.c_void_p] + [C.c_float] * 6 + [C.c_int64, C.c_int64, _f32p]
    L.refv_env_add_pointcloud.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_float, C.c_float, C.c_float]
    L.refv_env_set_attachment.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int64]
    L.refv_env_sort.argtypes = [C.c_void_p]
    L.refv_env_finalize.argtypes = [C.c_void_p]
    L.refv_env_min_distances.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.refv_env_min_distances.restype = C.c_int64
    L.refv_capt_sizes.argtypes = [C.c_void_p, C.c_int, _i64p]
    L.refv_capt_export.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _u32p, _f32p, _f32p, _f32p]
    L.refv_capt_collides.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int64, _u8p]
    L.refv_filter_pointcloud.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, _f32p, _f32p, _f32p, C.c_int, _f32p, C.c_int64]
    L.refv_filter_pointcloud.restype = C.c_int64
    L.refv_set_stable_sort.argtypes = [C.c_int]
    for r in ("panda", "fetch", "ur5", "baxter"):
        getattr(L, f"refv_{r}_bounds").argtypes = [_f32p, _f32p]
        getattr(L, f"refv_{r}_validate_motions").argtypes = [
            C.c_void_p, _f32p, _f32p, C.c_int64, C.c_int, _u8p, C.c_int]
        getattr(L, f"refv_{r}_sphere_fk").argtypes = [_f32p, C.c_int64, _f32p]
        getattr(L, f"refv_{r}_eefk").argtypes = [_f32p, C.c_int64, _f32p]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefEnvironment:
    """vamp::collision::Environment<float> -> Environment<FloatVector<8>> (environment.hh:14-85)."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.refv_env_create())
        self._final = False

    def __del__(self):
        try:
            self._L.refv_env_destroy(self._h)
        except Exception:
            pass

    def add_spheres(self, xyzr):
        for s in _f32(xyzr).reshape(-1, 4):
            self._L.refv_env_add_sphere(self._h, *[float(v) for v in s])
        self._final = False

    def add_cuboids(self, c15, z_aligned=False):
        for c in _f32(c15).reshape(-1, 15):
            self._L.refv_env_add_cuboid(self._h, np.ascontiguousarray(c), int(z_aligned))
        self._final = False

    def add_capsules(self, c7, z_aligned=False):
        for c in _f32(c7).reshape(-1, 7):
            self._L.refv_env_add_capsule(self._h, np.ascontiguousarray(c), int(z_aligned))
        self._final = False

    def add_heightfield(self, centre, scale, heights):
        h = _f32(heights)
        self._L.refv_env_add_heightfield(self._h, *[float(v) for v in centre], *[float(v) for v in scale], h.shape[1], h.shape[0], h.reshape(-1))
        self._final = False

    def add_pointcloud(self, xyz, r_min, r_max, r_point):
        xyz = _f32(xyz).reshape(-1, 3)
        self._L.refv_env_add_pointcloud(self._h, xyz, xyz.shape[0], r_min, r_max, r_point)
        self._final = False

    def set_attachment(self, tf, xyzr):
        """Attachment<float>(tf) + spheres (collision/attachments.hh:13-56); tf = 3x4 [R | t] relative to the end effector.
        With an attachment validate_motion runs Robot::fkcc_attach (planning/validate.hh:43,58)."""
        xyzr = _f32(xyzr).reshape(-1, 4)
        self._L.refv_env_set_attachment(self._h, _f32(tf).reshape(-1)[:12].copy(), xyzr, xyzr.shape[0])
        self._final = False

    def clear_attachment(self):
        self._L.refv_env_set_attachment(self._h, np.zeros(12, np.float32), np.zeros((0, 4), np.float32), 0)
        self._final = False

    def sort(self):
        self._L.refv_env_sort(self._h)
        self._final = False

    def finalize(self):
        self._L.refv_env_finalize(self._h)
        self._final = True

    def min_distances(self, kind: int) -> np.ndarray:
        n = self._L.refv_env_min_distances(self._h, kind, None)
        out = np.zeros(n, dtype=np.float32)
        if n:
            self._L.refv_env_min_distances(self._h, kind, out.ctypes.data_as(C.c_void_p))
        return out

    def capt_export(self, idx=0) -> dict:
        sizes = np.zeros(4, dtype=np.int64)
        assert self._L.refv_capt_sizes(self._h, idx, sizes) == 0
        nlog2, n_tests, n_leaves, n_aff = [int(v) for v in sizes]
        tests = np.zeros(n_tests, np.float32)
        aabbs = np.zeros((n_leaves, 6), np.float32)
        starts = np.zeros(n_leaves + 1, np.uint32)
        aff = np.zeros((3, n_aff * 8), np.float32)
        top = np.zeros(6, np.float32)
        radii = np.zeros(3, np.float32)
        assert self._L.refv_capt_export(self._h, idx, tests, aabbs, starts, aff, top, radii) == 0
        return dict(nlog2=nlog2, tests=tests, aabbs=aabbs, aff_starts=starts, aff=aff, top=top,
                    r_min=float(radii[0]), r_max=float(radii[1]), r_point=float(radii[2]))

    def capt_collides(self, xyzr, idx=0) -> np.ndarray:
        xyzr = _f32(xyzr).reshape(-1, 4)
        out = np.zeros(xyzr.shape[0], np.uint8)
        assert self._L.refv_capt_collides(self._h, idx, xyzr, xyzr.shape[0], out) == 0
        return out

    def validate_motions(self, robot: str, q0, q1, resolution: int = 0, nthreads: int = 1) -> np.ndarray:
        """validate_motion<Robot, 8, Robot::resolution> (resolution 0/32) or <.., 1> (resolution 1)."""
        if not self._final:
            self.finalize()
        q0 = _f32(q0)
        q1 = _f32(q1)
        assert q0.shape == q1.shape and q0.ndim == 2
        out = np.zeros(q0.shape[0], dtype=np.uint8)
        rc = getattr(self._L, f"refv_{robot}_validate_motions")(
            self._h, q0, q1, q0.shape[0], resolution, out, nthreads)
        if rc != 0:
            raise RuntimeError(f"refv_{robot}_validate_motions -> {rc}")
        return out

    def validate_states(self, robot: str, q, nthreads: int = 1) -> np.ndarray:
        """VampStateValidityChecker::isValid = validate_motion<Robot, rake, 1>(q, q)
        (src/ompl/vamp/VampStateValidityChecker.h:37-41)."""
        return self.validate_motions(robot, q, q, resolution=1, nthreads=nthreads)


def bounds(robot: str):
    L = lib()
    d = int(getattr(L, f"refv_{robot}_dimension")())
    lo = np.zeros(d, np.float32)
    hi = np.zeros(d, np.float32)
    getattr(L, f"refv_{robot}_bounds")(lo, hi)
    return lo, hi


def sphere_fk(robot: str, q) -> np.ndarray:
    """Robot::sphere_fk<8>: [M][n_spheres][4] fine spheres (x, y, z, r)."""
    L = lib()
    q = _f32(q)
    ns = int(getattr(L, f"refv_{robot}_n_spheres")())
    out = np.zeros((q.shape[0], ns, 4), np.float32)
    getattr(L, f"refv_{robot}_sphere_fk")(q, q.shape[0], out)
    return out


def eefk(robot: str, q) -> np.ndarray:
    """Robot::eefk: [M][3][4] end-effector pose [R | t]."""
    q = _f32(q)
    out = np.zeros((q.shape[0], 3, 4), np.float32)
    getattr(lib(), f"refv_{robot}_eefk")(q, q.shape[0], out.reshape(-1))
    return out


def hardware_threads() -> int:
    return int(lib().refv_hardware_threads())


def filter_pointcloud(xyz, min_dist, max_range, origin, ws_min, ws_max, cull=True, stable_ties=False) -> np.ndarray:
    """vamp::collision::filter_pointcloud (collision/filter.hh:176-274), compiled as is. Its sort (pdqsort, absent here) is std::sort, or
    with stable_ties std::stable_sort: equal curve codes then keep their previous order."""
    xyz = _f32(xyz).reshape(-1, 3)
    lib().refv_set_stable_sort(1 if stable_ties else 0)
    out = np.empty((len(xyz) + 1, 3), np.float32)
    n = lib().refv_filter_pointcloud(xyz, len(xyz), float(min_dist), float(max_range), _f32(origin), _f32(ws_min), _f32(ws_max),
                                     int(bool(cull)), out, len(out))
    lib().refv_set_stable_sort(0)
    return out[:n].copy()
