"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/liboracle.so (the CPU restatement / "port").

Same surface as oracle/refvamp.py so tests can swap one for the other. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
ROBOT_ID = {"panda": 0, "fetch": 1, "ur5": 2, "baxter": 3}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.or_env_create.restype = C.c_void_p
    L.or_env_destroy.argtypes = [C.c_void_p]
    L.or_env_add_sphere.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    L.or_env_add_cuboid.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.or_env_add_capsule.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.or_env_add_heightfield.argtypes = [C.c_void_p] + [C.c_float] * 6 + [C.c_int64, C.c_int64, _f32p]
    L.or_env_add_capt.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _u32p, _f32p, C.c_int64, _f32p,
                                  C.c_float, C.c_float, C.c_float]
    L.or_env_set_attachment.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int64]
    L.or_env_sort.argtypes = [C.c_void_p]
    L.or_env_min_distances.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.or_env_min_distances.restype = C.c_int64
    L.or_validate_motions.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int64, C.c_int, _u8p, C.c_void_p, C.c_int]
    L.or_edge_min_clearance.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int64, C.c_int, _f64p, C.c_int]
    L.or_sphere_fk.argtypes = [C.c_int, _f32p, C.c_int64, _f32p]
    L.or_eefk.argtypes = [C.c_int, _f32p, C.c_int64, _f32p]
    L.or_robot_info.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4 + [_f32p, _f32p]
    L.or_trig.argtypes = [_f32p, C.c_int64, _f32p, _f32p]
    L.or_knn.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                         _i64p, _f64p, C.c_int]
    L.or_knn_radius.argtypes = [_f64p, C.c_int64, _f64p, C.c_int, C.c_double, C.c_int, C.c_double, _i64p, _f64p, C.c_int64]
    L.or_knn_radius.restype = C.c_int64
    L.orf_filter_pointcloud.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, _f32p, _f32p, _f32p, C.c_int, _f32p, C.c_int64]
    L.orf_filter_pointcloud.restype = C.c_int64
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def robot_info(robot: str):
    L = lib()
    d, nf, nl, res = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lo = np.zeros(16, np.float32)
    hi = np.zeros(16, np.float32)
    assert L.or_robot_info(ROBOT_ID[robot], d, nf, nl, res, lo, hi) == 0
    return dict(dim=d.value, n_fine=nf.value, n_links=nl.value, resolution=res.value,
                lo=lo[:d.value].copy(), hi=hi[:d.value].copy())


def bounds(robot: str):
    i = robot_info(robot)
    return i["lo"], i["hi"]


def trig(x):
    x = _f32(x).reshape(-1)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    lib().or_trig(x, x.size, s, c)
    return s, c


def sphere_fk(robot: str, q) -> np.ndarray:
    q = _f32(q)
    nf = robot_info(robot)["n_fine"]
    out = np.zeros((q.shape[0], nf, 4), np.float32)
    assert lib().or_sphere_fk(ROBOT_ID[robot], q, q.shape[0], out) == 0
    return out


def eefk(robot: str, q) -> np.ndarray:
    """[M][3][4] end-effector pose (Robot::eefk) from the restatement's frames."""
    q = _f32(q)
    out = np.zeros((q.shape[0], 3, 4), np.float32)
    assert lib().or_eefk(ROBOT_ID[robot], q, q.shape[0], out.reshape(-1)) == 0
    return out


class PortEnvironment:
    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.or_env_create())

    def __del__(self):
        try:
            self._L.or_env_destroy(self._h)
        except Exception:
            pass

    def add_spheres(self, xyzr):
        for s in _f32(xyzr).reshape(-1, 4):
            self._L.or_env_add_sphere(self._h, *[float(v) for v in s])

    def set_attachment(self, tf, xyzr):
        """collision/attachments.hh:13-56 (see oracle_validate.cpp::pose_attachment); tf = 3x4 [R | t] relative to the end effector."""
        xyzr = _f32(xyzr).reshape(-1, 4)
        assert self._L.or_env_set_attachment(self._h, _f32(np.asarray(tf, np.float32).reshape(-1)[:12].copy()), xyzr, xyzr.shape[0]) == 0

    def clear_attachment(self):
        self._L.or_env_set_attachment(self._h, np.zeros(12, np.float32), np.zeros((0, 4), np.float32), 0)

    def add_cuboids(self, c15, z_aligned=False):
        for c in _f32(c15).reshape(-1, 15):
            self._L.or_env_add_cuboid(self._h, np.ascontiguousarray(c), int(z_aligned))

    def add_capsules(self, c7, z_aligned=False):
        for c in _f32(c7).reshape(-1, 7):
            self._L.or_env_add_capsule(self._h, np.ascontiguousarray(c), int(z_aligned))

    def add_heightfield(self, centre, scale, heights):
        """factory::heightfield::flat(centre, scale, heights.shape[1] x heights.shape[0] cells, row-major heights)"""
        h = _f32(heights)
        self._L.or_env_add_heightfield(self._h, *[float(v) for v in centre], *[float(v) for v in scale], h.shape[1], h.shape[0], h.reshape(-1))

    def add_capt(self, t: dict):
        aff = np.ascontiguousarray(t["aff"], np.float32)
        self._L.or_env_add_capt(self._h, int(t["nlog2"]), _f32(t["tests"]), _f32(t["aabbs"]),
                                np.ascontiguousarray(t["aff_starts"], np.uint32), aff, aff.shape[1] // 8,
                                _f32(t["top"]), t["r_min"], t["r_max"], t["r_point"])

    def sort(self):
        self._L.or_env_sort(self._h)

    def finalize(self):
        pass

    def min_distances(self, kind: int) -> np.ndarray:
        n = self._L.or_env_min_distances(self._h, kind, None)
        out = np.zeros(n, dtype=np.float32)
        if n:
            self._L.or_env_min_distances(self._h, kind, out.ctypes.data_as(C.c_void_p))
        return out

    def validate_motions(self, robot, q0, q1, resolution=0, nthreads=1, want_flops=False):
        q0 = _f32(q0)
        q1 = _f32(q1)
        out = np.zeros(q0.shape[0], np.uint8)
        fl = np.zeros(q0.shape[0], np.float64) if want_flops else None
        rc = self._L.or_validate_motions(self._h, ROBOT_ID[robot], q0, q1, q0.shape[0], resolution, out,
                                         fl.ctypes.data_as(C.c_void_p) if want_flops else None, nthreads)
        assert rc == 0
        return (out, fl) if want_flops else out

    def validate_states(self, robot, q, nthreads=1):
        return self.validate_motions(robot, q, q, resolution=1, nthreads=nthreads)

    def edge_min_clearance(self, robot, q0, q1, resolution=0, nthreads=1):
        q0 = _f32(q0)
        q1 = _f32(q1)
        out = np.zeros(q0.shape[0], np.float64)
        assert self._L.or_edge_min_clearance(self._h, ROBOT_ID[robot], q0, q1, q0.shape[0], resolution, out, nthreads) == 0
        return out


def knn(db, q, k, metric=0, param=0.0, exclude_self=False, nthreads=1):
    db = np.ascontiguousarray(db, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    idx = np.zeros((q.shape[0], k), np.int64)
    dist = np.zeros((q.shape[0], k), np.float64)
    assert lib().or_knn(db, db.shape[0], q, q.shape[0], db.shape[1], k, metric, param, int(exclude_self), idx, dist, nthreads) == 0
    return idx, dist


def knn_radius(db, q, radius, metric=0, param=0.0):
    db = np.ascontiguousarray(db, np.float64)
    q = np.ascontiguousarray(q, np.float64).reshape(-1)
    idx = np.zeros(db.shape[0], np.int64)
    dist = np.zeros(db.shape[0], np.float64)
    n = lib().or_knn_radius(db, db.shape[0], q, db.shape[1], radius, metric, param, idx, dist, db.shape[0])
    return idx[:n], dist[:n]


# ---------------------------------------------------------------------------------------------- roadmap sampler
def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; the Random123 reference algorithm). ctr [n][4], key [2] uint32 -> [n][4]."""
    c = ctr.astype(np.uint64).copy()
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c[:, 0]
        p1 = m1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def prm_candidates(robot: str, seed: int, first: int, count: int) -> np.ndarray:
    """Candidate states first .. first+count-1 of the roadmap sample stream (ompl_b200/csrc/roadmap.cu prm_sample_kernel):
    q_j = fma(u_j, hi_j - lo_j, lo_j), u_j = (Philox word j >> 8) * 2^-24, counter = (c_lo, c_hi, j // 4, 0), key = seed."""
    lo, hi = bounds(robot)
    dim = len(lo)
    c = np.arange(first, first + count, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    words = []
    for sub in range((dim + 3) // 4):
        ctr = np.stack([c & np.uint64(0xFFFFFFFF), c >> np.uint64(32), np.full(count, sub, np.uint64), np.zeros(count, np.uint64)], axis=1)
        words.append(philox4x32_10(ctr, key))
    w = np.concatenate(words, axis=1)[:, :dim]
    u = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    span = (hi - lo).astype(np.float32)
    # single-rounding fma: exact product and sum in float64 (24-bit x 24-bit products are exact), rounded once to float32
    return (u.astype(np.float64) * span.astype(np.float64) + lo.astype(np.float64)).astype(np.float32)


def prm_edge_list(nbr: np.ndarray, dist: np.ndarray, q_first: int = 0, q_count: int | None = None):
    """Edge list rule of the batched builder: for query i (in order) and slot s, j = nbr[i][s] is kept when i is not in nbr[j],
    else when i is the pair's keeper (lower index if i + j even, higher if odd); stored (min, max).
    Returns (edges int32 [E][2], weights [E])."""
    n, k = nbr.shape
    q_count = n - q_first if q_count is None else q_count
    e, w = [], []
    for i in range(q_first, q_first + q_count):
        for s in range(k):
            j = int(nbr[i, s])
            if j < 0 or j == i:
                continue
            mine = (i < j) if (i + j) % 2 == 0 else (i > j)
            if not mine and i in nbr[j]:
                continue
            e.append((min(i, j), max(i, j)))
            w.append(dist[i - q_first, s])
    return np.array(e, np.int32).reshape(-1, 2), np.array(w, np.float64)


def filter_pointcloud(xyz, min_dist, max_range, origin, ws_min, ws_max, cull=True) -> np.ndarray:
    """Restatement of vamp::collision::filter_pointcloud (oracle_filter.cpp; collision/filter.hh:176-274)."""
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty((len(xyz) + 1, 3), np.float32)
    n = lib().orf_filter_pointcloud(xyz, len(xyz), float(min_dist), float(max_range), _f32(origin), _f32(ws_min), _f32(ws_max),
                                    int(bool(cull)), out, len(out))
    return out[:n].copy()
