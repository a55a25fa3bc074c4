"""ctypes binding of the b200mp C ABI (include/b200mp.h). No compute happens in Python and there is no
CPU fallback: if the CUDA library is missing or no GPU is present, construction raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROBOTS = {"panda": 0, "fetch": 1, "ur5": 2, "baxter": 3}
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_vp = C.c_void_p
_i64 = C.c_int64


class RoadmapStats(C.Structure):
    """b200mp_roadmap_stats (include/b200mp.h)."""
    _fields_ = [("n_states", _i64), ("n_candidates", _i64), ("n_edges", _i64), ("n_valid_edges", _i64),
                ("ms_sample", C.c_double), ("ms_knn", C.c_double), ("ms_edges", C.c_double), ("ms_validate", C.c_double),
                ("ms_total", C.c_double), ("ms_gather", C.c_double), ("n_local_edges", _i64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class B200MPError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200mp error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    # B200MP_LIB: a differently built libb200mp.so (tools/build_variant.sh; development only)
    return os.environ.get("B200MP_LIB") or os.path.join(_HERE, "libb200mp.so")


def load_library():
    """Load libb200mp.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise B200MPError(-1, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    L = C.CDLL(p)
    L.b200mp_last_error.restype = C.c_char_p
    L.b200mp_create.argtypes = [C.c_int, C.POINTER(_vp)]
    L.b200mp_destroy.argtypes = [_vp]
    L.b200mp_device.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.b200mp_synchronize.argtypes = [_vp]
    L.b200mp_robot_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _vp, _vp]
    L.b200mp_env_create.argtypes = [_vp, C.POINTER(_vp)]
    L.b200mp_env_destroy.argtypes = [_vp]
    L.b200mp_env_add_spheres.argtypes = [_vp, _vp, C.c_int]
    L.b200mp_env_add_cuboids.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.b200mp_env_add_capsules.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.b200mp_env_add_capt.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _i64, _vp, C.c_float, C.c_float, C.c_float]
    L.b200mp_env_add_heightfield.argtypes = [_vp] + [C.c_float] * 6 + [_i64, _i64, _vp]
    L.b200mp_env_add_pointcloud.argtypes = [_vp, _vp, _i64, C.c_float, C.c_float, C.c_float]
    L.b200mp_env_add_pointcloud_device.argtypes = [_vp, _vp, _i64, C.c_float, C.c_float, C.c_float]
    L.b200mp_capt_build.argtypes = [_vp, _i64, C.c_float, C.c_float, C.c_float, C.POINTER(_vp)]
    L.b200mp_capt_destroy.argtypes = [_vp]
    L.b200mp_filter_pointcloud.argtypes = [_vp, _vp, _i64, C.c_float, C.c_float, _vp, _vp, _vp, C.c_int, _vp, _vp]
    L.b200mp_capt_sizes.argtypes = [_vp, _vp]
    L.b200mp_capt_export.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
    L.b200mp_env_sort.argtypes = [_vp]
    L.b200mp_env_set_attachment.argtypes = [_vp, C.c_void_p, C.c_void_p, C.c_int]
    L.b200mp_session_open.argtypes = [_vp, C.c_int, _vp, C.POINTER(_vp)]
    L.b200mp_session_close.argtypes = [_vp]
    L.b200mp_session_close.restype = None
    L.b200mp_session_check_motion.argtypes = [_vp, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.b200mp_session_check_motion_f64.argtypes = [_vp, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.b200mp_session_is_valid.argtypes = [_vp, C.c_void_p, C.c_void_p]
    L.b200mp_session_is_valid_f64.argtypes = [_vp, C.c_void_p, C.c_void_p]
    L.b200mp_session_stats.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.b200mp_env_set_grid.argtypes = [_vp, C.c_int]
    L.b200mp_env_commit.argtypes = [_vp]
    L.b200mp_env_min_distances.argtypes = [_vp, C.c_int, _vp]
    L.b200mp_env_min_distances.restype = _i64
    L.b200mp_validate_states.argtypes = [_vp, C.c_int, _vp, _vp, _i64, _vp]
    L.b200mp_validate_motions.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _i64, C.c_int, _vp]
    L.b200mp_validate_states_f64.argtypes = [_vp, C.c_int, _vp, _vp, _i64, _vp]
    L.b200mp_validate_motions_f64.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _i64, C.c_int, _vp]
    L.b200mp_validate_states_device.argtypes = [_vp, C.c_int, _vp, _vp, _i64, _vp, _vp]
    L.b200mp_validate_motions_device.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _i64, C.c_int, _vp, _vp]
    L.b200mp_validate_edges_device.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _i64, C.c_int, _vp, _vp]
    L.b200mp_prm_sample_device.argtypes = [_vp, C.c_int, _vp, C.c_uint64, C.c_uint64, _i64, _vp, _i64, C.POINTER(_i64), _vp]
    L.b200mp_prm_edges_device.argtypes = [_vp, _vp, _vp, _i64, C.c_int, _i64, _i64, _vp, _vp, _i64, C.POINTER(_i64), _vp]
    L.b200mp_prm_build.argtypes = [_vp, C.c_int, _vp, C.c_uint64, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _i64,
                                   C.POINTER(RoadmapStats)]
    L.b200mp_comm_unique_id.argtypes = [_vp, C.c_size_t]
    L.b200mp_comm_create.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.b200mp_comm_create_all.argtypes = [_vp, C.c_int, _vp]
    L.b200mp_comm_destroy.argtypes = [_vp]
    L.b200mp_comm_rank.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.b200mp_prm_build_multi.argtypes = [_vp, C.c_int, _vp, C.c_uint64, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _i64,
                                         C.POINTER(RoadmapStats)]
    L.b200mp_env_clone.argtypes = [_vp, _vp, C.POINTER(_vp)]
    L.b200mp_knn_query_part_ids_device_f32.argtypes = [_vp, _vp, _i64, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp]
    L.b200mp_knn_query_device_f32.argtypes = [_vp, _vp, _i64, C.c_int, _i64, C.c_int, _vp, _vp, _vp, _vp]
    L.b200mp_knn_query_part_device_f32.argtypes = [_vp, _vp, _i64, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
    L.b200mp_chain_horn_environment.argtypes = [C.c_int, C.c_double, _vp, C.POINTER(C.c_int)]
    L.b200mp_chain_create.argtypes = [_vp, C.c_int, C.c_double, _vp, C.c_int, C.c_double, C.POINTER(_vp)]
    L.b200mp_chain_destroy.argtypes = [_vp]
    L.b200mp_chain_longest_valid_segment.argtypes = [_vp]
    L.b200mp_chain_longest_valid_segment.restype = C.c_double
    L.b200mp_chain_distance.argtypes = [_vp, _vp, _vp, _i64, _vp]
    L.b200mp_chain_interpolate.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp]
    L.b200mp_chain_validate_states.argtypes = [_vp, _vp, _i64, _vp]
    L.b200mp_chain_check_motions.argtypes = [_vp, _vp, _vp, _i64, _vp, _vp, _vp]
    L.b200mp_chain_validate_states_device.argtypes = [_vp, _vp, _i64, _vp, _vp]
    L.b200mp_chain_check_motions_device.argtypes = [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]
    L.b200mp_set_option.argtypes = [_vp, C.c_int, C.c_int]
    L.b200mp_launch_count.argtypes = [_vp]
    L.b200mp_launch_count.restype = _i64
    L.b200mp_fp32_peak.argtypes = [_vp, C.c_int, C.POINTER(C.c_double)]
    L.b200mp_knn_create.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.b200mp_knn_create_chain.argtypes = [_vp, C.c_int, C.c_double, C.POINTER(_vp)]
    L.b200mp_knn_destroy.argtypes = [_vp]
    L.b200mp_knn_clear.argtypes = [_vp]
    L.b200mp_knn_set_index.argtypes = [_vp, C.c_int]
    L.b200mp_knn_rows_visited.argtypes = [_vp]
    L.b200mp_knn_rows_visited.restype = _i64
    L.b200mp_knn_search_stats.argtypes = [_vp, _vp]
    L.b200mp_knn_size.argtypes = [_vp]
    L.b200mp_knn_size.restype = _i64
    L.b200mp_knn_add.argtypes = [_vp, _vp, _i64, C.POINTER(_i64)]
    L.b200mp_knn_add_device_f32.argtypes = [_vp, _vp, _i64, C.c_float, _vp]
    L.b200mp_knn_remove.argtypes = [_vp, _i64]
    L.b200mp_knn_query.argtypes = [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp, _vp]
    L.b200mp_knn_query_device.argtypes = [_vp, _vp, _i64, C.c_int, _vp, _i64, C.c_int, _vp, _vp, _vp, _vp]
    L.b200mp_knn_radius.argtypes = [_vp, _vp, C.c_double, _vp, _vp, _i64, C.POINTER(_i64)]
    _LIB = L
    return L


def _check(rc: int):
    if rc != 0:
        raise B200MPError(rc, load_library().b200mp_last_error().decode())


def robot_info(robot: str) -> dict:
    L = load_library()
    d, ns, res = C.c_int(), C.c_int(), C.c_int()
    lo = np.zeros(16, np.float32)
    hi = np.zeros(16, np.float32)
    _check(L.b200mp_robot_info(ROBOTS[robot], d, ns, res, lo.ctypes.data, hi.ctypes.data))
    return dict(dim=d.value, n_spheres=ns.value, resolution=res.value, lo=lo[:d.value].copy(), hi=hi[:d.value].copy())


def capt_build(xyz, r_min: float, r_max: float, r_point: float) -> dict:
    """Host build of the point-cloud tree (b200mp_capt_build); returns the flat arrays in the layout of
    b200mp_env_add_capt. Needs no GPU."""
    L = load_library()
    pts = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    h = _vp()
    _check(L.b200mp_capt_build(pts.ctypes.data, pts.shape[0], r_min, r_max, r_point, C.byref(h)))
    try:
        sizes = np.zeros(4, np.int64)
        _check(L.b200mp_capt_sizes(h, sizes.ctypes.data))
        nlog2, nt, nl, na = (int(v) for v in sizes)
        tests = np.zeros(nt, np.float32)
        aabbs = np.zeros((nl, 6), np.float32)
        starts = np.zeros(nl + 1, np.uint32)
        aff = np.zeros((3, na * 8), np.float32)
        top = np.zeros(6, np.float32)
        _check(L.b200mp_capt_export(h, tests.ctypes.data, aabbs.ctypes.data, starts.ctypes.data, aff.ctypes.data, top.ctypes.data))
    finally:
        L.b200mp_capt_destroy(h)
    return dict(nlog2=nlog2, tests=tests, aabbs=aabbs, aff_starts=starts, aff=aff, top=top, r_min=r_min, r_max=r_max,
                r_point=r_point)


def _host(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """One engine context per GPU (b200mp_create)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = _vp()
        _check(self._L.b200mp_create(device, C.byref(h)))
        self._h = h
        dev, sms = C.c_int(), C.c_int()
        _check(self._L.b200mp_device(self._h, dev, sms))
        self.device, self.sm_count = dev.value, sms.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200mp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self._L.b200mp_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.b200mp_launch_count(self._h))

    def set_option(self, option: int, value: int):
        """b200mp_set_option; option 1 = B200MP_OPT_FIRST_LOOK."""
        _check(self._L.b200mp_set_option(self._h, option, value))

    def fp32_peak_tflops(self, iters: int = 4096) -> float:
        out = C.c_double()
        _check(self._L.b200mp_fp32_peak(self._h, iters, C.byref(out)))
        return out.value

    # ---- host-buffer entry points (what the OMPL adapters call) ----
    def filter_pointcloud(self, xyz, min_dist: float, max_range: float, origin, ws_min, ws_max, cull: bool = True) -> np.ndarray:
        """vamp.filter_pointcloud on the device (b200mp_filter_pointcloud); returns the kept points [m][3] in the reference's order."""
        pts = _host(np.asarray(xyz).reshape(-1, 3), np.float32)
        o, lo, hi = (_host(np.asarray(v).reshape(3), np.float32) for v in (origin, ws_min, ws_max))
        out = np.empty((pts.shape[0] + 1, 3), np.float32)
        m = C.c_int64(0)
        _check(self._L.b200mp_filter_pointcloud(self._h, pts.ctypes.data, pts.shape[0], float(min_dist), float(max_range), o.ctypes.data,
                                                lo.ctypes.data, hi.ctypes.data, int(bool(cull)), out.ctypes.data, C.byref(m)))
        return out[:m.value].copy()

    def validate_states(self, robot: str, env: "Environment", q) -> np.ndarray:
        q = np.asarray(q)
        dim = robot_info(robot)["dim"]
        f64 = q.dtype == np.float64
        q = _host(q, np.float64 if f64 else np.float32).reshape(-1, dim)
        out = np.empty(q.shape[0], np.uint8)
        fn = self._L.b200mp_validate_states_f64 if f64 else self._L.b200mp_validate_states
        _check(fn(self._h, ROBOTS[robot], env._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def validate_motions(self, robot: str, env: "Environment", q0, q1, resolution: int = 0, out=None) -> np.ndarray:
        q0 = np.asarray(q0)
        dim = robot_info(robot)["dim"]
        f64 = q0.dtype == np.float64
        dt = np.float64 if f64 else np.float32
        q0 = _host(q0, dt).reshape(-1, dim)
        q1 = _host(q1, dt).reshape(-1, dim)
        if q0.shape != q1.shape:
            raise ValueError("q0 and q1 must have the same shape")
        if out is None:
            out = np.empty(q0.shape[0], np.uint8)
        fn = self._L.b200mp_validate_motions_f64 if f64 else self._L.b200mp_validate_motions
        _check(fn(self._h, ROBOTS[robot], env._h, q0.ctypes.data, q1.ctypes.data, q0.shape[0], resolution, out.ctypes.data))
        return out

    # ---- raw-pointer entry points (device or pinned host memory owned by the caller, e.g. torch tensors) ----
    def validate_motions_ptr(self, robot: str, env: "Environment", q0_ptr: int, q1_ptr: int, m: int, valid_ptr: int,
                             resolution: int = 0, device: bool = True, stream: int = 0):
        if device:
            _check(self._L.b200mp_validate_motions_device(self._h, ROBOTS[robot], env._h, q0_ptr, q1_ptr, m, resolution,
                                                          valid_ptr, stream or None))
        else:
            _check(self._L.b200mp_validate_motions(self._h, ROBOTS[robot], env._h, q0_ptr, q1_ptr, m, resolution, valid_ptr))

    def validate_states_ptr(self, robot: str, env: "Environment", q_ptr: int, m: int, valid_ptr: int, device: bool = True,
                            stream: int = 0):
        if device:
            _check(self._L.b200mp_validate_states_device(self._h, ROBOTS[robot], env._h, q_ptr, m, valid_ptr, stream or None))
        else:
            _check(self._L.b200mp_validate_states(self._h, ROBOTS[robot], env._h, q_ptr, m, valid_ptr))


    def validate_edges_ptr(self, robot: str, env: "Environment", states_ptr: int, edges_ptr: int, m: int, valid_ptr: int,
                           resolution: int = 0, stream: int = 0):
        """Motions given as int32 index pairs into one device-resident state table (b200mp_validate_edges_device)."""
        _check(self._L.b200mp_validate_edges_device(self._h, ROBOTS[robot], env._h, states_ptr, edges_ptr, m, resolution, valid_ptr,
                                                    stream or None))

    # ---- roadmap phases (device pointers; see ompl_b200/roadmap.py for the multi-GPU composition) ----
    def prm_sample_ptr(self, robot: str, env: "Environment", seed: int, cand_first: int, cand_count: int, states_ptr: int,
                       capacity: int, stream: int = 0) -> int:
        n = _i64()
        _check(self._L.b200mp_prm_sample_device(self._h, ROBOTS[robot], env._h, seed, cand_first, cand_count, states_ptr, capacity,
                                                C.byref(n), stream or None))
        return int(n.value)

    def prm_edges_ptr(self, nbr_ptr: int, dist_local_ptr: int, n: int, k: int, q_first: int, q_count: int, edges_ptr: int,
                      weights_ptr: int, capacity: int, stream: int = 0) -> int:
        ne = _i64()
        _check(self._L.b200mp_prm_edges_device(self._h, nbr_ptr, dist_local_ptr or None, n, k, q_first, q_count, edges_ptr,
                                               weights_ptr or None, capacity, C.byref(ne), stream or None))
        return int(ne.value)

    def prm_build(self, robot: str, env: "Environment", n: int, k: int, seed: int, resolution: int = 0, want_weights: bool = True):
        """Single-GPU roadmap (b200mp_prm_build) -> dict(states f32[n][dim], edges i32[E][2], valid u8[E], weights f64[E], stats)."""
        dim = robot_info(robot)["dim"]
        cap = n * k
        states = np.empty((n, dim), np.float32)
        edges = np.empty((cap, 2), np.int32)
        valid = np.empty(cap, np.uint8)
        weights = np.empty(cap, np.float64) if want_weights else None
        st = RoadmapStats()
        _check(self._L.b200mp_prm_build(self._h, ROBOTS[robot], env._h, seed, n, k, resolution, states.ctypes.data, edges.ctypes.data,
                                        valid.ctypes.data, weights.ctypes.data if want_weights else None, cap, C.byref(st)))
        e = int(st.n_edges)
        return dict(states=states, edges=edges[:e], valid=valid[:e], weights=weights[:e] if want_weights else None, stats=st.as_dict())


UNIQUE_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """b200mp_comm_unique_id: made by rank 0, handed to the other ranks by any means (torch.distributed, MPI, a file)."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(load_library().b200mp_comm_unique_id(buf, UNIQUE_ID_BYTES))
    return buf.raw


class Comm:
    """One rank of a group of GPUs (b200mp_comm): the context of this process's GPU + an NCCL communicator."""

    def __init__(self, ctx: Context, unique_id: bytes | None, rank: int, world: int):
        self._L = load_library()
        self._ctx = ctx
        h = _vp()
        idbuf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES) if unique_id is not None else None
        _check(self._L.b200mp_comm_create(ctx._h, idbuf, rank, world, C.byref(h)))
        self._h = h
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._L.b200mp_comm_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prm_build(self, robot: str, env: "Environment", n: int, k: int, seed: int, resolution: int = 0, want_arrays: bool = True,
                  want_weights: bool = True, want_neighbors: bool = False, incremental: bool = False):
        """Collective roadmap build (b200mp_prm_build_multi); every rank receives the whole roadmap when want_arrays.
        incremental: B200MP_PRM_INCREMENTAL -- state i is connected to its k nearest among states 0 .. i - 1 only, stock PRM's graph."""
        dim = robot_info(robot)["dim"]
        cap = n * k
        st = RoadmapStats()
        flags = 1 if incremental else 0
        if not want_arrays:
            _check(self._L.b200mp_prm_build_multi(self._h, ROBOTS[robot], env._h, seed, n, k, resolution, flags, None, None, None, None, None, cap,
                                                  C.byref(st)))
            return dict(stats=st.as_dict())
        states = np.empty((n, dim), np.float32)
        edges = np.empty((cap, 2), np.int32)
        valid = np.empty(cap, np.uint8)
        weights = np.empty(cap, np.float64) if want_weights else None
        nbr = np.empty((n, min(k, n - 1)), np.int32) if want_neighbors and n > 1 else None
        _check(self._L.b200mp_prm_build_multi(self._h, ROBOTS[robot], env._h, seed, n, k, resolution, flags, states.ctypes.data, edges.ctypes.data,
                                              valid.ctypes.data, weights.ctypes.data if want_weights else None,
                                              nbr.ctypes.data if nbr is not None else None, cap, C.byref(st)))
        e = int(st.n_edges)
        # (views of the capacity-sized buffers, not copies: ~180 MB at 2^20 x 16)
        return dict(states=states, edges=edges[:e], valid=valid[:e],
                    weights=weights[:e] if want_weights else None, neighbors=nbr, stats=st.as_dict())


class Session:
    """The resident single-call kernel (b200mp_session_*): one checkMotion / isValid per call without a kernel launch."""

    def __init__(self, ctx: "Context", robot: str, env: "Environment"):
        self._L = load_library()
        self._ctx, self._env = ctx, env
        self._dim = robot_info(robot)["dim"]
        h = _vp()
        _check(self._L.b200mp_session_open(ctx._h, ROBOTS[robot], env._h, C.byref(h)))
        self._h = h
        self._out = np.zeros(1, np.uint8)

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200mp_session_close(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check_motion(self, q0, q1, resolution: int = 0) -> bool:
        a, b = np.ascontiguousarray(q0), np.ascontiguousarray(q1)
        assert a.shape == (self._dim,) and b.shape == (self._dim,) and a.dtype == b.dtype
        fn = self._L.b200mp_session_check_motion_f64 if a.dtype == np.float64 else self._L.b200mp_session_check_motion
        if a.dtype not in (np.float32, np.float64):
            a, b = a.astype(np.float32), b.astype(np.float32)
        _check(fn(self._h, a.ctypes.data, b.ctypes.data, int(resolution), self._out.ctypes.data))
        return bool(self._out[0])

    def is_valid(self, q) -> bool:
        a = np.ascontiguousarray(q)
        fn = self._L.b200mp_session_is_valid_f64 if a.dtype == np.float64 else self._L.b200mp_session_is_valid
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        _check(fn(self._h, a.ctypes.data, self._out.ctypes.data))
        return bool(self._out[0])

    def stats(self) -> dict:
        c, l = C.c_int64(0), C.c_int64(0)
        _check(self._L.b200mp_session_stats(self._h, C.byref(c), C.byref(l)))
        return {"calls": c.value, "launches": l.value}


class Environment:
    """Mirror of vamp::collision::Environment<float> (collision/environment.hh:14-85): shape lists per class,
    `sort()` by min_distance, then `commit()` uploads the scene to the GPU."""

    def __init__(self, ctx: Context):
        self._L = load_library()
        self._ctx = ctx
        h = _vp()
        _check(self._L.b200mp_env_create(ctx._h, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._L.b200mp_env_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_spheres(self, xyzr):
        a = _host(xyzr, np.float32).reshape(-1, 4)
        _check(self._L.b200mp_env_add_spheres(self._h, a.ctypes.data, a.shape[0]))

    def add_cuboids(self, c15, z_aligned: bool = False):
        a = _host(c15, np.float32).reshape(-1, 15)
        _check(self._L.b200mp_env_add_cuboids(self._h, a.ctypes.data, a.shape[0], int(z_aligned)))

    def add_capsules(self, c7, z_aligned: bool = False):
        a = _host(c7, np.float32).reshape(-1, 7)
        _check(self._L.b200mp_env_add_capsules(self._h, a.ctypes.data, a.shape[0], int(z_aligned)))

    def add_heightfield(self, centre, scale, heights):
        """factory::heightfield::flat: heights[ny][nx] row-major, cell size `scale`, centred at `centre`."""
        h = _host(np.asarray(heights), np.float32)
        _check(self._L.b200mp_env_add_heightfield(self._h, *[float(v) for v in centre], *[float(v) for v in scale], h.shape[1], h.shape[0],
                                                  h.ctypes.data))

    def add_capt(self, t: dict):
        tests = _host(t["tests"], np.float32)
        aabbs = _host(t["aabbs"], np.float32)
        starts = _host(t["aff_starts"], np.uint32)
        aff = _host(t["aff"], np.float32)
        top = _host(t["top"], np.float32)
        _check(self._L.b200mp_env_add_capt(self._h, int(t["nlog2"]), tests.ctypes.data, aabbs.ctypes.data, starts.ctypes.data,
                                           aff.ctypes.data, aff.shape[1] // 8, top.ctypes.data, t["r_min"], t["r_max"],
                                           t["r_point"]))

    def add_pointcloud(self, xyz, r_min: float, r_max: float, r_point: float):
        """vamp::collision::CAPT(points, r_min, r_max, r_point), built on the host by the engine."""
        a = _host(xyz, np.float32).reshape(-1, 3)
        _check(self._L.b200mp_env_add_pointcloud(self._h, a.ctypes.data, a.shape[0], r_min, r_max, r_point))

    def add_pointcloud_device(self, xyz, r_min: float, r_max: float, r_point: float):
        """The same tree structure built on the GPU (b200mp_env_add_pointcloud_device; at most 2^17 points)."""
        a = _host(xyz, np.float32).reshape(-1, 3)
        _check(self._L.b200mp_env_add_pointcloud_device(self._h, a.ctypes.data, a.shape[0], r_min, r_max, r_point))

    def set_attachment(self, tf, xyzr):
        """vamp::collision::Attachment (collision/attachments.hh:13-56): spheres xyzr[n][4] rigidly fixed to the end effector at the
        pose tf (3x4 [R | t], relative to Robot::eefk). Validation then runs Robot::fkcc_attach. No re-commit needed."""
        t = _host(np.asarray(tf, np.float32).reshape(-1)[:12].copy(), np.float32)
        a = _host(xyzr, np.float32).reshape(-1, 4)
        _check(self._L.b200mp_env_set_attachment(self._h, t.ctypes.data, a.ctypes.data, a.shape[0]))

    def clear_attachment(self):
        _check(self._L.b200mp_env_set_attachment(self._h, None, None, 0))

    def set_grid(self, enable: bool):
        """Candidate-grid culling of sorted scenes on/off (b200mp_env_set_grid); takes effect at the next commit()."""
        _check(self._L.b200mp_env_set_grid(self._h, int(enable)))

    def sort(self):
        _check(self._L.b200mp_env_sort(self._h))

    def commit(self):
        _check(self._L.b200mp_env_commit(self._h))

    finalize = commit

    def min_distances(self, kind: int) -> np.ndarray:
        n = self._L.b200mp_env_min_distances(self._h, kind, None)
        out = np.zeros(max(n, 0), np.float32)
        if n > 0:
            self._L.b200mp_env_min_distances(self._h, kind, out.ctypes.data)
        return out


class NearestNeighbors:
    """Mirror of ompl::NearestNeighbors<_T> (datastructures/NearestNeighbors.h:46-115) over point ids: add / remove /
    nearest / nearestK / nearestR / size / clear, results sorted ascending, double distances."""

    def __init__(self, ctx: Context, dim: int, metric: int = 0, link_length: float = 0.0):
        """metric 0: RealVectorStateSpace::distance; metric 1: KinematicChainSpace::distance with `link_length`."""
        self._L = load_library()
        self._ctx = ctx
        self.dim = dim
        h = _vp()
        if metric == 1:
            _check(self._L.b200mp_knn_create_chain(ctx._h, dim, link_length, C.byref(h)))
        else:
            _check(self._L.b200mp_knn_create(ctx._h, dim, metric, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._L.b200mp_knn_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        _check(self._L.b200mp_knn_clear(self._h))

    def rows_visited(self) -> int:
        return int(self._L.b200mp_knn_rows_visited(self._h))

    def search_stats(self) -> dict:
        """Counters of the cell-index search (collected only while B200MP_KNN_STAT_PAIRS is set)."""
        out = np.zeros(4, np.int64)
        _check(self._L.b200mp_knn_search_stats(self._h, out.ctypes.data))
        return dict(pairs=int(out[0]), exact=int(out[1]), insertions=int(out[2]), leaf_visits=int(out[3]))

    def set_index(self, enable: bool):
        """Cell-index pruning for large batches on/off (b200mp_knn_set_index)."""
        _check(self._L.b200mp_knn_set_index(self._h, int(enable)))

    def size(self) -> int:
        return int(self._L.b200mp_knn_size(self._h))

    def add(self, pts) -> int:
        a = _host(pts, np.float64).reshape(-1, self.dim)
        first = _i64()
        _check(self._L.b200mp_knn_add(self._h, a.ctypes.data, a.shape[0], C.byref(first)))
        return int(first.value)

    def add_device_f32(self, ptr: int, n: int, max_norm: float, stream: int = 0):
        _check(self._L.b200mp_knn_add_device_f32(self._h, ptr, n, max_norm, stream or None))

    def remove(self, idx: int):
        _check(self._L.b200mp_knn_remove(self._h, idx))

    def nearest_k(self, q, k: int, self_ids=None):
        """-> (idx[nq][k] int64 (-1 padded), dist[nq][k] float64 (+inf padded), counts[nq])"""
        q = _host(q, np.float64).reshape(-1, self.dim)
        nq = q.shape[0]
        idx = np.full((nq, k), -1, np.int64)
        dist = np.full((nq, k), np.inf, np.float64)
        cnt = np.zeros(nq, np.int32)
        sp = None
        if self_ids is not None:
            sid = _host(self_ids, np.int64)
            sp = sid.ctypes.data
        _check(self._L.b200mp_knn_query(self._h, q.ctypes.data, nq, k, sp, idx.ctypes.data, dist.ctypes.data, cnt.ctypes.data))
        return idx, dist, cnt

    def nearest(self, q) -> int:
        if self.size() == 0:
            raise B200MPError(6, "No elements found in nearest neighbors data structure")
        idx, _, _ = self.nearest_k(q, 1)
        return int(idx[0, 0])

    def nearest_r(self, q, radius: float):
        q = _host(q, np.float64).reshape(-1)
        cap = max(self.size(), 1)
        idx = np.zeros(cap, np.int64)
        dist = np.zeros(cap, np.float64)
        cnt = _i64()
        _check(self._L.b200mp_knn_radius(self._h, q.ctypes.data, radius, idx.ctypes.data, dist.ctypes.data, cap, C.byref(cnt)))
        return idx[:cnt.value].copy(), dist[:cnt.value].copy()

    def query_device_f32(self, q_ptr: int, nq: int, k: int, idx_ptr: int, dist_ptr: int, self_base: int = 0,
                         exclude_self: bool = False, counts_ptr: int = 0, stream: int = 0):
        _check(self._L.b200mp_knn_query_device_f32(self._h, q_ptr, nq, k, self_base, int(exclude_self), idx_ptr, dist_ptr,
                                                   counts_ptr or None, stream or None))

    def query_part_ids_device_f32(self, q_ptr: int, nq: int, k: int, ids_ptr: int, part: int, n_parts: int, self_base: int = 0,
                                  exclude_self: bool = False, stream: int = 0):
        _check(self._L.b200mp_knn_query_part_ids_device_f32(self._h, q_ptr, nq, k, self_base, int(exclude_self), part, n_parts, ids_ptr,
                                                            stream or None))

    def query_part_device_f32(self, q_ptr: int, nq: int, k: int, idx_ptr: int, dist_ptr: int, part: int, n_parts: int, self_base: int = 0,
                              exclude_self: bool = False, stream: int = 0):
        _check(self._L.b200mp_knn_query_part_device_f32(self._h, q_ptr, nq, k, self_base, int(exclude_self), part, n_parts, idx_ptr,
                                                        dist_ptr, stream or None))

    def query_device(self, q_ptr: int, nq: int, k: int, idx_ptr: int, dist_ptr: int, self_base: int = 0, exclude_self: bool = False,
                     self_ids_ptr: int = 0, counts_ptr: int = 0, stream: int = 0):
        _check(self._L.b200mp_knn_query_device(self._h, q_ptr, nq, k, self_ids_ptr or None, self_base, int(exclude_self), idx_ptr,
                                               dist_ptr, counts_ptr or None, stream or None))


def horn_environment(d: int, eps: float) -> np.ndarray:
    """createHornEnvironment(d, eps) (demos/KinematicChain.h:279-315) -> segments [2(d-1)][4]; host only."""
    L = load_library()
    n = C.c_int()
    _check(L.b200mp_chain_horn_environment(d, eps, None, C.byref(n)))
    out = np.zeros((n.value, 4), np.float64)
    _check(L.b200mp_chain_horn_environment(d, eps, out.ctypes.data, C.byref(n)))
    return out


class KinematicChain:
    """Planar n-link chain of demos/KinematicChain.h on the GPU (fp64): KinematicChainSpace::distance / interpolate,
    KinematicChainValidityChecker::isValid and DiscreteMotionValidator::checkMotion over it."""

    def __init__(self, ctx: Context, n_links: int, link_length: float, segments, lvs_fraction: float = 0.01):
        self._L = load_library()
        self._ctx = ctx
        self.n = n_links
        seg = _host(segments, np.float64).reshape(-1, 4)
        h = _vp()
        _check(self._L.b200mp_chain_create(ctx._h, n_links, link_length, seg.ctypes.data if len(seg) else None, len(seg), lvs_fraction,
                                           C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._L.b200mp_chain_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def longest_valid_segment(self) -> float:
        return float(self._L.b200mp_chain_longest_valid_segment(self._h))

    def distance(self, a, b) -> np.ndarray:
        a, b = _host(a, np.float64).reshape(-1, self.n), _host(b, np.float64).reshape(-1, self.n)
        out = np.zeros(len(a))
        _check(self._L.b200mp_chain_distance(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    def interpolate(self, a, b, t) -> np.ndarray:
        a, b = _host(a, np.float64).reshape(-1, self.n), _host(b, np.float64).reshape(-1, self.n)
        tt = _host(np.broadcast_to(t, (len(a),)), np.float64)
        out = np.zeros_like(a)
        _check(self._L.b200mp_chain_interpolate(self._h, a.ctypes.data, b.ctypes.data, tt.ctypes.data, len(a), out.ctypes.data))
        return out

    def validate_states(self, q) -> np.ndarray:
        q = _host(q, np.float64).reshape(-1, self.n)
        out = np.zeros(len(q), np.uint8)
        _check(self._L.b200mp_chain_validate_states(self._h, q.ctypes.data, len(q), out.ctypes.data))
        return out

    def check_motions(self, a, b):
        """-> (valid u8, last_valid_t f64 (-1 when valid), nd i32)"""
        a, b = _host(a, np.float64).reshape(-1, self.n), _host(b, np.float64).reshape(-1, self.n)
        out = np.zeros(len(a), np.uint8)
        last = np.zeros(len(a))
        nd = np.zeros(len(a), np.int32)
        _check(self._L.b200mp_chain_check_motions(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data, last.ctypes.data, nd.ctypes.data))
        return out, last, nd

    def validate_states_ptr(self, q_ptr: int, m: int, valid_ptr: int, stream: int = 0):
        _check(self._L.b200mp_chain_validate_states_device(self._h, q_ptr, m, valid_ptr, stream or None))

    def check_motions_ptr(self, a_ptr: int, b_ptr: int, m: int, valid_ptr: int, last_ptr: int = 0, nd_ptr: int = 0, stream: int = 0):
        _check(self._L.b200mp_chain_check_motions_device(self._h, a_ptr, b_ptr, m, valid_ptr, last_ptr or None, nd_ptr or None, stream or None))
