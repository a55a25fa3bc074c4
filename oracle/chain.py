"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the planar kinematic-chain checkers (BASELINE config 5):
`RefChain` = oracle/_ref/libref_chain.so, the reference's own demos/KinematicChain.h and DiscreteMotionValidator.cpp
compiled unmodified; `PortChain` = the restatement in oracle/oracle_chain.cpp (liboracle.so). Same surface for both.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libref_chain.so")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_libs = {}


def ref_available() -> bool:
    return os.path.exists(REF_PATH)


def _load(prefix):
    if prefix in _libs:
        return _libs[prefix]
    if prefix == "refc":
        L = C.CDLL(REF_PATH)
    else:
        from . import port
        L = port.lib()
    f = lambda n: getattr(L, f"{prefix}_{n}")  # noqa: E731
    f("horn").argtypes = [C.c_int, C.c_double, _f64p]
    f("create").restype = C.c_void_p
    f("create").argtypes = [C.c_int, C.c_double, _f64p, C.c_int, C.c_double]
    f("destroy").argtypes = [C.c_void_p]
    f("longest_valid_segment").restype = C.c_double
    f("longest_valid_segment").argtypes = [C.c_void_p]
    f("distance").argtypes = [C.c_void_p, _f64p, _f64p, C.c_int64, _f64p]
    f("interpolate").argtypes = [C.c_void_p, _f64p, _f64p, _f64p, C.c_int64, _f64p]
    f("valid_states").argtypes = [C.c_void_p, _f64p, C.c_int64, _u8p]
    f("check_motions").argtypes = [C.c_void_p, _f64p, _f64p, C.c_int64, C.c_int, _u8p, C.c_void_p, C.c_void_p]
    _libs[prefix] = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Chain:
    prefix = None

    def __init__(self, n_links: int, link_length: float, segments, lvs_fraction: float = 0.01):
        self._L = _load(self.prefix)
        self.n = n_links
        seg = _f64(segments).reshape(-1, 4)
        self._h = C.c_void_p(self._f("create")(n_links, link_length, seg if len(seg) else np.zeros((1, 4)), len(seg), lvs_fraction))

    def _f(self, name):
        return getattr(self._L, f"{self.prefix}_{name}")

    def __del__(self):
        try:
            self._f("destroy")(self._h)
        except Exception:
            pass

    @classmethod
    def horn(cls, d: int, eps: float) -> np.ndarray:
        out = np.zeros((2 * (d - 1), 4))
        n = getattr(_load(cls.prefix), f"{cls.prefix}_horn")(d, eps, out)
        assert n == len(out)
        return out

    @property
    def longest_valid_segment(self) -> float:
        return float(self._f("longest_valid_segment")(self._h))

    def distance(self, a, b):
        a, b = _f64(a).reshape(-1, self.n), _f64(b).reshape(-1, self.n)
        out = np.zeros(len(a))
        self._f("distance")(self._h, a, b, len(a), out)
        return out

    def interpolate(self, a, b, t):
        a, b = _f64(a).reshape(-1, self.n), _f64(b).reshape(-1, self.n)
        out = np.zeros_like(a)
        self._f("interpolate")(self._h, a, b, _f64(np.broadcast_to(t, (len(a),))), len(a), out)
        return out

    def validate_states(self, q):
        q = _f64(q).reshape(-1, self.n)
        out = np.zeros(len(q), np.uint8)
        self._f("valid_states")(self._h, q, len(q), out)
        return out

    def check_motions(self, a, b, form: int = 0):
        """form 0: checkMotion(s1, s2); form 1: checkMotion(s1, s2, lastValid). -> (valid u8, last_t f64 (-1 = untouched), nd i32)"""
        a, b = _f64(a).reshape(-1, self.n), _f64(b).reshape(-1, self.n)
        out = np.zeros(len(a), np.uint8)
        last = np.full(len(a), -1.0)
        nd = np.zeros(len(a), np.int32)
        self._f("check_motions")(self._h, a, b, len(a), form, out, last.ctypes.data_as(C.c_void_p), nd.ctypes.data_as(C.c_void_p))
        return out, last, nd


class RefChain(_Chain):
    prefix = "refc"


class PortChain(_Chain):
    prefix = "orc"
