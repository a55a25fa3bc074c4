"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_ref/libref_nn.so: the UNMODIFIED reference
ompl::NearestNeighborsLinear<unsigned long> (src/ompl/datastructures/NearestNeighborsLinear.h:65-142) driven through a
std::function distance exactly as OMPL's planners drive it (oracle/ref_nn.cpp). Only tests/, bench.py's checker /
cpu_baseline legs and tests/golden/make_golden.py may import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_nn.so")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.refnn_nearest_k_batch.argtypes = [_f64p, C.c_int64, C.c_int, _f64p, C.c_int64, C.c_int, _i64p, _f64p, C.c_int]
        L.refnn_nearest_k_batch.restype = None
        _lib = L
    return _lib


def nearest_k(db, queries, k: int, nthreads: int = 1, exclude_ids=None):
    """NearestNeighborsLinear::nearestK for every query row -> (idx[nq][k] int64, dist[nq][k] float64), ascending.
    exclude_ids[i] (optional): a database id query i must not return -- PRM queries a milestone BEFORE adding it
    (PRM.cpp:583,601); here the container already holds it, so k + 1 are requested and that id is dropped."""
    db = np.ascontiguousarray(db, np.float64)
    q = np.ascontiguousarray(queries, np.float64).reshape(-1, db.shape[1])
    kk = k + (1 if exclude_ids is not None else 0)
    idx = np.empty((len(q), kk), np.int64)
    dist = np.empty((len(q), kk), np.float64)
    lib().refnn_nearest_k_batch(db, db.shape[0], db.shape[1], q, len(q), kk, idx, dist, max(int(nthreads), 1))
    if exclude_ids is None:
        return idx, dist
    ex = np.asarray(exclude_ids, np.int64).reshape(-1, 1)
    keep = idx != ex
    # rows that did not contain the excluded id (possible only through a distance tie at the k-th place) drop their last entry
    drop_last = keep.all(axis=1)
    keep[drop_last, -1] = False
    return idx[keep].reshape(len(q), k), dist[keep].reshape(len(q), k)
