"""ompl_b200 -- B200-native batched motion validation + nearest neighbours behind OMPL's plugin surface.

The product is the CUDA library `libb200mp.so` (C ABI in include/b200mp.h) plus the C++ OMPL adapters in
include/b200mp_ompl.hpp. This Python package is thin ctypes plumbing over the same C ABI, used by bench.py,
the tests and the multi-GPU roadmap builder; it never falls back to a CPU implementation.
"""
from .engine import (B200MPError, Comm, Context, Environment, KinematicChain, NearestNeighbors, ROBOTS, Session, capt_build, horn_environment, lib_path,  # noqa: F401
                     comm_unique_id, load_library, robot_info)
