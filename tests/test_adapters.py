"""The C++ OMPL adapters (include/b200mp_ompl.hpp) compiled against the GENUINE OMPL base-class headers of the reference
checkout where it exists (State.h, StateValidityChecker.h, MotionValidator.h, NearestNeighbors.h, NearestNeighborsLinear.h,
Exception.h; Boost's concept_check and RealVectorStateSpace::StateType come from tests/ompl_shim because the image has
neither Boost nor Eigen), and against interface-identical mocks (tests/ompl_mock) everywhere else.
CPU: both builds compile and the binary refuses to run without a GPU. GPU: tests/cpp/adapter_test.cpp passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def compile_adapter_test(flavour: str) -> str:
    import __graft_entry__ as g
    g.build()
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, f"adapter_test_{flavour}")
    src = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in os.listdir(os.path.join(ROOT, "include"))]
    hdrs += [os.path.join(ROOT, "tests", "ompl_shim", h) for h in os.listdir(os.path.join(ROOT, "tests", "ompl_shim")) if h.endswith(".h")]
    lib = os.path.join(ROOT, "ompl_b200", "libb200mp.so")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(p) for p in [src, lib] + hdrs):
        return out
    inc = [f"-I{ROOT}/include", f"-I{ROOT}/tests/ompl_shim"]
    inc += [f"-I{REF_SRC}"] if flavour == "genuine" else [f"-I{ROOT}/tests/ompl_mock"]
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", *inc,
           '-DB200MP_OMPL_STATE_HEADER="b200_mock_real_vector_state.h"', '-DB200MP_OMPL_PLANNERDATA_HEADER="b200_mock_planner_data.h"', src, "-o", out,
           f"-L{ROOT}/ompl_b200", "-lb200mp", f"-Wl,-rpath,{ROOT}/ompl_b200", "-Wl,-rpath,$ORIGIN/../../../ompl_b200"]
    subprocess.run(cmd, check=True)
    return out


def flavours():
    return (["genuine"] if os.path.isdir(os.path.join(REF_SRC, "ompl", "base")) else []) + ["mock"]


@pytest.mark.parametrize("flavour", ["genuine", "mock"])
def test_adapters_compile_and_fail_loudly_without_gpu(flavour):
    if flavour not in flavours():
        pytest.skip("reference checkout not present")
    exe = compile_adapter_test(flavour)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_adapters_on_gpu():
    ran = 0
    for flavour in ("genuine", "mock"):
        exe = os.path.join(BUILD, f"adapter_test_{flavour}")
        if flavour in flavours():
            exe = compile_adapter_test(flavour)  # (rebuilt when the headers or the test are newer than the binary)
        elif not os.path.exists(exe):
            continue
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "all passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
        ran += 1
    assert ran >= 1
