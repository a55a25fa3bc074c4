"""CPU tests: the C-ABI library loads, exports every symbol include/*.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            syms |= set(re.findall(r"\b(b200mp_[a-z0-9_]+)\s*\(", open(os.path.join(inc, f)).read()))
    return sorted(syms)


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return ctypes.CDLL(os.path.join(ROOT, "ompl_b200", "libb200mp.so"))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_robot_info_matches_reference_constants(lib):
    import ompl_b200
    p = ompl_b200.robot_info("panda")
    assert (p["dim"], p["n_spheres"], p["resolution"]) == (7, 59, 32)   # robots/panda.hh:14-18
    f = ompl_b200.robot_info("fetch")
    assert (f["dim"], f["n_spheres"], f["resolution"]) == (8, 111, 32)  # robots/fetch.hh:14-18


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ompl_b200
    with pytest.raises(ompl_b200.B200MPError) as e:
        ompl_b200.Context(0)
    assert e.value.code == 3  # B200MP_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    # only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may touch oracle/
    for top in ("ompl_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                    src = open(os.path.join(d, f)).read()
                    for needle in ("from oracle", "import oracle", "liboracle", "libref_", "oracle/"):
                        assert needle not in src, (f, needle)
