"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/dtqn_b200.h declares
(no compute calls -- there is no GPU here) and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dtqn_b200.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(dtqn_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "dtqn_b200", "libdtqn_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 5
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dtqn_b200.h but not exported"
    lib.dtqn_version.restype = ctypes.c_int
    header_ver = int(re.search(r"#define DTQN_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "dtqn_b200.h")).read()).group(1))
    assert lib.dtqn_version() == header_ver


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dtqn_b200 import _lib
    from dtqn_b200.envs import BatchedEnv
    with pytest.raises(_lib.DtqnLibError):
        BatchedEnv("DiscreteCarFlag-v0", 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dtqn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
