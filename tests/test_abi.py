"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/swarm_abi.h declares (no compute calls without a GPU), and the product path refuses to run
without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest

import gym_swarm_b200  # noqa: F401
from gym_swarm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return _lib.LIB_PATH


def header_functions():
    src = open(os.path.join(ROOT, "include", "swarm_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swarm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/swarm_abi.h but not exported"
    assert lib.swarm_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # field counts / sizes of the ctypes mirrors (all pointers 8 bytes, int32/float 4, double 8)
    assert ctypes.sizeof(_lib.Config) == 15 * 4
    assert ctypes.sizeof(_lib.State) == 14 * 8
    assert ctypes.sizeof(_lib.StepIn) == 3 * 8 + 3 * 8 + 4 + 4 + 8
    assert ctypes.sizeof(_lib.StepOut) == 14 * 8


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_swarm_b200.envs import BatchedSwarmEnv
    with pytest.raises(_lib.SwarmLibraryError):
        BatchedSwarmEnv(4, 4, 20)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "gym-swarm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
