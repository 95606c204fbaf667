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


def _header_struct_fields(name):
    """Field names of `typedef struct <name> { ... } <name>;` in include/swarm_abi.h, in declaration order."""
    src = open(_lib.HEADER_PATH).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]


def test_struct_layout_matches_header():
    # the ctypes mirrors carry the header's fields in the header's order ...
    for cname, mirror in (("swarm_config", _lib.Config), ("swarm_state", _lib.State), ("swarm_step_in", _lib.StepIn),
                          ("swarm_step_out", _lib.StepOut)):
        assert _header_struct_fields(cname) == [f[0] for f in mirror._fields_], cname
    # ... and their sizes (all pointers 8 bytes, int32/uint32/float 4, double 8)
    assert ctypes.sizeof(_lib.Config) == 19 * 4
    assert ctypes.sizeof(_lib.State) == 14 * 8
    assert ctypes.sizeof(_lib.StepIn) == 3 * 8 + 3 * 8 + 4 + 4 + 8
    assert ctypes.sizeof(_lib.StepOut) == 15 * 8


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


def _cfg(**kw):
    base = dict(abi_version=_lib.ABI_VERSION, device=0, num_envs=4, num_agents=4, grid_size=20, attraction_thresh=5,
                repulsion_thresh=2, predator_eat_rew=-10.0, slab_len=32, place_len=2 * 4 + 64, pred_attempts=8,
                reset_slots=1, auto_reset=0, track_stats=0, path=_lib.PATH_AUTO)
    base.update(kw)
    return _lib.Config(**base)


@pytest.mark.parametrize("kw,needle", [
    (dict(abi_version=_lib.ABI_VERSION + 1), "ABI version"),
    (dict(num_envs=0), "num_envs"),
    (dict(num_agents=1), "num_agents"),
    (dict(num_agents=16385), "num_agents"),
    (dict(grid_size=1), "grid_size"),
    (dict(num_agents=25, grid_size=5, place_len=200), "more cells"),
    (dict(place_len=7), "tape geometry"),
    (dict(reset_slots=0), "tape geometry"),
    (dict(num_agents=300, place_len=2 * 300 + 64, path=_lib.PATH_FUSED), "fused path"),
])
def test_create_rejects_bad_configs_before_touching_cuda(built, kw, needle):
    """Argument validation (include/swarm_abi.h conventions: negative code + swarm_last_error text) runs before the
    first CUDA call, so it is checkable on a box without a GPU."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _cfg(**kw)
    rc = lib.swarm_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc < 0 and not h.value
    assert needle in lib.swarm_last_error(None).decode()


def test_create_without_gpu_reports_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _cfg()
    rc = lib.swarm_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc < 0 and not h.value
    assert "no CPU fallback" in lib.swarm_last_error(None).decode()


def test_null_arguments(built):
    lib = _lib.load()
    assert lib.swarm_create(None, None) < 0
    assert lib.swarm_reset(None, None) < 0
    lib.swarm_destroy(None)                                   # destroying a null handle is a no-op
    assert lib.swarm_path_name(None) == b""
    assert lib.swarm_last_launch_count(None) == 0


def _stub_gym(monkeypatch):
    """A stand-in for the gym registry (gym is not in this image): records register() calls and resolves entry points
    the way gym.make does ("module:Class")."""
    import importlib
    import sys
    import types
    registry = {}
    gym = types.ModuleType("gym")
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")

    def register(id, entry_point, **kw):
        registry[id] = entry_point

    def make(id, **kw):
        mod, cls = registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)(**kw)

    reg.register = register
    gym.make = make
    gym.envs = envs
    envs.registration = reg
    for name, m in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, m)
    return gym, registry


def test_register_gym_ids_and_entry_points(monkeypatch):
    """register_gym() registers the two ids of gym_swarm/__init__.py:6-14 and their entry points resolve to the drop-in
    classes (construction needs a GPU: tests/test_gpu_parity.py::test_gym_make_roundtrip)."""
    import importlib
    from gym_swarm_b200 import envs
    _, registry = _stub_gym(monkeypatch)
    assert envs.register_gym() == ["Discrete-Swarm-v0", "Continuous-Swarm-v0"]
    assert set(registry) == {"Discrete-Swarm-v0", "Continuous-Swarm-v0"}
    for id_, cls in (("Discrete-Swarm-v0", envs.DiscreteSwarmEnv), ("Continuous-Swarm-v0", envs.ContinuousSwarmEnv)):
        mod, name = registry[id_].split(":")
        assert getattr(importlib.import_module(mod), name) is cls
