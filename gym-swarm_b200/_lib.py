"""ctypes binding of ``libswarm_b200.so`` (C ABI: ``include/swarm_abi.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present the product path raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWARM_B200_LIB selects another build of the same library (kernel A/B experiments); never a CPU path
LIB_PATH = os.environ.get("SWARM_B200_LIB") or os.path.join(_HERE, "libswarm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "swarm_abi.h")

ABI_VERSION = 3
PATH_AUTO, PATH_STAGED, PATH_FUSED, PATH_GROUP, PATH_STAGED_PAIRWISE = 0, 1, 2, 3, 4

E_NOT_READY = -5
FLAG_NO_FAR_BORDER, FLAG_VF_DENSE, FLAG_FUSED_SPLIT, FLAG_FUSED_ONE_LAUNCH, FLAG_NO_BOX_NEAR = 1, 4, 8, 16, 32
FLAG_TINY_SCALAR = 64

ERR_SLAB_OVERFLOW = 1
ERR_PLACE_OVERFLOW = 2
ERR_PRED_OVERFLOW = 4
ERR_BAD_ACTION = 8
ERR_STEP_WHEN_DONE = 16

MAX_AGENTS = 16384
MAX_GRID = 16384
FUSED_MAX_AGENTS = 256


class SwarmLibraryError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("num_envs", C.c_int32),
                ("num_agents", C.c_int32), ("grid_size", C.c_int32), ("attraction_thresh", C.c_int32),
                ("repulsion_thresh", C.c_int32), ("predator_eat_rew", C.c_float), ("slab_len", C.c_int32),
                ("place_len", C.c_int32), ("pred_attempts", C.c_int32), ("reset_slots", C.c_int32),
                ("auto_reset", C.c_int32), ("track_stats", C.c_int32), ("path", C.c_int32),
                ("flags", C.c_uint32), ("tiny16_min_envs", C.c_int32), ("env_offset", C.c_uint32),
                ("reserved", C.c_int32)]


class State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x", "y", "px", "py", "target", "pred_orient", "frozen", "err", "reset_count", "tcell",
                 "ep_rec", "fin_count", "fin_len", "fin_ret")]


class StepIn(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("retarget", C.c_void_p), ("slab", C.c_void_p),
                ("w_attraction", C.c_double), ("w_repulsion", C.c_double), ("w_alignment", C.c_double),
                ("vf_size", C.c_int32), ("rng_step", C.c_uint32), ("rng_seed", C.c_uint64)]


class StepOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("done", "eaten", "reward", "counts", "pred_prev", "pred_next", "step_orient", "step_target",
                 "consumed", "term_x", "term_y", "agent_reward", "agent_counts", "eff_action", "agent_terms")]


# every symbol include/swarm_abi.h declares: name -> (restype, argtypes)
_VP, _I32, _U64, _DP = C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_double)
EXPORTS = {
    "swarm_create": (C.c_int, [C.POINTER(Config), C.POINTER(_VP)]),
    "swarm_destroy": (C.c_int, [_VP]),
    "swarm_last_error": (C.c_char_p, [_VP]),
    "swarm_abi_version": (C.c_int, []),
    "swarm_bind_state": (C.c_int, [_VP, C.POINTER(State)]),
    "swarm_bind_reset_tapes": (C.c_int, [_VP, _VP, _VP]),
    "swarm_set_device_reset_tapes": (C.c_int, [_VP, C.c_uint64]),
    "swarm_reset": (C.c_int, [_VP, _VP]),
    "swarm_step": (C.c_int, [_VP, C.POINTER(StepIn), C.POINTER(StepOut), _VP]),
    "swarm_path_name": (C.c_char_p, [_VP]),
    "swarm_last_launch_count": (C.c_int, [_VP]),
    "swarm_move": (C.c_int, [_I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "swarm_predator": (C.c_int, [_I32, _I32, _I32] + [_VP] * 14),
    "swarm_pairwise": (C.c_int, [_I32, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _U64, _VP]),
    "swarm_pairwise_scratch_bytes": (_U64, [_I32, _I32]),
    "swarm_autoreset": (C.c_int, [_VP, _VP, _VP]),
    "swarm_observe_vf": (C.c_int, [_I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "swarm_stats_begin": (C.c_int, [_VP, _I32, _VP]),
    "swarm_stats_end": (C.c_int, [_VP, _DP, _I32]),
    "swarm_stats_read": (C.c_int, [_VP, _DP, _VP]),
    "swarm_nccl_adopt": (C.c_int, [_VP, _VP]),
    "swarm_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "swarm_nccl_init": (C.c_int, [_VP, C.POINTER(C.c_uint8), _I32, _I32]),
    "swarm_stats_allreduce": (C.c_int, [_VP, _DP, _VP]),
    "swarm_int32_peak": (C.c_int, [_I32, _I32, _DP, _VP]),
}

_lib = None


def load():
    """Load the CUDA library; raises SwarmLibraryError (never falls back to a CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SwarmLibraryError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.swarm_abi_version() != ABI_VERSION:
            raise SwarmLibraryError("libswarm_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().swarm_last_error(handle)
        raise SwarmLibraryError(f"swarm C-ABI error {rc}: {msg.decode() if msg else '?'}")
