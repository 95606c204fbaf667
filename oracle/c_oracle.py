"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/swarm_oracle.c (the plain-C restatement).

Exposes the same batch interface as ``swarm_oracle.OracleBatch`` (``reset()`` / ``step(...)`` returning
dicts of numpy arrays).  Used by tests/ as a fast second oracle and by bench.py as the CPU baseline
("port", OpenMP over envs).  Never imported by the product path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "swarm_oracle.c")
LIB = os.path.join(HERE, "libswarm_oracle.so")

_lib = None


def build(force=False):
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_int] * 5 + [C.c_double] + [C.c_int] * 5
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_set_reset_tapes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_reset.argtypes = [C.c_void_p]
        lib.oracle_step.argtypes = [C.c_void_p] * 4 + [C.c_double] * 3 + [C.c_int] + [C.c_void_p] * 14
        lib.oracle_get_state.argtypes = [C.c_void_p] * 8
        lib.oracle_set_threads.restype = C.c_int
        lib.oracle_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_threads(n):
    """Use n OpenMP threads (returns the count in effect)."""
    return int(load().oracle_set_threads(int(n)))


class COracleBatch:
    def __init__(self, E, N, S, att=5, rep=2, eat=-10.0, auto_reset=False, K=None, PL=None, KP=None, R=None):
        self.lib = load()
        self.E, self.N, self.S = E, N, S
        self.att, self.rep, self.eat, self.auto_reset = int(np.ceil(att)), int(np.ceil(rep)), float(eat), auto_reset
        self.K, self.PL, self.KP, self.R = K, PL, KP, R
        self.h = None

    def set_reset_tapes(self, place, ptape):
        self.place = np.ascontiguousarray(place, dtype=np.int16)
        self.ptape = np.ascontiguousarray(ptape, dtype=np.int16)
        self.R, _, self.PL = self.place.shape
        self.KP = self.ptape.shape[2]

    def _ensure(self, K):
        if self.h is None:
            self.K = K
            self.h = self.lib.oracle_create(self.E, self.N, self.S, self.att, self.rep, self.eat, self.K, self.PL,
                                            self.KP, self.R, int(self.auto_reset))
            self.lib.oracle_set_reset_tapes(self.h, _p(self.place), _p(self.ptape))

    def close(self):
        if self.h is not None:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def state(self):
        E, N = self.E, self.N
        st = dict(x=np.zeros((E, N), np.int16), y=np.zeros((E, N), np.int16), px=np.zeros(E, np.int16),
                  py=np.zeros(E, np.int16), target=np.zeros(E, np.int32), err=np.zeros(E, np.uint32),
                  frozen=np.zeros(E, np.uint8))
        self.lib.oracle_get_state(self.h, *[_p(st[k]) for k in ("x", "y", "px", "py", "target", "err", "frozen")])
        return st

    def reset(self, K=32):
        self._ensure(self.K if self.K is not None else K)
        self.lib.oracle_reset(self.h)
        return self.state()

    def step(self, actions, slab, retarget, weights=(1.0, 1.0, 1.0), vf=None, want_agents=True, want_state=True):
        E, N = self.E, self.N
        if vf is not None and vf < 0:          # SDE:355 clears every pair (D >= 0 > vf): alignment is identically 0
            vf, weights = None, (weights[0], weights[1], 0.0)
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        slab = np.ascontiguousarray(slab, dtype=np.uint8)
        retarget = np.ascontiguousarray(retarget, dtype=np.uint8)
        assert slab.shape == (E, self.K), (slab.shape, self.K)
        o = dict(done=np.zeros(E, np.uint8), eaten=np.full(E, -1, np.int32), glob=np.zeros((E, 5)),
                 cnt=np.zeros((E, 2), np.int64), pred_prev=np.zeros((E, 2), np.int16),
                 pred_next=np.zeros((E, 2), np.int16), pred_orient=np.zeros(E, np.uint8),
                 step_target=np.zeros(E, np.int32), consumed=np.zeros(E, np.int32),
                 term_x=np.zeros((E, N), np.int16), term_y=np.zeros((E, N), np.int16),
                 eff=np.zeros((E, N), np.uint8))
        if want_agents:
            o["agent"] = np.zeros((E, N, 5))
            o["agent_cnt"] = np.zeros((E, N, 2), np.int64)
        self.lib.oracle_step(self.h, _p(actions), _p(retarget), _p(slab), float(weights[0]), float(weights[1]),
                             float(weights[2]), -1 if vf is None else int(np.floor(vf)), _p(o["done"]), _p(o["eaten"]),
                             _p(o["glob"]), _p(o["cnt"]), _p(o["pred_prev"]), _p(o["pred_next"]), _p(o["pred_orient"]),
                             _p(o["step_target"]), _p(o["consumed"]), _p(o["term_x"]), _p(o["term_y"]),
                             _p(o.get("agent")), _p(o.get("agent_cnt")), _p(o["eff"]))
        if want_state:
            o.update(self.state())
        return o
