"""Explicit random streams ("tapes") shared by the CUDA path and the oracle.

The reference draws from the process-global ``np.random`` at seven call sites
(``swarm_discrete_env.py:92,101,138,202,208,216,250``).  Here every one of those
draws is materialised once on the host and handed to both implementations:

reset tapes (one set per reset slot r = 0..R-1, per env)
    ``place``  int16[R, E, 2N + 2*KR]  values in [0,S) in consumption order: N x's,
               N y's, then per re-draw round L x's followed by L y's
               (``swarm_discrete_env.py:202-203, 208-209``).
    ``ptape``  int16[R, E, KP, 2]      predator placement attempts (``:92, :101``).
    ``orient`` uint8[R, E, N]          render-only orientation draws (``:216``); used
               by the single-env compatibility class and by the reference shim, never
               by the device.
step tapes (per step, per env)
    ``retarget`` uint8[E]   1 iff ``np.random.random() < 0.1`` (``:138-139``).
    ``slab``     uint8[E,K] k-th collision re-draw action (0..7) of this step (``:250``).

A finished env that auto-resets for the n-th time uses reset slot ``n % R``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def default_slab_len(N: int) -> int:
    """Re-draw slab length per env-step (32 for N<=128, 512 at N=16384)."""
    if N <= 128:
        return 32
    return int(min(4096, max(64, N // 32)))


def default_place_extra(N: int) -> int:
    """Extra (x,y) pairs per reset for placement re-draw rounds (a 16-pair tape was exhausted by one env in
    8 M x 100 steps at N = 8 on 16 x 16; 32 pairs leaves orders of magnitude of margin, benches at >= 1 M envs use 64)."""
    return int(max(32, N // 4))


@dataclass
class ResetTapes:
    place: np.ndarray    # int16[R,E,2N+2KR]
    ptape: np.ndarray    # int16[R,E,KP,2]
    orient: np.ndarray   # uint8[R,E,N]


@dataclass
class StepTapes:
    actions: np.ndarray   # uint8[T,E,N]
    retarget: np.ndarray  # uint8[T,E]
    slab: np.ndarray      # uint8[T,E,K]
    roll: np.ndarray      # float64[T,E] the uniform draw behind ``retarget``


def make_reset_tapes(rng: np.random.Generator, R: int, E: int, N: int, S: int,
                     KR: int | None = None, KP: int = 8) -> ResetTapes:
    KR = default_place_extra(N) if KR is None else KR
    place = rng.integers(0, S, size=(R, E, 2 * N + 2 * KR), dtype=np.int16)
    ptape = rng.integers(0, S, size=(R, E, KP, 2), dtype=np.int16)
    orient = rng.integers(0, 8, size=(R, E, N), dtype=np.uint8)
    return ResetTapes(place, ptape, orient)


def make_step_tapes(rng: np.random.Generator, T: int, E: int, N: int, K: int | None = None,
                    n_actions: int = 8, p_retarget: float = 0.1) -> StepTapes:
    K = default_slab_len(N) if K is None else K
    actions = rng.integers(0, n_actions, size=(T, E, N), dtype=np.uint8)
    roll = rng.random(size=(T, E))
    retarget = (roll < p_retarget).astype(np.uint8)
    slab = rng.integers(0, 8, size=(T, E, K), dtype=np.uint8)
    return StepTapes(actions, retarget, slab, roll)
