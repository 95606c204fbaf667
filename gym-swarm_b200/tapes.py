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


# ------------------------------------------------------------------------------------------------------
# Counter-based tapes: the NumPy twin of the in-kernel Philox4x32-10 streams of
# gym-swarm_b200/csrc/swarm_common.cuh.  A draw is a pure function of (seed, stream, env, index); these functions
# materialise exactly the values the kernels compute when no tape is bound, so the oracle can be fed the same draws.
# ------------------------------------------------------------------------------------------------------
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 arrays (broadcast); returns the four output words."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _M32 for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & _M32, np.uint64(k1) & _M32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & _M32, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & _M32
        k0 = (k0 + np.uint64(0x9E3779B9)) & _M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & _M32
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def philox_step_tapes(seed, step, E, K):
    """(retarget u8[E], slab u8[E,K]) of step ``step``: stream 1, counter (env, step, block, 1)."""
    k0, k1 = _key(seed)
    env = np.arange(E, dtype=np.uint64)
    retarget = (philox4x32_10(env, step, 0, 1, k0, k1)[0] < np.uint32(429496730)).astype(np.uint8)
    slab = np.zeros((E, K), dtype=np.uint8)
    for blk in range((K + 15) // 16):
        w = philox4x32_10(env, step, 1 + blk, 1, k0, k1)
        for d in range(16 * blk, min(K, 16 * blk + 16)):
            slab[:, d] = ((w[(d >> 2) & 3] >> np.uint32(8 * (d & 3))) & np.uint32(7)).astype(np.uint8)
    return retarget, slab


def philox_reset_tapes(seed, R, E, N, S, KR=None, KP=8):
    """ResetTapes for reset indices 0..R-1: stream 2 (placement) and stream 3 (predator attempts).  The device uses the
    reset index itself as the stream counter, so these equal its draws as long as no env resets more than R times."""
    KR = default_place_extra(N) if KR is None else KR
    PL = 2 * N + 2 * KR
    k0, k1 = _key(seed)
    env = np.arange(E, dtype=np.uint64)
    place = np.zeros((R, E, PL), dtype=np.int16)
    ptape = np.zeros((R, E, KP, 2), dtype=np.int16)
    for r in range(R):
        for blk in range((PL + 3) // 4):
            w = philox4x32_10(env, r, blk, 2, k0, k1)
            for i in range(4 * blk, min(PL, 4 * blk + 4)):
                place[r, :, i] = ((w[i & 3].astype(np.uint64) * np.uint64(S)) >> np.uint64(32)).astype(np.int16)
        for blk in range((KP + 1) // 2):
            w = philox4x32_10(env, r, blk, 3, k0, k1)
            for a in range(2 * blk, min(KP, 2 * blk + 2)):
                ptape[r, :, a, 0] = ((w[2 * (a & 1)].astype(np.uint64) * np.uint64(S)) >> np.uint64(32)).astype(np.int16)
                ptape[r, :, a, 1] = ((w[2 * (a & 1) + 1].astype(np.uint64) * np.uint64(S)) >> np.uint64(32)).astype(np.int16)
    return ResetTapes(place, ptape, np.zeros((R, E, N), dtype=np.uint8))
