#!/usr/bin/env python
"""Race canary for the lane-group kernels (compute-sanitizer is closed on this pool).

The fused kernels synchronise the lanes of one environment with `__syncwarp(mask)` instead of `__syncthreads()` and
exchange data through per-env shared memory.  A missing barrier there is invisible as long as a converged warp runs in
lockstep.  Build the library with -DSWARM_RACE_JITTER (tools/build_variant.sh jitter -DSWARM_RACE_JITTER): after every
lane-group barrier each lane then sleeps a lane- and call-dependent number of nanoseconds, which pushes the lanes out of
lockstep between barriers, so an exchange that relies on lockstep instead of a barrier produces different results.

    SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_jitter.so python tools/race_canary.py [steps]

Every shape below (dense grids: many collision rounds, resets and border-pass pairs; both launch modes; both
visual-field modes) is stepped `steps` times (default 1000) against the C oracle: positions, predator, target, done,
draw counts and integer reward counts must be bit-identical, float rewards within 1e-6 relative.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_utils import assert_float_parity  # noqa: E402
from gym_swarm_b200 import _lib, tapes  # noqa: E402
from gym_swarm_b200.envs import BatchedSwarmEnv  # noqa: E402
from oracle import c_oracle  # noqa: E402

SHAPES = [  # E, N, S, vf, flags, path
    (96, 16, 6, None, 0, "auto"), (64, 40, 9, None, 0, "auto"), (48, 128, 16, None, 0, "auto"),
    (48, 128, 128, None, 0, "auto"), (48, 128, 128, None, _lib.FLAG_FUSED_SPLIT, "auto"),
    (48, 128, 128, None, _lib.FLAG_FUSED_SPLIT | _lib.FLAG_NO_BOX_NEAR, "auto"), (32, 255, 40, None, 0, "auto"),
    (32, 200, 300, None, _lib.FLAG_FUSED_SPLIT, "auto"), (64, 8, 6, None, 0, "group"), (48, 64, 64, 3, 0, "auto"),
    (48, 128, 128, 5, _lib.FLAG_FUSED_SPLIT, "auto"), (48, 60, 24, 2, _lib.FLAG_VF_DENSE, "auto"),
    (48, 128, 64, None, _lib.FLAG_NO_FAR_BORDER, "auto"),
]


def run(E, N, S, vf, flags, path, T):
    rng = np.random.default_rng(E * 7 + N + S)
    dense = S * S < 6 * N
    env = BatchedSwarmEnv(E, N, S, auto_reset=True, reset_slots=5, per_agent=True, debug_outputs=True, path=path, flags=flags,
                          slab_len=1024 if dense else 512, place_extra=4000 if dense else 600, pred_attempts=64)
    rt = tapes.make_reset_tapes(rng, env.R, E, N, S, KR=env.KR, KP=env.KP)
    env.set_reset_tapes(rt.place, rt.ptape)
    ora = c_oracle.COracleBatch(E, N, S, 5, 2, -10.0, auto_reset=True, K=env.K)
    ora.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    ora.reset()
    done = 0
    for t0 in range(0, T, 50):
        st = tapes.make_step_tapes(rng, min(50, T - t0), E, N, K=env.K, n_actions=9)
        for t in range(st.actions.shape[0]):
            w = (1.0, 0.5, 0.25)
            a = env.step_numpy(st.actions[t], st.slab[t], st.retarget[t], w, vf)
            b = ora.step(st.actions[t], st.slab[t], st.retarget[t], w, vf, True)
            for k in ("x", "y", "px", "py", "target", "done", "eaten", "cnt", "consumed", "agent_cnt", "eff"):
                if not np.array_equal(a[k], b[k]):
                    raise AssertionError(f"{env.path_name} N={N} S={S} vf={vf} flags={flags}: step {t0 + t}: {k} differs")
            assert_float_parity(a["glob"], b["glob"], err_msg=f"step {t0 + t} glob")
            assert_float_parity(a["agent"], b["agent"], err_msg=f"step {t0 + t} agent")
            assert not np.any(a["err"] & ~np.uint32(16)), f"error bits {a['err'].max()}"
            done += int(a["done"].sum())
    env.close()
    ora.close()
    print(f"ok  {env.path_name:38s} E={E:3d} N={N:3d} S={S:3d} vf={vf} flags={flags:2d}: {T} steps, {done} episodes finished, "
          f"bit-identical to the oracle", flush=True)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    print(f"library: {_lib.LIB_PATH}", flush=True)
    for shp in SHAPES:
        run(*shp, T)
    print("race canary passed", flush=True)


if __name__ == "__main__":
    main()
