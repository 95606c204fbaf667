#!/usr/bin/env python
"""A short run of every kernel path (tiny, lane group incl. vf + box, staged box / pairwise / vf, standalone kernels,
resets, statistics) for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_swarm_b200 import _lib, tapes  # noqa: E402
from gym_swarm_b200.envs import BatchedSwarmEnv  # noqa: E402


def run(E, N, S, T, path="auto", vf=None, **kw):
    rng = np.random.default_rng(N + S)
    env = BatchedSwarmEnv(E, N, S, auto_reset=True, track_stats=True, reset_slots=3, per_agent=True, debug_outputs=True,
                          path=path, **kw)
    rt = tapes.make_reset_tapes(rng, 3, E, N, S, KR=env.KR, KP=env.KP)
    st = tapes.make_step_tapes(rng, T, E, N, K=env.K, n_actions=9)
    env.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    for t in range(T):
        env.step(st.actions[t], {"attraction": 1.0, "repulsion": 0.5, "alignment": 0.25, "vf_size": vf},
                 retarget=st.retarget[t], slab=st.slab[t])
    torch.cuda.synchronize()
    s = env.stats()
    print(f"ok E={E} N={N} S={S} path={env.path_name} vf={vf} episodes={s['episodes']:.0f}", flush=True)
    env.close()


def main():
    run(70, 8, 6, 6)                                   # tiny + masked reset kernel (dense: many resets / re-draws)
    run(40, 5, 9, 4, path="group")                     # fused<8,1>
    run(40, 8, 6, 4, vf=2)                             # tiny handle, vf step -> fused<8,1, vf>
    run(20, 16, 8, 5)                                  # fused<8,2> ring + bitmap
    run(9, 40, 9, 4, slab_len=2048, place_extra=4000, pred_attempts=64)   # fused<16,4>, dense
    run(6, 128, 64, 3)                                 # fused<32,4>
    run(6, 128, 64, 3, vf=3)                           # fused<32,4, vf>
    run(6, 120, 64, 3, flags=_lib.FLAG_NO_FAR_BORDER)  # fused<32,4>, all four thresholds in the ring sweep
    run(4, 200, 300, 3)                                # fused<32,8>, no bitmap (all-pairs collision counts)
    run(3, 300, 96, 3)                                 # staged + box
    run(3, 300, 96, 2, path="staged_pairwise")         # staged symmetric pairwise
    run(2, 300, 96, 2, vf=4)                           # staged vf (TMA-staged row sweep)
    run(2, 1000, 100, 2)                               # staged, non word-aligned rows, slab overflow fallback
    run(5, 12, 4, 6, path="staged", slab_len=64, place_extra=4000, pred_attempts=64)   # staged, tiny dense env
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for (E, N, S) in ((33, 37, 50), (64, 8, 16), (3, 4096, 256)):
        x = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev)
        y = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev)
        a = torch.randint(0, 9, (E, N), dtype=torch.uint8, device=dev)
        ox, oy = torch.empty_like(x), torch.empty_like(y)
        _lib.check(lib.swarm_move(E, N, S, p(x), p(y), p(a), p(ox), p(oy), s))
        ret = torch.ones(E, dtype=torch.uint8, device=dev)
        px = torch.zeros(E, dtype=torch.int16, device=dev)
        py = torch.zeros(E, dtype=torch.int16, device=dev)
        tg = torch.zeros(E, dtype=torch.int32, device=dev)
        ori = torch.zeros(E, dtype=torch.uint8, device=dev)
        prev = torch.zeros((E, 2), dtype=torch.int16, device=dev)
        done = torch.zeros(E, dtype=torch.uint8, device=dev)
        eat = torch.zeros(E, dtype=torch.int32, device=dev)
        _lib.check(lib.swarm_predator(E, N, S, p(x), p(y), p(ox), p(oy), p(ret), p(px), p(py), p(tg), p(ori), p(prev),
                                      p(done), p(eat), None, s))
        out = torch.zeros((E, N, 2), dtype=torch.int32, device=dev)
        nb = int(lib.swarm_pairwise_scratch_bytes(E, N))
        scr = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(lib.swarm_pairwise(E, N, S, 5, 2, p(x), p(y), None, p(out), p(scr), nb, s))
    torch.cuda.synchronize()
    print("sanitize case done", flush=True)


if __name__ == "__main__":
    main()
