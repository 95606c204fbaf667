#!/usr/bin/env python
"""Roofline runs of the standalone C-ABI kernels (swarm_move / swarm_predator / swarm_pairwise) at
streaming sizes (>= 1 GB of algorithmic traffic per launch for the HBM-bound ones, so L2 residency cannot
flatter the number).  One JSON line per kernel; CUDA events on the launching stream, 3 warm-up launches.

    python tools/kernel_roofline.py [move|predator|pairwise|peaks ...]
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_swarm_b200 import _lib  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(PEAKS["hbm_gbs"])


def ptr(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    v = [a.elapsed_time(b) for a, b in ts]
    return float(np.mean(v)), float(np.min(v))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run_move(lib, dev):
    E, N, S = 1 << 24, 8, 16                       # 134 M agents: 9 B each = 1.2 GB per launch
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    y = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    a = torch.randint(0, 8, (E, N), dtype=torch.uint8, device=dev, generator=g)
    ox, oy = torch.empty_like(x), torch.empty_like(y)
    fn = lambda: _lib.check(lib.swarm_move(E, N, S, ptr(x), ptr(y), ptr(a), ptr(ox), ptr(oy), stream()))
    mean, best = timeit(fn)
    # spot-check against the closed form (SDE:60-80)
    dx = torch.tensor([-1, -1, 0, 1, 1, 1, 0, -1, 0], device=dev, dtype=torch.int16)[a[:4096].long()]
    ref = (x[:4096] + dx) % S
    assert torch.equal(ref.to(torch.int16), ox[:4096]), "swarm_move mismatch"
    alg = E * N * 9
    return {"kernel": "swarm_move (move_kernel)", "shape": f"{E} envs x {N} agents, S={S}", "bound": "hbm",
            "algorithmic_bytes": alg, "bytes_per_agent": 9, "ms_mean": mean, "ms_best": best,
            "achieved": alg / (mean * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s", "frac": alg / (mean * 1e-3) / 1e9 / HBM}


def run_predator(lib, dev, N=128, E=1 << 20, S=128, use_tcell=True):
    g = torch.Generator(device=dev).manual_seed(2)
    xo = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    yo = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    xn = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    yn = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    ret = (torch.rand((E,), device=dev, generator=g) < 0.1).to(torch.uint8)
    px = torch.randint(0, S, (E,), dtype=torch.int16, device=dev, generator=g)
    py = torch.randint(0, S, (E,), dtype=torch.int16, device=dev, generator=g)
    tgt = torch.randint(0, N, (E,), dtype=torch.int32, device=dev, generator=g)
    ori = torch.zeros((E,), dtype=torch.uint8, device=dev)
    prev = torch.zeros((E, 2), dtype=torch.int16, device=dev)
    done = torch.zeros((E,), dtype=torch.uint8, device=dev)
    eaten = torch.zeros((E,), dtype=torch.int32, device=dev)
    # the cached cell of the current target (swarm_state.tcell), as the staged step keeps it
    tcell = torch.stack([xo.gather(1, tgt.long().unsqueeze(1)).squeeze(1), yo.gather(1, tgt.long().unsqueeze(1)).squeeze(1)],
                        1).contiguous() if use_tcell else None
    fn = lambda: _lib.check(lib.swarm_predator(E, N, S, ptr(xo), ptr(yo), ptr(xn), ptr(yn), ptr(ret), ptr(px), ptr(py),
                                               ptr(tgt), ptr(ori), ptr(prev), ptr(done), ptr(eaten),
                                               ptr(tcell) if use_tcell else None, stream()))
    mean, best = timeit(fn)
    # new positions always (termination test), old positions for the 10 % of envs that retarget + the target's cell
    # (4 B read; + 4 B written back with the cache), per env: px,py,target read+write (16), retarget 1, orient 1,
    # pred_prev 4, done 1, eaten 4
    alg = int(E * (4 * N + 0.1 * 4 * N + 4 + 27 + (4 if use_tcell else 0)))
    return {"kernel": "swarm_predator" + (" (target cell cached)" if use_tcell else " (no tcell)"),
            "shape": f"{E} envs x {N} agents, S={S}", "bound": "hbm",
            "algorithmic_bytes": alg, "bytes_per_env": alg / E, "ms_mean": mean, "ms_best": best,
            "achieved": alg / (mean * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s", "frac": alg / (mean * 1e-3) / 1e9 / HBM}


def run_observe(lib, dev, E=1 << 18, N=128, S=128, vf=2):
    g = torch.Generator(device=dev).manual_seed(4)
    x = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    y = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    px = torch.randint(0, S, (E,), dtype=torch.int16, device=dev, generator=g)
    py = torch.randint(0, S, (E,), dtype=torch.int16, device=dev, generator=g)
    W = 2 * vf + 1
    out = torch.empty((E, N, W, W), dtype=torch.uint8, device=dev)
    fn = lambda: _lib.check(lib.swarm_observe_vf(E, N, S, vf, ptr(x), ptr(y), ptr(px), ptr(py), ptr(out), stream()))
    mean, best = timeit(fn)
    alg = E * N * (4 + W * W) + E * 4                 # positions in, windows out
    return {"kernel": "swarm_observe_vf (observe_vf_warp_kernel)", "shape": f"{E} envs x {N} agents, S={S}, vf={vf}", "bound": "hbm",
            "algorithmic_bytes": alg, "ms_mean": mean, "ms_best": best, "achieved": alg / (mean * 1e-3) / 1e9, "peak": HBM,
            "unit": "GB/s", "frac": alg / (mean * 1e-3) / 1e9 / HBM}


def int32_peaks(lib):
    out = {}
    for kind, name in ((0, "iadd3"), (1, "imnmx"), (2, "iadd3+imad"), (3, "viaddmnmx_s16x2"), (4, "viaddmnmx_s16x2+imad")):
        v = C.c_double()
        _lib.check(lib.swarm_int32_peak(0, kind, C.byref(v), stream()))
        out[name] = v.value
    return out


def run_pairwise(lib, dev, peaks):
    E, N, S = 16, 16384, 1024                      # BASELINE configs[3]
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    y = torch.randint(0, S, (E, N), dtype=torch.int16, device=dev, generator=g)
    out = torch.empty((E, N, 2), dtype=torch.int32, device=dev)
    nbytes = int(lib.swarm_pairwise_scratch_bytes(E, N))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    fn = lambda: _lib.check(lib.swarm_pairwise(E, N, S, 5, 2, ptr(x), ptr(y), None, ptr(out), ptr(scratch), nbytes,
                                               stream()))
    mean, best = timeit(fn, reps=5)
    pairs = E * N * (N - 1)
    ops14 = pairs * 14                              # 14 scalar INT32 lane-ops per ordered pair
    # executed by pairwise_sym_kernel per word evaluation (2 unordered = 4 ordered pairs): 9 ALU-pipe instructions
    # (8 DPX + loop share) and 6 FMA-pipe IMADs -> 2.25 ALU lane-ops per ordered pair
    exec_ops = pairs * 2.25
    pk = peaks["iadd3"]
    return {"kernel": "swarm_pairwise (pairprep + pairwise_sym_kernel + pair_counts)", "shape": f"{E} envs x {N} agents, S={S}",
            "bound": "int32-pipe", "ordered_pairs": pairs, "ms_mean": mean, "ms_best": best,
            "algorithmic_scalar_ops": ops14, "achieved_scalar_T": ops14 / (mean * 1e-3) / 1e12,
            "executed_alu_lane_ops": exec_ops, "achieved": exec_ops / (mean * 1e-3) / 1e12, "peak": pk / 1e12,
            "unit": "T lane-ops/s", "frac": exec_ops / (mean * 1e-3) / pk,
            "note": "frac = executed ALU-pipe lane-ops (2.25 per ordered pair: DPX 16x2 handles two pairs per instruction and "
                    "the symmetric sweep credits both directions) over the measured IADD3 issue peak; achieved_scalar_T "
                    "counts the 14 scalar ops per ordered pair of SURVEY.md 8d and therefore exceeds the scalar peak"}


def main():
    which = sys.argv[1:] or ["peaks", "move", "predator", "pairwise", "observe"]
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    peaks = int32_peaks(lib)
    if "peaks" in which:
        print(json.dumps({"kernel": "int32_peak", "lane_ops_per_s_T": {k: v / 1e12 for k, v in peaks.items()},
                          "hbm_gbs": HBM}), flush=True)
    if "move" in which:
        print(json.dumps(run_move(lib, dev)), flush=True)
        torch.cuda.empty_cache()
    if "predator" in which:
        print(json.dumps(run_predator(lib, dev)), flush=True)
        torch.cuda.empty_cache()
        print(json.dumps(run_predator(lib, dev, N=16384, E=8192, S=1024)), flush=True)
        torch.cuda.empty_cache()
        print(json.dumps(run_predator(lib, dev, N=8, E=1 << 24, S=16)), flush=True)
        torch.cuda.empty_cache()
        print(json.dumps(run_predator(lib, dev, use_tcell=False)), flush=True)      # round-1 form: dependent random read
        torch.cuda.empty_cache()
        print(json.dumps(run_predator(lib, dev, N=8, E=1 << 24, S=16, use_tcell=False)), flush=True)
        torch.cuda.empty_cache()
    if "pairwise" in which:
        print(json.dumps(run_pairwise(lib, dev, peaks)), flush=True)
        torch.cuda.empty_cache()
    if "observe" in which:
        print(json.dumps(run_observe(lib, dev)), flush=True)


if __name__ == "__main__":
    main()
