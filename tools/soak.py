#!/usr/bin/env python
"""Soak run: the full-size property checks of tests/test_gpu_properties.py over hundreds of steps (rare events: many
resets, multi-round re-draws, border cases) and cross-path equality over a longer horizon."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from test_gpu_properties import _run_properties  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for (E, N, S, per_agent) in [(65536, 128, 128, True), (1 << 20, 8, 16, False), (16, 16384, 1024, True), (4096, 16, 32, False),
                             (1 << 17, 16, 32, False)]:
    t0 = time.time()
    _run_properties(E, N, S, T=T, per_agent=per_agent)
    print(f"soak ok: {E} x {N} @ {S}, {T} steps, {time.time() - t0:.1f} s", flush=True)
for (E, N, S, paths) in [(8192, 128, 128, ("auto", "staged")), (1 << 16, 8, 16, ("auto", "group")), (4, 16384, 1024, ("auto", "staged_pairwise")),
                         (1 << 16, 16, 32, ("auto", "group"))]:
    ref = None
    for path in paths:
        out = _run_properties(E, N, S, T=60, per_agent=True, path=path, seed=11)
        if ref is None:
            ref = out
            continue
        for k in ref:
            ok = torch.allclose(ref[k], out[k], rtol=1e-6, atol=1e-9) if ref[k].is_floating_point() else torch.equal(ref[k], out[k])
            assert ok, (E, N, S, path, k)
    print(f"soak paths agree: {E} x {N} @ {S}: {paths}", flush=True)
