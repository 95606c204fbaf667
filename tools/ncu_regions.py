#!/usr/bin/env python
"""Share of executed warp instructions / stall samples per code region of fused_step.cuh, from an ncu report
captured with --import-source on (regions are delimited by marker comments in the embedded source)."""
import csv, io, subprocess, sys
MARKERS = [("helpers (Seg, ld/st)", "struct Seg {"), ("occupancy_counts", "void occupancy_counts("), ("bitmap_counts", "bool bitmap_counts("),
           ("closest_target", "int closest_target("), ("reset_group", "void reset_group("), ("ring sweep", "void ring_sweep("),
           ("visual-field hits -> sums", "void vf_sparse_accumulate("), ("box counts (near thresholds)", "void box_near_counts("),
           ("far-threshold border pass", "void far_border_counts("), ("FusedCfg", "struct FusedCfg {"),
           ("terminal step (out of line)", "void terminal_step("), ("kernel prologue + loads", "fused_step_kernel(const __grid_constant__"),
           ("move + collision rounds", "// ---- move + collision"), ("predator", "// ---- predator (SDE"),
           ("termination + early per-env stores + commit", "// ---- termination"), ("stage pair words", "// ---- reward (SDE"),
           ("sweep / box dispatch", "const unsigned TR = rep2"), ("alignment + epilogue + outputs", "// the running-episode record is fetched"),
           ("reset kernel", "fused_reset_kernel(")]
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, per_file = None, {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try: ln, inst, samp = int(r[0]), int(r[7]), int(r[6])
        except ValueError: continue
        d = per_file.setdefault(cur, {}).setdefault(ln, [r[1], 0, 0]); d[1] += inst; d[2] += samp
ti = sum(d[1] for f in per_file.values() for d in f.values()) or 1
ts = sum(d[2] for f in per_file.values() for d in f.values()) or 1
f = per_file.get("fused_step.cuh", {})
# markers are looked up in the source file as it was when the report was captured (argv[2]; default: working tree)
import os
path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym-swarm_b200", "csrc", "fused_step.cuh")
text = open(path).read().splitlines()
starts = []
for name, mark in MARKERS:
    ls = [i + 1 for i, l in enumerate(text) if mark in l]
    if ls: starts.append((min(ls), name))
starts.sort()
print(f"# {sys.argv[1]}: regions of fused_step.cuh (share of executed warp instructions, of stall samples)")
for i, (ln, name) in enumerate(starts):
    hi = starts[i + 1][0] if i + 1 < len(starts) else 10**9
    a = sum(d[1] for l, d in f.items() if ln <= l < hi); b = sum(d[2] for l, d in f.items() if ln <= l < hi)
    print(f"  {name:28s} lines {ln:4d}-{min(hi, max(f)) :4d}  inst={a / ti * 100:5.1f}%  samples={b / ts * 100:5.1f}%")
for fn, d in per_file.items():
    if fn != "fused_step.cuh":
        print(f"  [{fn}] inst={sum(x[1] for x in d.values()) / ti * 100:5.1f}%  samples={sum(x[2] for x in d.values()) / ts * 100:5.1f}%")
