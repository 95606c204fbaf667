#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_cfg3.csv            > profiles/r01_cfg3_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof_cfg3_fused.ncu-rep      > profiles/r01_cfg3_fused_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
] + [f"smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio" for r in (
    "long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "no_instruction", "not_selected",
    "lg_throttle", "mio_throttle", "branch_resolving", "dispatch_stall", "drain", "membar", "sleeping", "tex_throttle",
    "imc_miss")]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {path}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    for k, a in agg.items():
        print(f"{k:92s} n={a[0]:3d} total_us={a[1] / 1e3:10.1f} avg_us={a[1] / a[0] / 1e3:9.1f} share={a[1] / tot:.3f}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on : {path}")
    ki = hdr.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        print(f"## launch {n}: {r[ki][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:74s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    data, cur = {}, None
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] not in ("", "Line No"):
            try:
                ln, inst, samp = int(r[0]), int(r[7]), int(r[6])
            except ValueError:
                continue
            d = data.setdefault((cur, ln), [r[1].strip()[:100], 0, 0])
            d[1] += inst
            d[2] += samp
    ti, ts = sum(d[1] for d in data.values()) or 1, sum(d[2] for d in data.values()) or 1
    print("## hottest source lines (share of executed warp instructions / of stall samples, all captured launches)")
    for k, d in sorted(data.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"  {k[0]:18s}:{k[1]:4d} inst={d[1] / ti * 100:5.1f}% samples={d[2] / ts * 100:5.1f}%  {d[0]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
