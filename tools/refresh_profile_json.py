#!/usr/bin/env python
"""profiles/ncu_traffic.json and profiles/ncu_pipe.json from the raw ncu csv files of tools/gpu_final_profiles_r02.sh
(gpurun_out/r02_*_raw.csv): DRAM bytes and pipe utilisation of the dominant kernel(s) of every bench config."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def rows_of(tag):
    path = os.path.join(OUT, f"r02_{tag}_raw.csv")
    if not os.path.isfile(path):
        return None, None
    rows = list(csv.reader(open(path)))
    return rows[0], rows[2:]


def val(hdr, r, key, default=0.0):
    if key not in hdr:
        return default
    try:
        return float(r[hdr.index(key)].replace(",", ""))
    except ValueError:
        return default


def unit_scale(hdr, units_row, key):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units_row[hdr.index(key)], 1.0)


def summarise(tag):
    path = os.path.join(OUT, f"r02_{tag}_raw.csv")
    if not os.path.isfile(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    out = []
    for r in body:
        rd = val(hdr, r, "dram__bytes_read.sum") * unit_scale(hdr, units, "dram__bytes_read.sum")
        wr = val(hdr, r, "dram__bytes_write.sum") * unit_scale(hdr, units, "dram__bytes_write.sum")
        dur = val(hdr, r, "gpu__time_duration.sum")
        dur_us = dur * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
        out.append(dict(kernel=r[hdr.index("Kernel Name")][:80], read=rd, write=wr, us=dur_us,
                        alu=val(hdr, r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
                        fma=val(hdr, r, "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
                        xu=val(hdr, r, "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"),
                        lsu=val(hdr, r, "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
                        issue=val(hdr, r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        warps=val(hdr, r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                        inst=val(hdr, r, "smsp__inst_executed.sum"), regs=val(hdr, r, "launch__registers_per_thread"),
                        conflicts=val(hdr, r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                        wavefronts=val(hdr, r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")))
    return out


def main():
    traffic = json.load(open(os.path.join(PROF, "ncu_traffic.json")))
    pipe = json.load(open(os.path.join(PROF, "ncu_pipe.json")))
    for tag, key in (("cfg3_fused", "cfg3:fused<32,4>+bitmap+box (vf: x2 launches)"), ("cfg3_vf5", "cfg3:vf5:fused<32,4>+bitmap+box (vf: x2 launches)"),
                     ("cfg5_tiny", "cfg5:tiny<8> (vf: fused<8,1>)"), ("cfg4_staged", "cfg4:staged+box (vf: staged pairwise)")):
        ks = summarise(tag)
        if not ks:
            continue
        # a capture may hold several launches of the SAME kernel (several steps of a single-launch path): one per name
        seen = {}
        for k in ks:
            seen.setdefault(k["kernel"], k)
        ks = list(seen.values())
        rd, wr, us = sum(k["read"] for k in ks), sum(k["write"] for k in ks), sum(k["us"] for k in ks)
        names = " + ".join(k["kernel"].split("(")[0].replace("void ", "") for k in ks)
        traffic[key] = {"traffic_bytes": int(rd + wr),
                        "source": f"profiles/r02_{tag}_full.txt ({rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB write; {names})"}
        top = max(ks, key=lambda k: k["us"])            # the longest kernel of the step carries the pipe numbers
        tot_inst = sum(k["inst"] for k in ks)
        pipe[key] = {"alu_pipe_pct": round(top["alu"], 2), "fma_pipe_pct": round(top["fma"], 2), "xu_pipe_pct": round(top["xu"], 2),
                     "lsu_pipe_pct": round(top["lsu"], 2), "issue_active_pct": round(top["issue"], 2),
                     "warps_active_pct": round(top["warps"], 2), "registers": int(top["regs"]),
                     "kernel": top["kernel"], "duration_us": round(top["us"], 1),
                     "step_kernels": [{"kernel": k["kernel"], "us": round(k["us"], 1), "inst": int(k["inst"]),
                                       "alu_pipe_pct": round(k["alu"], 2), "issue_active_pct": round(k["issue"], 2)} for k in ks],
                     "inst_executed_per_step": int(tot_inst),
                     "smem_bank_conflict_share": round(top["conflicts"] / top["wavefronts"], 4) if top["wavefronts"] else None,
                     "source": f"profiles/r02_{tag}_full.txt"}
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    json.dump(pipe, open(os.path.join(PROF, "ncu_pipe.json"), "w"), indent=1)
    print("refreshed", sorted(traffic), sorted(pipe))


if __name__ == "__main__":
    main()
