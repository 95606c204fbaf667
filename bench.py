#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched swarm transition on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--config cfg3] [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one ``swarm_step`` (reference ``DiscreteSwarmEnv.step``, swarm_discrete_env.py:223-267) over
every environment this rank owns, on synthetic tapes of the named shape.  The default workload is
BASELINE.json's headline configuration (configs[2]: 65,536 envs x 128 agents on a 128x128 grid, per-agent
reward dict + reward curriculum); environments are independent, so each rank owns a full copy of the
workload ("scaling": "weak"; ``--scaling strong`` shards the 65,536 envs over the ranks instead).

One JSON line on rank 0:
  value      whole-job agent-steps/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the public API with HOST (pinned) inputs and outputs: per step the H2D copy of
             actions / retarget flags / re-draw slab and the D2H read of observation, reward(s) and done
  roofline   the dominant kernel's algorithmic HBM bytes / its live CUDA-event duration / measured HBM peak
  roofline_int32  the same kernel's algorithmic INT32 lane-ops against the INT32-pipe peak measured in this run
  cpu_baseline    the C port of the reference (oracle/swarm_oracle.c, OpenMP) on this box's host cores
``--impl reference`` times only that CPU port (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (E, N, S, per_agent, curriculum, auto_reset, track_stats, description)
CONFIGS = {
    "cfg2": dict(E=4096, N=16, S=32, per_agent=False, curriculum=False,
                 desc="configs[1]: 4,096 envs x 16 agents, 32x32 grid, global reward"),
    "cfg3": dict(E=65536, N=128, S=128, per_agent=True, curriculum=True,
                 desc="configs[2]: 65,536 envs x 128 agents, 128x128 grid, per-agent reward dict + curriculum"),
    "cfg4": dict(E=16, N=16384, S=1024, per_agent=True, curriculum=False,
                 desc="configs[3]: 16 envs x 16,384 agents, 1024x1024 grid, pairwise dominated"),
    "cfg5": dict(E=1048576, N=8, S=16, per_agent=False, curriculum=False, place_extra=64,
                 desc="configs[4]: 1,048,576 envs x 8 agents, 16x16 grid, auto-reset + episode statistics"),
    # streaming shape for an honest DRAM number: 16 M envs x 8 agents
    "stream": dict(E=16777216, N=8, S=16, per_agent=False, curriculum=False, place_extra=64,
                   desc="streaming shape: 16,777,216 envs x 8 agents, 16x16 grid (>= 1 GB of traffic per launch)"),
}
METRIC = "agent-steps/sec"
UNIT = "agent-steps/s"


def curriculum_weights(t, on):
    """reward_type weights at step t (w_a = 1, w_r ramps 0 -> 1 over 500 steps, w_l = 0.5)."""
    if not on:
        return {"attraction": True, "repulsion": True, "alignment": True, "indiv_rewards": True, "vf_size": None}
    return {"attraction": 1.0, "repulsion": min(1.0, t / 500.0), "alignment": 0.5, "indiv_rewards": True,
            "vf_size": None}


def algorithmic_bytes(E, N, per_agent):
    """9 B per agent-step (x,y in+out int16, action u8), 34 B per env-step (predator state r/w,
    retarget flag, done, 4 x f32 global reward), + 16 B per agent-step with per-agent rewards (4 x f32)."""
    return E * N * (9 + (16 if per_agent else 0)) + E * 34


def algorithmic_int_ops(E, N):
    """14 INT32 lane-ops per ordered pair for the reward + 2 for the collision check."""
    return E * N * (N - 1) * 16


# ------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        inside = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        for ln in inside:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the C port of the reference (oracle/swarm_oracle.c), OpenMP over environments
# ------------------------------------------------------------------------------------------------------
def cpu_port_run(cfg, sample_envs, steps, warmup, seed=123):
    """Returns (agent-steps/s, seconds per step, threads).  Test infrastructure used as a *measured baseline* only."""
    from gym_swarm_b200 import tapes
    from oracle import c_oracle
    E, N, S = sample_envs, cfg["N"], cfg["S"]
    threads = c_oracle.set_threads(host_cores())        # torchrun exports OMP_NUM_THREADS=1: override it
    rng = np.random.default_rng(seed)
    K = tapes.default_slab_len(N)
    rt = tapes.make_reset_tapes(rng, 2, E, N, S)
    T = max(1, min(4, steps))
    st = tapes.make_step_tapes(rng, T, E, N, K=K)
    ora = c_oracle.COracleBatch(E, N, S, 5, 2, -10.0, auto_reset=True, K=K)
    ora.set_reset_tapes(rt.place, rt.ptape)
    ora.reset()
    for t in range(warmup):
        w = curriculum_weights(t, cfg["curriculum"])
        ora.step(st.actions[t % T], st.slab[t % T], st.retarget[t % T],
                 (float(w["attraction"]), float(w["repulsion"]), float(w["alignment"])), None, cfg["per_agent"],
                 want_state=False)
    t0 = time.perf_counter()
    for t in range(steps):
        w = curriculum_weights(t, cfg["curriculum"])
        ora.step(st.actions[t % T], st.slab[t % T], st.retarget[t % T],
                 (float(w["attraction"]), float(w["repulsion"]), float(w["alignment"])), None, cfg["per_agent"],
                 want_state=False)
    dt = time.perf_counter() - t0
    ora.close()
    return E * N * steps / dt, dt / steps, threads


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_envs(cfg, target_s=1.0):
    """Envs per CPU step so that one step is roughly target_s of wall time on all cores (bounded sample)."""
    N = cfg["N"]
    per_env = 3.0e-9 * N * N * 2 + 2e-7 * N          # literal O(N^2) collision + reward loops
    cores = host_cores()
    e = int(target_s * cores / per_env)
    return int(max(cores, min(cfg["E"], e)))


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    E_s = cpu_sample_envs(cfg, target_s=0.5)
    v, spp, threads = cpu_port_run(cfg, E_s, args.steps, args.warmup)
    sample = f"{E_s} envs x {cfg['N']} agents per step (bounded sample of {cfg['E']} envs), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "envs": cfg["E"], "agents": cfg["N"], "grid": cfg["S"],
                   "reward": "per-agent dict + curriculum" if cfg["per_agent"] else "global", "auto_reset": True},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "what": "oracle/swarm_oracle.c: plain-C restatement of swarm_discrete_env.py (the Python "
                                 "reference cannot travel to the GPU box), OpenMP over environments"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--envs", type=int, default=None, help="override the number of envs per rank")
    ap.add_argument("--tape-steps", type=int, default=8, help="distinct pre-drawn tape steps rotated through")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--path", default="auto", choices=["auto", "fused", "group", "staged", "staged_pairwise"])
    ap.add_argument("--vf", type=int, default=None, help="reward_type['vf_size'] (default None)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="time the HBM-resident loop as CUDA-graph replays of --tape-steps steps (auto: when a step has "
                         "fewer than 1M agents and the host launch cost would dominate)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist

    from gym_swarm_b200 import _lib, tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    E_total = cfg["E"] if args.envs is None else args.envs
    if args.scaling == "strong" and world > 1:
        E = E_total // world                              # contiguous slice [rank*E, (rank+1)*E)
    else:
        E = E_total
    N, S = cfg["N"], cfg["S"]
    per_agent = cfg["per_agent"]
    T = args.tape_steps

    # ---- synthetic tapes, seeded per (config, rank) -------------------------------------
    rng = np.random.default_rng(sorted(CONFIGS).index(args.config) * 1000 + rank)
    env = BatchedSwarmEnv(E, N, S, device=f"cuda:{local_rank}", auto_reset=True, track_stats=True, reset_slots=4,
                          per_agent=per_agent, debug_outputs=False, path=args.path,
                          place_extra=cfg.get("place_extra"))
    K = env.K
    place = torch.from_numpy(rng.integers(0, S, size=(env.R, E, env.PL), dtype=np.int16))
    ptape = torch.from_numpy(rng.integers(0, S, size=(env.R, E, env.KP, 2), dtype=np.int16))
    env.set_reset_tapes(place, ptape)
    del place, ptape
    h_actions = torch.from_numpy(rng.integers(0, 8, size=(T, E, N), dtype=np.uint8)).pin_memory()
    h_retarget = torch.from_numpy((rng.random(size=(T, E)) < 0.1).astype(np.uint8)).pin_memory()
    h_slab = torch.from_numpy(rng.integers(0, 8, size=(T, E, K), dtype=np.uint8)).pin_memory()
    d_actions, d_retarget, d_slab = h_actions.to(dev), h_retarget.to(dev), h_slab.to(dev)
    env.reset()
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_no = [0]

    def dev_step():
        t = step_no[0]
        k = t % T
        env.step(d_actions[k], curriculum_weights(t, cfg["curriculum"]) if args.vf is None else
                 dict(curriculum_weights(t, cfg["curriculum"]), vf_size=args.vf),
                 retarget=d_retarget[k], slab=d_slab[k])
        step_no[0] = t + 1
        return env.last_launch_count()

    use_graph = args.graph == "on" or (args.graph == "auto" and E * N < (1 << 20) and args.vf is None)
    roll = None
    if use_graph:                                        # T steps per graph launch, tapes rotate inside the graph
        roll = env.capture_rollout(d_actions, d_retarget, d_slab,
                                   [curriculum_weights(t, cfg["curriculum"]) for t in range(T)])
        graph_launches_per_step = roll.launches / T
        n_replays = max(1, -(-args.steps // T))
        args.steps = n_replays * T

    # ---- warm-up + device-resident timed region ---------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        dev_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches = 0
    w0 = time.time()
    ev0.record()
    if roll is not None:
        for _ in range(n_replays):
            roll.replay()
        launches = int(round(graph_launches_per_step * args.steps))
        step_no[0] += args.steps
    else:
        for _ in range(args.steps):
            launches += dev_step()
    ev1.record()
    barrier()
    w1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    agent_steps = float(E) * N * args.steps * world
    value = agent_steps / (ms * 1e-3)
    env.check_errors()

    # ---- dominant-kernel duration, live CUDA events around single launches ----------------------------
    kt = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev_step()
        b.record()
        kt.append((a, b))
    torch.cuda.synchronize(dev)
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    alg_b = algorithmic_bytes(E, N, per_agent)
    ach = alg_b / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath) and args.envs is None and args.scaling == "weak":
        key = f"{args.config}:{env.path_name}" if args.vf is None else f"{args.config}:vf{args.vf}:{env.path_name}"
        ent = json.load(open(tpath)).get(key)
        if ent:
            traffic, traffic_src = ent["traffic_bytes"], ent["source"]
    roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": env.path_name, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_b,
                "peak_source": peak_src,
                "note": "whole transition fused into one kernel for N <= 256; at this shape it is INT32/DPX-pipe "
                        "bound (pairwise reward), see roofline_int32" if env.path_name.startswith("fused") else
                        "staged path: duration is the sum of the step's kernels"}
    roofline_int32 = None
    if rank == 0:
        import ctypes as C
        lib = _lib.load()
        pk = {}
        for kind, name in ((0, "iadd3"), (3, "viaddmnmx_s16x2"), (4, "viaddmnmx_s16x2+imad")):
            out = C.c_double()
            _lib.check(lib.swarm_int32_peak(local_rank, kind, C.byref(out),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            pk[name] = out.value
        ops = algorithmic_int_ops(E, N)
        ach_i = ops / (k_ms * 1e-3)
        roofline_int32 = {"bound": "int32-pipe", "achieved": ach_i / 1e12, "peak": pk["iadd3"] / 1e12,
                          "unit": "T lane-ops/s", "frac": ach_i / pk["iadd3"], "algorithmic_ops_per_launch": ops,
                          "measured_peaks_T": {k: v / 1e12 for k, v in pk.items()},
                          "peak_source": "swarm_int32_peak micro-benchmark in this run (independent IADD3 chains)"}

    # ---- end to end through the public API with host buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        n_e2e = args.e2e_steps or max(5, min(args.steps, 40))
        h_x = torch.empty((E, N), dtype=torch.int16).pin_memory()
        h_y = torch.empty((E, N), dtype=torch.int16).pin_memory()
        h_rew = torch.empty((E, 5), dtype=torch.float32).pin_memory()
        h_done = torch.empty((E,), dtype=torch.uint8).pin_memory()
        h_agent = torch.empty((E, N, 4), dtype=torch.float32).pin_memory() if per_agent else None
        d_a, d_r, d_s = torch.empty_like(d_actions[0]), torch.empty_like(d_retarget[0]), torch.empty_like(d_slab[0])

        def host_step():
            t = step_no[0]
            k = t % T
            d_a.copy_(h_actions[k], non_blocking=True)          # H2D from pinned host memory
            d_r.copy_(h_retarget[k], non_blocking=True)
            d_s.copy_(h_slab[k], non_blocking=True)
            (x, y), reward, done, _ = env.step(d_a, curriculum_weights(t, cfg["curriculum"]), retarget=d_r, slab=d_s)
            h_x.copy_(x, non_blocking=True)                     # D2H: observation, rewards, done
            h_y.copy_(y, non_blocking=True)
            h_rew.copy_(reward["global"], non_blocking=True)
            h_done.copy_(done, non_blocking=True)
            if per_agent:
                h_agent.copy_(reward["agents"], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()        # the caller reads the result before the next step
            step_no[0] = t + 1

        for _ in range(3):
            host_step()
        barrier()
        ev0.record()
        for _ in range(n_e2e):
            host_step()
        ev1.record()
        barrier()
        e_ms = max_over_ranks(ev0.elapsed_time(ev1))
        h2d = d_a.numel() + d_r.numel() + d_s.numel()
        d2h = (h_x.numel() + h_y.numel()) * 2 + h_rew.numel() * 4 + h_done.numel() + \
              (h_agent.numel() * 4 if per_agent else 0)
        e2e = {"value": float(E) * N * n_e2e * world / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "ms_per_step": e_ms / n_e2e,
               "what": "BatchedSwarmEnv.step with pinned-host actions/retarget/slab in, observation + global reward "
                       "+ done" + (" + per-agent rewards" if per_agent else "") + " out, stream sync every step"}

    # ---- episode statistics: NCCL all-reduce (the only collective of the path), outside the timed region -
    stats = env.stats()
    if world > 1:
        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        env.init_nccl(rank, world, bcast)
        stats = env.stats_allreduce()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        E_s = cpu_sample_envs(cfg, target_s=0.5)
        v, spp, threads = cpu_port_run(cfg, E_s, 4, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{E_s} envs x {N} agents x 4 steps of the same workload ({spp * 4:.1f} s wall, "
                         f"{threads} threads)",
               "what": "oracle/swarm_oracle.c (C restatement of the Python reference), OpenMP over environments"}

    if rank == 0:
        ws_mb = (E * N * (4 + 1 + (16 if per_agent else 0)) + E * (K + 40)) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "envs_per_gpu": E, "envs_total": E * world, "agents": N, "grid": S,
                       "reward": "per-agent dict + curriculum" if per_agent else "global", "auto_reset": True,
                       "kernel_path": env.path_name, "launches_per_step": env.last_launch_count(),
                       "l2": f"no flush: per-step working set {ws_mb:.0f} MB, {T} rotating tape steps "
                             f"({ws_mb * 1.0 + T * (E * N + E * K) / 1e6:.0f} MB touched per rotation) vs 126 MB L2",
                       "parallelism": f"env-sharded x{world}" if world > 1 else "single GPU",
                       "timed_loop": (f"CUDA graph replays of {T} steps" if roll is not None else "eager swarm_step calls")},
            "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "roofline_int32": roofline_int32, "cpu_baseline": cpu, "episode_stats": stats,
        }
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
