#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched swarm transition on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--config cfg3] [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one ``swarm_step`` (reference ``DiscreteSwarmEnv.step``, swarm_discrete_env.py:223-267) over
every environment this rank owns, on synthetic tapes of the named shape.  The headline workload is
BASELINE.json's configs[2] (65,536 envs x 128 agents on a 128x128 grid, per-agent reward dict + reward
curriculum); environments are independent, so each rank owns a full copy of the workload ("scaling": "weak";
``--scaling strong`` shards the 65,536 envs over the ranks instead).

One JSON line on rank 0:
  value      whole-job agent-steps/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: per step ONE packed H2D copy of
             actions / retarget flags / re-draw slab and ONE packed D2H copy of observation, done, global reward and
             per-agent rewards (BatchedSwarmEnv.step_host / fetch_host), stream sync every step
  roofline   the dominant kernel's algorithmic HBM bytes / its live CUDA-event duration / measured HBM peak
  roofline_int32  frac_executed = the kernel's ALU-pipe utilisation from its ncu capture (profiles/ncu_pipe.json);
             scalar_equivalent = SURVEY.md's 16 scalar ops per ordered pair against the IADD3 peak measured in this run
             (not a utilisation: the kernels execute packed 16x2 ops and visit unordered pairs once)
  configs    every OTHER named BASELINE config, timed in this same invocation (each < 1 s of GPU time):
             cfg2 (CUDA-graph replay), cfg3_strong (65,536 / N envs per rank), cfg4 (16 / N envs per rank),
             cfg5 (1,048,576 / N envs per rank, auto-reset, episode statistics reduced -- and NCCL all-reduced
             when N > 1 -- every 64 steps INSIDE the timed loop, asynchronously), cfg3_e2e_terms (the e2e loop with
             the compact int16 per-agent credit instead of the f32 rewards)
  cpu_baseline    the UNMODIFIED Python reference (baseline/_ref, one process per host core) on a bounded sample,
             with the C port (oracle/swarm_oracle.c, OpenMP) beside it
``--impl reference`` times only the CPU side (the Python reference; the port is reported inside ``cpu_baseline``).
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

CONFIGS = {
    "cfg2": dict(E=4096, N=16, S=32, per_agent=False, curriculum=False,
                 desc="configs[1]: 4,096 envs x 16 agents, 32x32 grid, global reward"),
    "cfg3": dict(E=65536, N=128, S=128, per_agent=True, curriculum=True,
                 desc="configs[2]: 65,536 envs x 128 agents, 128x128 grid, per-agent reward dict + curriculum"),
    "cfg4": dict(E=16, N=16384, S=1024, per_agent=True, curriculum=False,
                 desc="configs[3]: 16 envs x 16,384 agents, 1024x1024 grid, pairwise dominated"),
    "cfg5": dict(E=1048576, N=8, S=16, per_agent=False, curriculum=False, place_extra=64,
                 desc="configs[4]: 1,048,576 envs x 8 agents, 16x16 grid, auto-reset + episode statistics"),
    # not a BASELINE config: the 8-agents-per-lane shape of the lane-group kernels (N = 129 .. 256)
    "n256": dict(E=32768, N=256, S=128, per_agent=True, curriculum=True,
                 desc="32,768 envs x 256 agents, 128x128 grid, per-agent reward dict + curriculum (N = 129..256 kernel shape)"),
    # streaming shape for an honest DRAM number: 16 M envs x 8 agents
    "stream": dict(E=16777216, N=8, S=16, per_agent=False, curriculum=False, place_extra=64,
                   desc="streaming shape: 16,777,216 envs x 8 agents, 16x16 grid (>= 1 GB of traffic per launch)"),
}
METRIC = "agent-steps/sec"
UNIT = "agent-steps/s"
STATS_EVERY = 64


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


def config_dict(cfg, scaling, world):
    """The workload, identically in both arms (GPU and --impl reference): a pure function of the command line."""
    total = cfg["E"] * (world if scaling == "weak" else 1)
    ws_mb = (cfg["E"] * cfg["N"] * (4 + 1 + (16 if cfg["per_agent"] else 0)) + cfg["E"] * 72) / 1e6
    return {"workload": cfg["desc"], "envs_total": total, "agents": cfg["N"], "grid": cfg["S"],
            "reward": "per-agent dict + curriculum" if cfg["per_agent"] else "global", "auto_reset": True,
            "parallelism": f"env-sharded x{world}" if world > 1 else "single GPU",
            "l2": f"no flush: one step touches {ws_mb:.0f} MB per GPU (state + per-agent outputs + rotating tapes) vs the "
                  f"126 MB L2" if ws_mb > 126 else f"no flush: the named config's working set is {ws_mb:.1f} MB per step "
                  f"(L2 resident by nature)"}


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
# CPU arms.  (1) the UNMODIFIED Python reference, one process per core (oracle/reference_runner.py);
#            (2) the C port of the reference (oracle/swarm_oracle.c), OpenMP over environments
# ------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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


def cpu_sample_envs(cfg, target_s=1.0):
    """Envs per CPU-port step so that one step is roughly target_s of wall time on all cores (bounded sample)."""
    N = cfg["N"]
    per_env = 3.0e-9 * N * N * 2 + 2e-7 * N          # literal O(N^2) collision + reward loops
    cores = host_cores()
    e = int(target_s * cores / per_env)
    return int(max(cores, min(cfg["E"], e)))


def port_baseline(cfg, steps=4, warmup=1, target_s=0.5):
    E_s = cpu_sample_envs(cfg, target_s=target_s)
    v, spp, threads = cpu_port_run(cfg, E_s, steps, warmup)
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{E_s} envs x {cfg['N']} agents x {steps} steps of the same workload ({spp * steps:.1f} s wall, "
                      f"{threads} threads)",
            "what": "oracle/swarm_oracle.c (plain-C restatement of swarm_discrete_env.py), OpenMP over environments"}


def reference_sample(cfg, cores, budget_s):
    """(envs, steps) of the Python reference's bounded sample: SURVEY.md 8d asks for 64 envs x 100 steps (1 env x 1000
    steps at configs[0]); at N = 128 one env-step costs ~40 ms of one core, so the step count is cut to the budget."""
    N = cfg["N"]
    per_env_step = 6.0e-4 + 2.2e-6 * N * N             # measured: 0.6 ms at N = 4, 36 ms at N = 128
    envs = max(cores, 64)
    steps = int(max(3, min(100, budget_s * cores / (envs * per_env_step))))
    return envs, steps


def reference_baseline(cfg, steps=None, warmup=1, budget_s=15.0):
    """The unmodified Python reference on every host core; None when it is not available or not feasible."""
    from oracle import reference_runner, reference_shim
    if not reference_shim.reference_available():
        return None, "baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)"
    if cfg["N"] > 2048:
        return None, (f"the reference's O(N^2) Python loops and N x N float64 temporaries make N = {cfg['N']} infeasible "
                      f"(SURVEY.md 6: minutes per call, ~8 x 2.1 GB)")
    cores = host_cores()
    envs, auto_steps = reference_sample(cfg, cores, budget_s)
    steps = auto_steps if steps is None else steps
    rt = curriculum_weights(250, cfg["curriculum"]) if cfg["per_agent"] else None
    r = reference_runner.time_reference(cfg["N"], cfg["S"], envs, steps, warmup=warmup, reward_type=rt, workers=cores)
    rec = {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": "reference",
           "per_core": r["per_core_mean"], "seconds_per_step": r["seconds_per_step"],
           "sample": f"{envs} envs x {cfg['N']} agents x {steps} steps (+{warmup} warm-up) of the same shape, random actions, "
                     f"reset on done; {r['workers']} worker processes, slowest worker {r['seconds_per_step'] * steps:.1f} s",
           "what": "UNMODIFIED gym_swarm/envs/swarm_discrete_env.py (pip --target baseline/_ref copy) under import stubs for "
                   "gym / matplotlib, raw NumPy + scikit-learn, process-global np.random; one process per core"}
    return rec, None


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    sample_steps = args.steps
    ref, why = reference_baseline(cfg, steps=sample_steps, warmup=max(1, min(args.warmup, 3)))
    port = port_baseline(cfg, steps=4, warmup=1)
    if ref is not None:
        v, spp, cpu = ref["value"], ref["seconds_per_step"], dict(ref, port=port)
    else:                                               # the reference cannot run this shape: the port stands in
        v, cpu = port["value"], dict(port, reference_unavailable=why)
        spp = None
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None if spp is None else spp * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int64/f64 (NumPy)", "data": "synthetic",
        "config": config_dict(cfg, args.scaling, args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class Job:
    """Process-group plumbing shared by the headline run and the sub-records."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def bcast_bytes(self, b):
        obj = [b]
        self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def timed(self, fn, n):
        """n calls of fn between barriers, CUDA events on the launching stream; max over ranks, in ms."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1))


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def profile_entry(fname, key):
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.isfile(path):
        return None
    return json.load(open(path)).get(key)


def device_tapes(job, env, T, seed):
    """Synthetic tapes drawn on the device (sub-records: a host draw of [T, E, K] bytes costs seconds at 1 M envs)."""
    torch = job.torch
    g = torch.Generator(device=job.dev)
    g.manual_seed(seed)
    E, N, S = env.E, env.N, env.S
    place = torch.randint(0, S, (env.R, E, env.PL), dtype=torch.int16, device=job.dev, generator=g)
    ptape = torch.randint(0, S, (env.R, E, env.KP, 2), dtype=torch.int16, device=job.dev, generator=g)
    env.set_reset_tapes(place, ptape)
    actions = torch.randint(0, 8, (T, E, N), dtype=torch.uint8, device=job.dev, generator=g)
    retarget = (torch.rand((T, E), device=job.dev, generator=g) < 0.1).to(torch.uint8)
    slab = torch.randint(0, 8, (T, E, env.K), dtype=torch.uint8, device=job.dev, generator=g)
    return actions, retarget, slab


def run_sub(job, name, steps, warmup, strong=True, graph=False, stats_every=0, T=8):
    """One named BASELINE config as a sub-record: value / ms_per_step / kernel_path / roofline fraction."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    torch = job.torch
    cfg = CONFIGS[name]
    world = job.world
    if strong:
        if cfg["E"] % world:
            return {"skipped": f"{cfg['E']} envs do not split over {world} ranks"}
        E = cfg["E"] // world
    else:
        E = cfg["E"]
    N, S = cfg["N"], cfg["S"]
    env = BatchedSwarmEnv(E, N, S, device=f"cuda:{job.local_rank}", auto_reset=True, track_stats=True, reset_slots=4,
                          per_agent=cfg["per_agent"], debug_outputs=False, place_extra=cfg.get("place_extra"),
                          env_offset=job.rank * E if strong else 0)
    actions, retarget, slab = device_tapes(job, env, T, 7000 + 13 * job.rank + sorted(CONFIGS).index(name))
    env.reset()
    torch.cuda.synchronize(job.dev)
    if stats_every and world > 1:
        env.init_nccl(job.rank, world, job.bcast_bytes)
    step_no = [0]
    pending = [0]

    def one_step(with_stats):
        t = step_no[0]
        k = t % T
        env.step(actions[k], curriculum_weights(t, cfg["curriculum"]), retarget=retarget[k], slab=slab[k])
        step_no[0] = t + 1
        if with_stats and stats_every and (t + 1) % stats_every == 0:
            if pending[0] >= 3:                       # collect the oldest result: long finished, never blocks in practice
                env.stats_end(wait=True)
                pending[0] -= 1
            env.stats_begin(allreduce=world > 1)      # reduction on this stream; NCCL + D2H on the side stream
            pending[0] += 1

    rec = {"workload": cfg["desc"], "envs_per_gpu": E, "envs_total": E * world,
           "scaling": "strong" if strong else "weak (one replica per GPU)", "kernel_path": env.path_name}
    roll = None
    if graph:
        roll = env.capture_rollout(actions, retarget, slab, [curriculum_weights(t, cfg["curriculum"]) for t in range(T)])
        n_replays = max(1, -(-steps // T))
        rps = max(1, stats_every // T) if stats_every else 0      # replays per statistics interval
        if stats_every:
            n_replays = max(2 * rps, -(-n_replays // rps) * rps)
        steps = n_replays * T
        for _ in range(max(3, -(-warmup // T))):
            roll.replay()
        ms_plain = job.timed(roll.replay, n_replays)
        ms = ms_plain
        rec["timed_loop"] = f"CUDA graph replays of {T} steps"
        rec["gpu_launches"] = int(roll.launches * n_replays)
        if stats_every:
            done_replays = [0]

            def replay_with_stats():
                roll.replay()
                done_replays[0] += 1
                if done_replays[0] % rps == 0:
                    if pending[0] >= 3:                       # the oldest result: long finished, never blocks in practice
                        env.stats_end(wait=True)
                        pending[0] -= 1
                    env.stats_begin(allreduce=world > 1)
                    pending[0] += 1

            ms = job.timed(replay_with_stats, n_replays)
            results = []
            while pending[0] > 0:
                results.append(env.stats_end(wait=True))
                pending[0] -= 1
            rec["timed_loop"] = (f"CUDA graph replays of {T} steps (per-GPU batches of this config are launch bound in eager "
                                 f"mode from 4 GPUs up); every {stats_every} steps swarm_stats_begin (multi-CTA reduction on the "
                                 f"step stream, {'NCCL all-reduce of 8 doubles on the device buffer + ' if world > 1 else ''}"
                                 f"D2H on the library's side stream), results collected 3 intervals later, no host sync")
            rec["ms_per_step_without_stats"] = ms_plain / steps
            rec["stats_overhead_frac"] = ms / ms_plain - 1.0
            rec["stats_reductions_in_timed_loop"] = steps // stats_every
            rec["episode_stats" + ("_allreduced" if world > 1 else "")] = results[-1] if results else None
            rec["gpu_launches"] += 2 * (steps // stats_every)
    else:
        if stats_every:
            steps = max(2 * stats_every, -(-steps // stats_every) * stats_every)
        for _ in range(max(3, warmup)):
            one_step(False)
        ms_plain = job.timed(lambda: one_step(False), steps)
        ms = ms_plain
        rec["timed_loop"] = "eager swarm_step calls"
        rec["gpu_launches"] = int(env.last_launch_count() * steps)
        if stats_every:
            step_no[0] = 0
            ms = job.timed(lambda: one_step(True), steps)
            results = []
            while pending[0] > 0:
                results.append(env.stats_end(wait=True))
                pending[0] -= 1
            rec["timed_loop"] = (f"eager swarm_step calls; every {stats_every} steps swarm_stats_begin (multi-CTA reduction on "
                                 f"the step stream, {'NCCL all-reduce of 8 doubles on the device buffer + ' if world > 1 else ''}"
                                 f"D2H on the library's side stream), results collected 3 intervals later, no host sync")
            rec["ms_per_step_without_stats"] = ms_plain / steps
            rec["stats_overhead_frac"] = ms / ms_plain - 1.0
            rec["stats_reductions_in_timed_loop"] = steps // stats_every
            rec["episode_stats" + ("_allreduced" if world > 1 else "")] = results[-1] if results else None
            rec["gpu_launches"] += 2 * (steps // stats_every)
    env.check_errors()
    total_envs = E * world
    value = float(total_envs) * N * steps / (ms * 1e-3)
    peak, _ = hbm_peak()
    alg = algorithmic_bytes(E, N, cfg["per_agent"])
    rec.update(value=value, unit=UNIT, steps=steps, ms_per_step=ms / steps,
               roofline={"bound": "hbm", "achieved": alg / (ms / steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / steps * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": alg},
               l2=config_dict(cfg, "strong", world)["l2"])
    env.close()
    del env, actions, retarget, slab, roll
    torch.cuda.empty_cache()
    return rec


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
    ap.add_argument("--no-subconfigs", action="store_true", help="skip the sub-records of the other named configs")
    ap.add_argument("--path", default="auto", choices=["auto", "fused", "group", "staged", "staged_pairwise"])
    ap.add_argument("--vf", type=int, default=None, help="reward_type['vf_size'] (default None)")
    ap.add_argument("--flags", type=int, default=0, help="swarm_config.flags (kernel A/B switches, include/swarm_abi.h)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="time the HBM-resident loop as CUDA-graph replays of --tape-steps steps (auto: when a step has "
                         "fewer than 1M agents and the host launch cost would dominate)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])

    if args.impl == "reference":
        run_reference(args, cfg, int(os.environ.get("RANK", "0")))
        return

    job = Job()
    torch, dist = job.torch, job.dist
    rank, local_rank, world, dev = job.rank, job.local_rank, job.world, job.dev

    from gym_swarm_b200 import _lib, sharding, tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv

    # pinned host buffers are allocated by this process below: make them local to this GPU's PCIe root first
    numa = sharding.bind_to_gpu_numa_node(local_rank)

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
                          place_extra=cfg.get("place_extra"), flags=args.flags,
                          env_offset=rank * E if (args.scaling == "strong" and world > 1) else 0)
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

    step_no = [0]

    def rtype(t):
        w = curriculum_weights(t, cfg["curriculum"])
        return w if args.vf is None else dict(w, vf_size=args.vf)

    def dev_step():
        t = step_no[0]
        k = t % T
        env.step(d_actions[k], rtype(t), retarget=d_retarget[k], slab=d_slab[k])
        step_no[0] = t + 1
        return env.last_launch_count()

    use_graph = args.graph == "on" or (args.graph == "auto" and E * N < (1 << 20) and args.vf is None)
    roll = None
    if use_graph:                                        # T steps per graph launch, tapes rotate inside the graph
        roll = env.capture_rollout(d_actions, d_retarget, d_slab, [rtype(t) for t in range(T)])
        graph_launches_per_step = roll.launches / T
        n_replays = max(1, -(-args.steps // T))
        args.steps = n_replays * T

    # ---- warm-up + device-resident timed region ---------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        dev_step()
    job.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.barrier()
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
    job.barrier()
    w1 = time.time()
    ms = job.max_over_ranks(ev0.elapsed_time(ev1))
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
    peak, peak_src = hbm_peak()
    alg_b = algorithmic_bytes(E, N, per_agent)
    ach = alg_b / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    prof_key = f"{args.config}:{env.path_name}" if args.vf is None else f"{args.config}:vf{args.vf}:{env.path_name}"
    if args.envs is None and args.scaling == "weak":
        ent = profile_entry("ncu_traffic.json", prof_key)
        if ent:
            traffic, traffic_src = ent["traffic_bytes"], ent["source"]
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": env.path_name, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_b, "peak_source": peak_src,
                "note": "whole transition fused into one kernel for N <= 256; at this shape it is INT32/DPX-pipe and "
                        "latency bound (pairwise reward), not HBM bound: see roofline_int32.frac_executed"
                        if env.path_name.startswith("fused") else "staged path: duration is the sum of the step's kernels"}
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
        pipe = profile_entry("ncu_pipe.json", prof_key) or {}
        roofline_int32 = {
            "bound": "int32-pipe", "unit": "fraction of the ALU pipe's issue slots",
            "frac_executed": None if "alu_pipe_pct" not in pipe else pipe["alu_pipe_pct"] / 100.0,
            "frac_executed_source": pipe.get("source"),
            "issue_active": None if "issue_active_pct" not in pipe else pipe["issue_active_pct"] / 100.0,
            "contract_kernel": profile_entry("ncu_pipe.json", "pairwise_sym_kernel:16x16384"),
            "scalar_equivalent": {"what": "SURVEY.md 8d's 16 scalar INT32 lane-ops per ordered pair / live kernel time / "
                                          "IADD3 peak: NOT a utilisation (packed 16x2 DPX ops, unordered pairs visited "
                                          "once, far thresholds from a border pass)",
                                  "ops_per_launch": ops, "T_lane_ops_per_s": ops / (k_ms * 1e-3) / 1e12,
                                  "ratio_to_iadd3_peak": ops / (k_ms * 1e-3) / pk["iadd3"]},
            "measured_peaks_T": {k: v / 1e12 for k, v in pk.items()},
            "peak_source": "swarm_int32_peak micro-benchmark in this run (independent dependent-chain kernels)"}

    # ---- end to end through the public API with host buffers: ONE packed copy each way per step -----------
    e2e = None
    if not args.no_e2e:
        n_e2e = args.e2e_steps or max(5, min(args.steps, 40))
        for k in range(T):                                  # T pinned, packed input blocks: actions | retarget | slab
            hin = env.host_inputs(slot=k)
            hin["actions"].copy_(h_actions[k]); hin["retarget"].copy_(h_retarget[k]); hin["slab"].copy_(h_slab[k])

        def host_step():
            t = step_no[0]
            env.step_host(rtype(t), slot=t % T)             # 1 H2D of this step's inputs + the step's kernels
            env.fetch_host(sync=True)                       # 1 D2H: x, y, done, reward[, per-agent rewards]; then sync:
            step_no[0] = t + 1                              # the caller reads the result before the next step

        for _ in range(3):
            host_step()
        e_ms = job.timed(host_step, n_e2e)
        h2d, d2h = env.input_bytes, env.fetch_bytes
        e2e = {"value": float(E) * N * n_e2e * world / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "ms_per_step": e_ms / n_e2e,
               "pcie_gb_s_per_rank": (h2d + d2h) / (e_ms / n_e2e * 1e-3) / 1e9, "numa": numa,
               "what": "BatchedSwarmEnv.step_host + fetch_host: one packed pinned H2D (actions | retarget | slab) and one "
                       "packed D2H (observation x, y | done | global reward" + (" | per-agent f32 rewards" if per_agent else "")
                       + ") per step, stream sync every step"}

    stats = env.stats()
    env.close()
    del env, d_actions, d_retarget, d_slab
    torch.cuda.empty_cache()

    # ---- the other named configs, as sub-records of the same invocation -------------------------------------
    configs = None
    if not args.no_subconfigs and args.config == "cfg3" and args.envs is None and args.vf is None:
        configs = {}
        configs["cfg2"] = run_sub(job, "cfg2", steps=800, warmup=40, strong=False, graph=True)
        configs["cfg3_strong"] = run_sub(job, "cfg3", steps=100, warmup=10, strong=True)
        configs["cfg4"] = run_sub(job, "cfg4", steps=96, warmup=16, strong=True, graph=True)
        # (from 4 ranks up the per-GPU batch of cfg5 is launch bound in eager mode: CUDA-graph replays there; the replays also
        # record rewards / dones per step, which costs ~8 % where the step is GPU bound, so large batches stay eager)
        configs["cfg5"] = run_sub(job, "cfg5", steps=3 * STATS_EVERY, warmup=10, strong=True, stats_every=STATS_EVERY,
                                  graph=CONFIGS["cfg5"]["E"] // job.world <= 262144)
        if not args.no_e2e:
            configs["cfg3_e2e_terms"] = run_e2e_terms(job, cfg, h_actions, h_retarget, h_slab, T, numa)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        port = port_baseline(cfg)
        ref, why = reference_baseline(cfg)
        cpu = dict(ref, port=port) if ref is not None else dict(port, reference_unavailable=why)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int32/int16x2 (f64 reward combine)", "data": "synthetic",
            "config": config_dict(cfg, args.scaling, world),
            "run": {"envs_per_gpu": E, "kernel_path": prof_key.split(":", 1)[1], "launches_per_step": launches / max(1, args.steps),
                    "tape_steps_rotated": T,
                    "timed_loop": (f"CUDA graph replays of {T} steps" if roll is not None else "eager swarm_step calls")},
            "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "roofline_int32": roofline_int32, "cpu_baseline": cpu, "episode_stats": stats, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e_terms(job, cfg, h_actions, h_retarget, h_slab, T, numa):
    """The headline e2e loop with the compact per-agent credit (int16 x 4 per agent: half the D2H bytes)."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    torch = job.torch
    E, N, S = cfg["E"], cfg["N"], cfg["S"]
    env = BatchedSwarmEnv(E, N, S, device=f"cuda:{job.local_rank}", auto_reset=True, track_stats=True, reset_slots=4,
                          per_agent=False, agent_terms=True, debug_outputs=False)
    device_tapes(job, env, 1, 99 + job.rank)
    env.reset()
    for k in range(T):
        hin = env.host_inputs(slot=k)
        hin["actions"].copy_(h_actions[k]); hin["retarget"].copy_(h_retarget[k]); hin["slab"].copy_(h_slab[k])
    step_no = [0]

    def host_step():
        t = step_no[0]
        env.step_host(curriculum_weights(t, cfg["curriculum"]), slot=t % T)
        env.fetch_host(sync=True)
        step_no[0] = t + 1

    for _ in range(3):
        host_step()
    n = 20
    e_ms = job.timed(host_step, n)
    h = env.fetch_host(sync=True)
    t0 = time.perf_counter()
    sample = h["agent_terms"][:4096]
    env.decode_agent_terms(sample, curriculum_weights(0, cfg["curriculum"]))
    dec = (time.perf_counter() - t0) / sample.shape[0] * E
    rec = {"value": float(E) * N * n * job.world / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms / n, "steps": n,
           "h2d_bytes_per_step": env.input_bytes, "d2h_bytes_per_step": env.fetch_bytes,
           "pcie_gb_s_per_rank": (env.input_bytes + env.fetch_bytes) / (e_ms / n * 1e-3) / 1e9, "numa": numa,
           "decode_ms_per_step_one_core": dec * 1e3,
           "what": "as e2e, but the per-agent credit crosses PCIe as swarm_step_out.agent_terms (int16 x 4: rew_rep_i, "
                   "rew_attr_i, U_i, Q_i) instead of 4 x f32; a consumer that needs the floats forms them with three "
                   "multiplies per agent (BatchedSwarmEnv.decode_agent_terms; decode_ms_per_step_one_core = torch on one "
                   "host core, NOT inside the timed region)"}
    env.close()
    del env
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    main()
