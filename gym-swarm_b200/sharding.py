"""Multi-GPU plumbing: environments are independent (no cross-env term anywhere in
``swarm_discrete_env.py``), so the path shards by giving every rank a contiguous slice of the
environments and of every tape.  There is no data-path collective; the only exchange is the
optional sum all-reduce of the 8 episode statistics.

Two ways to do that all-reduce:
  * ``BatchedSwarmEnv.stats_allreduce()`` -- the library's own NCCL communicator
    (``swarm_nccl_init`` / ``swarm_stats_allreduce`` in include/swarm_abi.h), GPU only;
  * ``allreduce_stats()`` below -- ``torch.distributed`` on whatever backend the job initialised
    (``nccl`` on the GPU box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np

STAT_KEYS = ("episodes", "sum_length", "sum_return", "sum_survival", "sum_attraction", "sum_repulsion",
             "sum_alignment", "running_steps")


def shard_range(num_envs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the environments owned by ``rank``; the first ``num_envs % world_size``
    ranks own one extra environment.  An environment is never split across ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(num_envs), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slice_reset_tapes(rt, lo: int, hi: int):
    """ResetTapes (tapes.py) restricted to envs [lo, hi): the env axis is axis 1."""
    return replace(rt, place=np.ascontiguousarray(rt.place[:, lo:hi]), ptape=np.ascontiguousarray(rt.ptape[:, lo:hi]),
                   orient=np.ascontiguousarray(rt.orient[:, lo:hi]))


def slice_step_tapes(st, lo: int, hi: int):
    """StepTapes restricted to envs [lo, hi): the env axis is axis 1 (axis 0 is the step)."""
    return replace(st, actions=np.ascontiguousarray(st.actions[:, lo:hi]), retarget=np.ascontiguousarray(st.retarget[:, lo:hi]),
                   slab=np.ascontiguousarray(st.slab[:, lo:hi]), roll=np.ascontiguousarray(st.roll[:, lo:hi]))


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of a CUDA device's PCIe root (from sysfs), or -1 when unknown."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return -1


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the GPU's NUMA node, so that the pinned host buffers it allocates next
    (first touch) are local to that GPU's PCIe root.  With one process per GPU on a two-socket box this keeps half of the
    host <-> device traffic off the socket interconnect.  Returns what was done (for the bench record); never raises."""
    import os
    node = gpu_numa_node(device_index)
    info = {"numa_node": node, "bound": False}
    if node < 0:
        return info
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:                       # containers may hide sysfs or forbid the call: run unbound
        info["error"] = type(exc).__name__
    return info


def allreduce_stats(stats: dict, group=None, device=None) -> dict:
    """Sum the episode statistics over all ranks with torch.distributed (fp64, one 8-element all-reduce)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    if device is None and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_KEYS, [float(v) for v in t.cpu()]))


class ShardedSwarmEnv:
    """This rank's slice of a job-wide batch of ``num_envs_total`` environments.

    ``env`` is a ``BatchedSwarmEnv`` over ``hi - lo`` environments on this rank's GPU; tapes given in global
    (all-env) shape are sliced before they are bound.  ``global_stats()`` all-reduces the episode statistics.
    """

    def __init__(self, num_envs_total, num_agents, obs_space_size, rank=None, world_size=None, use_library_nccl=False,
                 **env_kwargs):
        import torch.distributed as dist

        from .envs import BatchedSwarmEnv
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.rank, self.world_size = rank, world_size
        self.num_envs_total = int(num_envs_total)
        self.lo, self.hi = shard_range(num_envs_total, world_size, rank)
        env_kwargs.setdefault("track_stats", True)
        # env_offset = lo: host tape generators differ per shard, and the in-kernel Philox streams are keyed by the
        # GLOBAL env index, so a sharded device_tapes job draws exactly what one big batch would draw
        env_kwargs.setdefault("env_offset", self.lo)
        self.env = BatchedSwarmEnv(self.hi - self.lo, num_agents, obs_space_size, **env_kwargs)
        self._library_nccl = bool(use_library_nccl) and world_size > 1
        if self._library_nccl:
            def bcast(b):
                obj = [b]
                dist.broadcast_object_list(obj, src=0)
                return obj[0]
            self.env.init_nccl(rank, world_size, bcast)

    def set_reset_tapes_global(self, rt):
        s = slice_reset_tapes(rt, self.lo, self.hi)
        self.env.set_reset_tapes(s.place, s.ptape)

    def reset(self):
        return self.env.reset()

    def step(self, actions, reward_type=None, retarget=None, slab=None):
        """``actions`` / ``retarget`` / ``slab`` are this rank's slice (use ``lo:hi`` of the global arrays)."""
        return self.env.step(actions, reward_type, retarget=retarget, slab=slab)

    def global_stats(self):
        if self._library_nccl:
            return self.env.stats_allreduce()
        return allreduce_stats(self.env.stats())

    def close(self):
        self.env.close()
