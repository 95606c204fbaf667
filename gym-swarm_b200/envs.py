"""Host-side mirror of the reference Gym surface for the swarm transition.

``BatchedSwarmEnv``  -- E independent environments stepped together on one B200; tensor I/O.
``SwarmEnv``         -- drop-in for the reference ``DiscreteSwarmEnv`` (one env, dict I/O):
                        same constructor defaults, ``reset()``, ``step(action[, reward_type])``,
                        ``set_env_parameters``, ``set_reward_parameters`` and public attributes as
                        ``/root/reference/gym_swarm/envs/swarm_discrete_env.py:168-406`` ("SDE").
``ContinuousSwarmEnv`` -- alias: the reference's swarm_continuous_env.py is byte-identical to the
                        discrete file except for the class name.

All compute happens in ``libswarm_b200.so`` (sm_100a CUDA) through ctypes; PyTorch only provides
device memory and the current stream.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from . import tapes as _tapes

REWARD_KEYS = ("survival", "attraction", "repulsion", "alignment", "sum")
DEFAULT_REWARD_TYPE = {"attraction": True, "repulsion": True, "alignment": True,
                       "indiv_rewards": False, "vf_size": None}           # SDE:223-227

# SDE:467-475 / 477-485
action_to_move = {0: np.array([-1, 0]), 1: np.array([-1, -1]), 2: np.array([0, -1]), 3: np.array([1, -1]),
                  4: np.array([1, 0]), 5: np.array([1, 1]), 6: np.array([0, 1]), 7: np.array([-1, 1]),
                  8: np.array([0, 0])}
ACTION_LOOKUP = {0: "left", 1: "left-down", 2: "down", 3: "right-down", 4: "right", 5: "right-up", 6: "up",
                 7: "left-up", 8: "no-move"}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _as_dev(a, dtype, device):
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=dtype)
    else:
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device)
    return t.contiguous()


class BatchedSwarmEnv:
    """E environments, N agents + 1 predator each, on an S x S grid, on one GPU.

    Randomness is explicit (tapes.py): reset tapes are bound once, the per-step ``retarget`` flags
    and re-draw ``slab`` either come from the caller (parity runs) or from an internal host
    generator that pre-draws chunks.
    """

    def __init__(self, num_envs, num_agents=4, obs_space_size=20, attraction_thresh=5, repulsion_thresh=2,
                 predator_eat_rew=-10, device="cuda:0", auto_reset=True, track_stats=False, slab_len=None,
                 place_extra=None, pred_attempts=8, reset_slots=4, per_agent=True, debug_outputs=False,
                 path="auto", seed=0, tape_chunk=16, device_tapes=False, env_offset=0, flags=0, tiny16_min_envs=0,
                 agent_terms=False):
        self.lib = _lib.load()                      # raises when the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.SwarmLibraryError("no CUDA device: the swarm transition has no CPU fallback")
        self.device = torch.device(device)
        self.E, self.N, self.S = int(num_envs), int(num_agents), int(obs_space_size)
        self.attraction_thresh, self.repulsion_thresh = attraction_thresh, repulsion_thresh
        self.predator_eat_rew = predator_eat_rew
        self.auto_reset, self.track_stats = bool(auto_reset), bool(track_stats)
        self.K = _tapes.default_slab_len(self.N) if slab_len is None else int(slab_len)
        self.KR = _tapes.default_place_extra(self.N) if place_extra is None else int(place_extra)
        self.PL = 2 * self.N + 2 * self.KR
        self.KP, self.R = int(pred_attempts), int(reset_slots)
        self.per_agent, self.debug_outputs = bool(per_agent), bool(debug_outputs)
        self.with_terms = bool(agent_terms)             # compact per-agent credit int16[E,N,4] (swarm_step_out.agent_terms)
        # env_offset = global index of this batch's env 0 (a shard's ``lo``): host tapes are drawn from a generator
        # keyed by (seed, env_offset) so that shards do not replay each other, and the in-kernel Philox counters use
        # env + env_offset so that a sharded device_tapes run draws exactly what the unsharded run draws
        self.env_offset = int(env_offset)
        self._rng = np.random.default_rng(seed if self.env_offset == 0 else [int(seed), self.env_offset])
        # device_tapes: every random draw of the transition is computed inside the kernels by Philox4x32-10 keyed with
        # ``seed`` (no tape memory, no host draws); tapes.philox_* materialise the identical values on the host
        self.device_tapes = bool(device_tapes)
        self.seed = int(seed)
        self.step_index = 0
        self._tape_chunk = int(tape_chunk)
        self._chunk = None
        self._chunk_pos = 0

        cfg = _lib.Config(
            abi_version=_lib.ABI_VERSION, device=self.device.index or 0, num_envs=self.E, num_agents=self.N,
            grid_size=self.S, attraction_thresh=int(math.ceil(attraction_thresh)),
            repulsion_thresh=int(math.ceil(repulsion_thresh)), predator_eat_rew=float(predator_eat_rew),
            slab_len=self.K, place_len=self.PL, pred_attempts=self.KP, reset_slots=self.R,
            auto_reset=int(self.auto_reset), track_stats=int(self.track_stats),
            path={"auto": _lib.PATH_AUTO, "staged": _lib.PATH_STAGED, "fused": _lib.PATH_FUSED,
                  "group": _lib.PATH_GROUP, "staged_pairwise": _lib.PATH_STAGED_PAIRWISE}[path],
            flags=int(flags), tiny16_min_envs=int(tiny16_min_envs), env_offset=self.env_offset & 0xFFFFFFFF, reserved=0)
        self._h = C.c_void_p()
        _lib.check(self.lib.swarm_create(C.byref(cfg), C.byref(self._h)))
        self.path_name = self.lib.swarm_path_name(self._h).decode()

        E, N, dev = self.E, self.N, self.device
        # Everything a caller reads back per step -- the observation (x, y), done, the global reward and the per-agent
        # credit -- is carved out of ONE contiguous device block ("result arena"), so that a host-side consumer fetches
        # a step's results with a single device-to-host copy (fetch_host) instead of one per tensor.
        arena_spec = [("x", (E, N), torch.int16), ("y", (E, N), torch.int16), ("done", (E,), torch.uint8),
                      ("reward", (E, 5), torch.float32)]
        if self.per_agent:
            arena_spec.append(("agent_reward", (E, N, 4), torch.float32))
        if self.with_terms:
            arena_spec.append(("agent_terms", (E, N, 4), torch.int16))
        self._arena_layout, off = [], 0
        for name, shape, dt in arena_spec:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._arena_layout.append((name, shape, dt, off, nbytes))
            off = (off + nbytes + 255) // 256 * 256
        self.arena = torch.zeros(off, dtype=torch.uint8, device=dev)
        self.agent_reward = self.agent_terms = None
        for name, shape, dt, o, nbytes in self._arena_layout:
            setattr(self, name, self.arena[o:o + nbytes].view(dt).view(shape))
        self._host_arena = None
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        # persistent state (SoA, int16 positions; x, y live in the arena)
        self.px, self.py = z((E,), torch.int16), z((E,), torch.int16)
        self.target = z((E,), torch.int32)
        self.pred_orient, self.frozen = z((E,), torch.uint8), z((E,), torch.uint8)
        self.err, self.reset_count = z((E,), torch.int32), z((E,), torch.int32)
        # cell of the predator's target: state of the staged kernels only (include/swarm_abi.h, swarm_state.tcell)
        self.tcell = z((E, 2), torch.int16) if self.path_name.startswith("staged") else None
        if self.track_stats:
            # one 16-byte running-episode record per env: word 0 = length, words 1..3 = f32 bits of the running
            # attraction / repulsion / alignment return (``ep_len`` / ``ep_ret`` below are views of it)
            self.ep_rec = z((E, 4), torch.int32)
            self.fin_count, self.fin_len = z((E,), torch.int32), z((E,), torch.int32)
            self.fin_ret = z((E, 4), torch.float32)
        else:
            self.ep_rec = self.fin_count = self.fin_len = self.fin_ret = None
        st = _lib.State(x=_ptr(self.x), y=_ptr(self.y), px=_ptr(self.px), py=_ptr(self.py), target=_ptr(self.target),
                        pred_orient=_ptr(self.pred_orient), frozen=_ptr(self.frozen), err=_ptr(self.err),
                        reset_count=_ptr(self.reset_count), tcell=_ptr(self.tcell), ep_rec=_ptr(self.ep_rec),
                        fin_count=_ptr(self.fin_count), fin_len=_ptr(self.fin_len), fin_ret=_ptr(self.fin_ret))
        _lib.check(self.lib.swarm_bind_state(self._h, C.byref(st)), self._h)
        # remaining step outputs (owned by the env, overwritten by the next step)
        self.eaten = z((E,), torch.int32)
        self.counts = z((E, 2), torch.int32)
        self.pred_prev, self.pred_next = z((E, 2), torch.int16), z((E, 2), torch.int16)
        self.step_orient, self.step_target = z((E,), torch.uint8), z((E,), torch.int32)
        self.consumed = z((E,), torch.int32)
        self.term_x, self.term_y = z((E, N), torch.int16), z((E, N), torch.int16)
        self.agent_counts = z((E, N, 2), torch.int32) if self.debug_outputs else None
        self.eff_action = z((E, N), torch.uint8) if self.debug_outputs else None
        self._out = _lib.StepOut(
            done=_ptr(self.done), eaten=_ptr(self.eaten), reward=_ptr(self.reward), counts=_ptr(self.counts),
            pred_prev=_ptr(self.pred_prev), pred_next=_ptr(self.pred_next), step_orient=_ptr(self.step_orient),
            step_target=_ptr(self.step_target), consumed=_ptr(self.consumed), term_x=_ptr(self.term_x),
            term_y=_ptr(self.term_y), agent_reward=_ptr(self.agent_reward), agent_counts=_ptr(self.agent_counts),
            eff_action=_ptr(self.eff_action), agent_terms=_ptr(self.agent_terms))
        self._place = self._ptape = None
        self._nccl_ready = False

    # ---- lifecycle -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.swarm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- tapes -------------------------------------------------------------------------------------
    def set_reset_tapes(self, place, ptape):
        """place int16[R,E,2N+2KR], ptape int16[R,E,KP,2] (tapes.make_reset_tapes)."""
        place = _as_dev(place, torch.int16, self.device)
        ptape = _as_dev(ptape, torch.int16, self.device)
        assert tuple(place.shape) == (self.R, self.E, self.PL), (tuple(place.shape), (self.R, self.E, self.PL))
        assert tuple(ptape.shape) == (self.R, self.E, self.KP, 2), tuple(ptape.shape)
        self._place, self._ptape = place, ptape
        self._dev_reset_bound = False
        _lib.check(self.lib.swarm_bind_reset_tapes(self._h, _ptr(place), _ptr(ptape)), self._h)

    def _ensure_reset_tapes(self):
        """Reset tapes for callers that bind none: the in-kernel counter streams (Philox keyed by ``seed``, stream index =
        the env's reset count).  They never wrap, so an env that resets for the n-th time gets a fresh placement -- a
        materialised tape with ``reset_slots`` slots would cycle through ``reset_slots`` placements per env -- and the
        only bounds left are the attempts per reset (``place_extra`` re-draws, ``pred_attempts``).  Explicit tapes
        (``set_reset_tapes``: parity runs, the reference-compatible single env) take precedence."""
        if self._place is None and not getattr(self, "_dev_reset_bound", False):
            _lib.check(self.lib.swarm_set_device_reset_tapes(self._h, C.c_uint64(self.reset_seed)), self._h)
            self._dev_reset_bound = True

    def _next_step_tapes(self):
        """Tapes for callers that do not bring their own: drawn in chunks of ``tape_chunk`` steps.  Small batches draw
        on the host (NumPy generator, reproducible from ``seed`` alone); large ones draw on the device with a torch
        generator seeded from the same NumPy stream -- a host draw of [T, E, K] bytes costs seconds at a million envs."""
        if self._chunk is None or self._chunk_pos >= self._tape_chunk:
            T = self._tape_chunk
            if self.E * self.K * T <= (1 << 24):
                roll = self._rng.random(size=(T, self.E))
                retarget = torch.from_numpy((roll < 0.1).astype(np.uint8)).to(self.device)
                slab = torch.from_numpy(self._rng.integers(0, 8, size=(T, self.E, self.K), dtype=np.uint8)).to(self.device)
            else:
                g = torch.Generator(device=self.device)
                g.manual_seed(int(self._rng.integers(0, 2 ** 62)))
                retarget = (torch.rand((T, self.E), device=self.device, generator=g) < 0.1).to(torch.uint8)
                slab = torch.randint(0, 8, (T, self.E, self.K), dtype=torch.uint8, device=self.device, generator=g)
            self._chunk, self._chunk_pos = (retarget, slab), 0
        r, s = self._chunk[0][self._chunk_pos], self._chunk[1][self._chunk_pos]
        self._chunk_pos += 1
        return r, s

    @property
    def reset_seed(self):
        """Key of the in-kernel reset streams (placement, predator attempts) in device_tapes mode."""
        return (self.seed * 0x9E3779B97F4A7C15 + 1) & 0xFFFFFFFFFFFFFFFF

    # ---- Gym-style surface (tensor I/O) ----------------------------------------------------------
    def reset(self):
        """SDE:197-221 for every env. Returns the observation (x, y): int16[E,N] views of the state."""
        self._ensure_reset_tapes()
        _lib.check(self.lib.swarm_reset(self._h, self._stream()), self._h)
        return self.x, self.y

    def step(self, actions, reward_type=None, retarget=None, slab=None):
        """SDE:223-267 for every env.

        actions: uint8[E,N] (ids 0..8).  Returns (obs, reward, done, info):
          obs    = (x, y) int16[E,N] (post auto-reset where done)
          reward = {"global": f32[E,5] (survival, attraction, repulsion, alignment, sum),
                    "agents": f32[E,N,4] (attraction, repulsion, alignment, sum) or None,
                    "counts": i32[E,2] integer rew_rep / rew_attr}
          done   = uint8[E]
          info   = predator_state / predator_next_state int16[E,2], predator_orientation, eaten, ...
        """
        rt = DEFAULT_REWARD_TYPE if reward_type is None else reward_type
        vf = rt.get("vf_size", None)
        w_alignment = float(rt["alignment"])
        if vf is not None and vf < 0:                   # SDE:355 `unalign[dist_1 > vf_size] = 0` then clears every
            vf, w_alignment = None, 0.0                 # pair (D >= 0 > vf): the alignment term is identically zero
        actions = _as_dev(actions, torch.uint8, self.device)
        assert tuple(actions.shape) == (self.E, self.N), tuple(actions.shape)
        if self.device_tapes and retarget is None and slab is None:
            retarget = slab = None                      # drawn in the kernels: Philox(seed, step_index)
        else:
            if retarget is None or slab is None:
                r2, s2 = self._next_step_tapes()
                retarget = r2 if retarget is None else retarget
                slab = s2 if slab is None else slab
            retarget = _as_dev(retarget, torch.uint8, self.device)
            slab = _as_dev(slab, torch.uint8, self.device)
            assert tuple(retarget.shape) == (self.E,) and tuple(slab.shape) == (self.E, self.K), (retarget.shape, slab.shape)
        sin = _lib.StepIn(actions=_ptr(actions), retarget=_ptr(retarget), slab=_ptr(slab),
                          w_attraction=float(rt["attraction"]), w_repulsion=float(rt["repulsion"]),
                          w_alignment=w_alignment, vf_size=-1 if vf is None else int(math.floor(vf)),
                          rng_step=self.step_index & 0xFFFFFFFF, rng_seed=self.seed & 0xFFFFFFFFFFFFFFFF)
        self.step_index += 1
        self._keep = (actions, retarget, slab)      # keep inputs alive until the kernels ran
        _lib.check(self.lib.swarm_step(self._h, C.byref(sin), C.byref(self._out), self._stream()), self._h)
        reward = {"global": self.reward, "agents": self.agent_reward, "counts": self.counts, "terms": self.agent_terms}
        info = {"predator_state": self.pred_prev, "predator_next_state": self.pred_next,
                "predator_orientation": self.step_orient, "predator_target": self.step_target, "eaten": self.eaten,
                "terminal_state": (self.term_x, self.term_y), "redraws": self.consumed, "errors": self.err}
        return (self.x, self.y), reward, self.done, info

    # ---- host-side consumers: one packed copy per step ------------------------------------------------------
    def fetch_host(self, sync=True):
        """The step's results on the host with ONE device-to-host copy of the result arena into a pinned mirror
        (allocated on first use, by the calling thread: bind the process to the GPU's NUMA node first --
        ``sharding.bind_to_gpu_numa_node`` -- so that the pages are local to the GPU's PCIe root).  Returns
        ``{"x", "y", "done", "reward"[, "agent_reward"][, "agent_terms"]}`` as CPU tensors viewing the mirror; they are
        valid once the stream has been synchronised (``sync=True`` does that) and are overwritten by the next fetch."""
        if self._host_arena is None:
            self._host_arena = torch.empty(self.arena.numel(), dtype=torch.uint8).pin_memory()
            self._host_views = {name: self._host_arena[o:o + nbytes].view(dt).view(shape)
                                for name, shape, dt, o, nbytes in self._arena_layout}
        self._host_arena.copy_(self.arena, non_blocking=True)          # a single cudaMemcpyAsync
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return self._host_views

    def host_inputs(self, slot=0):
        """Pinned host views ``{"actions" u8[E,N], "retarget" u8[E], "slab" u8[E,K]}`` of ONE packed input block (only
        ``actions`` in device_tapes mode): fill them, then ``step_host(slot=...)`` moves the block with a single
        host-to-device copy.  ``slot`` selects one of several independent blocks (a producer can fill the next block
        while the previous step runs)."""
        if getattr(self, "_in_host", None) is None:
            spec = [("actions", (self.E, self.N))]
            if not self.device_tapes:
                spec += [("retarget", (self.E,)), ("slab", (self.E, self.K))]
            self._in_layout, off = [], 0
            for name, shape in spec:
                nb = int(np.prod(shape))
                self._in_layout.append((name, shape, off, nb))
                off = (off + nb + 255) // 256 * 256
            self._in_bytes = off
            self._in_host, self._in_host_views = {}, {}
            self._in_dev = torch.zeros(off, dtype=torch.uint8, device=self.device)
            self._in_dev_views = {n: self._in_dev[o:o + nb].view(sh) for n, sh, o, nb in self._in_layout}
        if slot not in self._in_host:
            blk = torch.zeros(self._in_bytes, dtype=torch.uint8).pin_memory()
            self._in_host[slot] = blk
            self._in_host_views[slot] = {n: blk[o:o + nb].view(sh) for n, sh, o, nb in self._in_layout}
        return self._in_host_views[slot]

    def step_host(self, reward_type=None, slot=0):
        """``step`` on the inputs currently in ``host_inputs(slot)``: one packed H2D copy + the step's kernels."""
        self.host_inputs(slot)
        self._in_dev.copy_(self._in_host[slot], non_blocking=True)     # a single cudaMemcpyAsync
        d = self._in_dev_views
        return self.step(d["actions"], reward_type, retarget=d.get("retarget"), slab=d.get("slab"))

    @property
    def input_bytes(self):
        self.host_inputs()
        return int(self._in_bytes)

    @property
    def fetch_bytes(self):
        """Bytes one fetch_host moves (the arena, 256-byte padding between its tensors included)."""
        return int(self.arena.numel())

    def decode_agent_terms(self, terms, reward_type=None):
        """float64[...,4] (attraction, repulsion, alignment, sum) from the compact per-agent credit (int16[...,4] =
        rew_rep_i, rew_attr_i, U_i, Q_i; include/swarm_abi.h) for the weights of ``reward_type`` -- the same values
        ``agent_reward`` holds, formed on the consumer's side (SDE:371-379)."""
        rt = DEFAULT_REWARD_TYPE if reward_type is None else reward_type
        t = terms.to(torch.float64) if isinstance(terms, torch.Tensor) else torch.as_tensor(np.asarray(terms), dtype=torch.float64)
        inv = 1.0 / (self.N * (self.N - 1))
        w_ali = float(rt["alignment"])
        vf = rt.get("vf_size", None)
        if vf is not None and vf < 0:
            w_ali = 0.0
        rep = float(rt["repulsion"]) * t[..., 0] * inv
        att = float(rt["attraction"]) * t[..., 1] * inv
        ali = -w_ali * (0.25 * t[..., 2] - 0.35355339059327376220 * t[..., 3]) * inv
        return torch.stack([att, rep, ali, att + rep + ali], -1)

    def observe_visual_field(self, vf_size=2, out=None):
        """Egocentric occupancy windows of the current state: uint8[E,N,2vf+1,2vf+1], [dy+vf][dx+vf]; 0 empty / outside
        the grid, 1 self, 2 other agent, 3 predator (README figure "Visual Field 5x5"; include/swarm_abi.h)."""
        W = 2 * int(vf_size) + 1
        if out is None:
            out = torch.empty((self.E, self.N, W, W), dtype=torch.uint8, device=self.device)
        assert tuple(out.shape) == (self.E, self.N, W, W) and out.dtype == torch.uint8 and out.is_contiguous()
        _lib.check(self.lib.swarm_observe_vf(self.E, self.N, self.S, int(vf_size), _ptr(self.x), _ptr(self.y), _ptr(self.px),
                                             _ptr(self.py), _ptr(out), self._stream()))
        return out

    def last_launch_count(self):
        return int(self.lib.swarm_last_launch_count(self._h))

    def check_errors(self):
        """Raise if any env has a sticky error bit other than STEP_WHEN_DONE."""
        bad = (self.err & ~_lib.ERR_STEP_WHEN_DONE).nonzero()
        if bad.numel():
            e = int(bad[0])
            raise _lib.SwarmLibraryError(f"env {e} error bits {int(self.err[e])} (tape exhausted or bad action)")

    # ---- statistics -------------------------------------------------------------------------------
    STAT_KEYS = ("episodes", "sum_length", "sum_return", "sum_survival", "sum_attraction", "sum_repulsion",
                 "sum_alignment", "running_steps")

    @property
    def ep_len(self):
        """Length of every env's running episode: int32[E] view of the running-episode record."""
        return None if self.ep_rec is None else self.ep_rec[:, 0]

    @property
    def ep_ret(self):
        """Running attraction / repulsion / alignment return: f32[E,3] view of the running-episode record."""
        return None if self.ep_rec is None else self.ep_rec.view(torch.float32)[:, 1:]

    def stats(self):
        """Episode statistics of this batch (synchronous: reduces on the current stream and waits for the result)."""
        out = (C.c_double * 8)()
        _lib.check(self.lib.swarm_stats_read(self._h, out, self._stream()), self._h)
        return dict(zip(self.STAT_KEYS, [float(v) for v in out]))

    def stats_begin(self, allreduce=False):
        """Asynchronous statistics for a step loop (e.g. every 64 steps): the reduction is enqueued on the current
        stream, the optional NCCL sum all-reduce over the ranks (on the device buffer) and the copy to pinned host
        memory run on the library's side stream.  Never blocks the host; collect with ``stats_end``."""
        _lib.check(self.lib.swarm_stats_begin(self._h, int(bool(allreduce)), self._stream()), self._h)

    def stats_end(self, wait=False):
        """Oldest result begun and not yet collected, or None while it is still in flight (``wait=False``)."""
        out = (C.c_double * 8)()
        rc = self.lib.swarm_stats_end(self._h, out, int(bool(wait)))
        if rc == _lib.E_NOT_READY:
            return None
        _lib.check(rc, self._h)
        return dict(zip(self.STAT_KEYS, [float(v) for v in out]))

    def init_nccl(self, rank, world_size, broadcast_bytes):
        """Bootstrap the library's own NCCL communicator; ``broadcast_bytes(b, src=0)`` must return
        rank 0's 128-byte unique id on every rank (e.g. via torch.distributed.broadcast_object_list)."""
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(self.lib.swarm_nccl_unique_id(uid))
        raw = broadcast_bytes(bytes(uid))
        uid = (C.c_uint8 * 128).from_buffer_copy(raw)
        _lib.check(self.lib.swarm_nccl_init(self._h, uid, world_size, rank), self._h)
        self._nccl_ready = True

    def stats_allreduce(self):
        """Sum the episode statistics over all ranks (NCCL over NVLink, 8 doubles, on the device buffer); synchronous."""
        buf = (C.c_double * 8)()
        _lib.check(self.lib.swarm_stats_allreduce(self._h, buf, self._stream()), self._h)
        return dict(zip(self.STAT_KEYS, [float(v) for v in buf]))

    # ---- checkpoint / resume (the reference's state is its attributes + the global NumPy RNG) ----
    _STATE_TENSORS = ("x", "y", "px", "py", "target", "pred_orient", "frozen", "err", "reset_count", "tcell",
                      "ep_rec", "fin_count", "fin_len", "fin_ret")

    def state_dict(self):
        """Everything needed to continue bit-identically: device state, statistics, reset tapes, the host tape
        generator and the unconsumed part of the current tape chunk."""
        torch.cuda.synchronize(self.device)
        d = {"shape": (self.E, self.N, self.S), "tensors": {}, "rng": self._rng.bit_generator.state,
             "chunk_pos": self._chunk_pos, "seed": self.seed, "step_index": self.step_index}
        for k in self._STATE_TENSORS:
            t = getattr(self, k)
            if t is not None:
                d["tensors"][k] = t.detach().cpu().clone()
        d["place"] = None if self._place is None else self._place.cpu().clone()
        d["ptape"] = None if self._ptape is None else self._ptape.cpu().clone()
        d["chunk"] = None if self._chunk is None else tuple(t.cpu().clone() for t in self._chunk)
        return d

    def load_state_dict(self, d):
        assert tuple(d["shape"]) == (self.E, self.N, self.S), "checkpoint is for another shape"
        for k, v in d["tensors"].items():
            if getattr(self, k, None) is not None:
                getattr(self, k).copy_(v)               # in place: the bound device pointers stay valid
        if self.tcell is not None:                      # (a checkpoint of another kernel path carries no target cell)
            idx = self.target.long().clamp(0, self.N - 1).unsqueeze(1)
            self.tcell[:, 0].copy_(self.x.gather(1, idx).squeeze(1))
            self.tcell[:, 1].copy_(self.y.gather(1, idx).squeeze(1))
        self._rng.bit_generator.state = d["rng"]
        self.seed, self.step_index = d.get("seed", self.seed), d.get("step_index", self.step_index)
        if d["place"] is not None:
            self.set_reset_tapes(d["place"], d["ptape"])
        else:                                           # in-kernel reset streams: re-key them with the restored seed
            self._place = self._ptape = None
            self._dev_reset_bound = False
            self._ensure_reset_tapes()
        self._chunk = None if d["chunk"] is None else tuple(t.to(self.device) for t in d["chunk"])
        self._chunk_pos = d["chunk_pos"]

    # ---- CUDA graph rollouts --------------------------------------------------------------------------
    def capture_rollout(self, actions, retarget, slab, reward_types=None):
        """Capture T consecutive steps (actions u8[T,E,N], retarget u8[T,E], slab u8[T,E,K] resident on the device) into
        one CUDA graph.  ``swarm_step`` only launches kernels, so the whole rollout replays with a single graph launch:
        this removes the per-step host cost that dominates small batches (cfg2: 4,096 x 16).  Returns a
        ``GraphedRollout``; refill the three tape tensors in place and call ``replay()``.  Per-step global rewards and
        done flags are recorded into ``rewards f32[T,E,5]`` / ``dones u8[T,E]``."""
        return GraphedRollout(self, actions, retarget, slab, reward_types)

    # ---- numpy views for tests / the compatibility class ------------------------------------------
    def state_numpy(self):
        torch.cuda.synchronize(self.device)
        return dict(x=self.x.cpu().numpy(), y=self.y.cpu().numpy(), px=self.px.cpu().numpy(), py=self.py.cpu().numpy(),
                    target=self.target.cpu().numpy(), err=self.err.cpu().numpy().astype(np.uint32),
                    frozen=self.frozen.cpu().numpy())

    def step_numpy(self, actions, slab, retarget, weights=(1.0, 1.0, 1.0), vf=None):
        rt = {"attraction": weights[0], "repulsion": weights[1], "alignment": weights[2], "vf_size": vf}
        self.step(actions, rt, retarget=retarget, slab=slab)
        o = self.state_numpy()
        o.update(done=self.done.cpu().numpy(), eaten=self.eaten.cpu().numpy(),
                 glob=self.reward.cpu().numpy().astype(np.float64), cnt=self.counts.cpu().numpy().astype(np.int64),
                 pred_prev=self.pred_prev.cpu().numpy(), pred_next=self.pred_next.cpu().numpy(),
                 pred_orient=self.step_orient.cpu().numpy(), step_target=self.step_target.cpu().numpy(),
                 consumed=self.consumed.cpu().numpy(), term_x=self.term_x.cpu().numpy(), term_y=self.term_y.cpu().numpy())
        if self.agent_reward is not None:
            ar = self.agent_reward.cpu().numpy().astype(np.float64)       # attraction, repulsion, alignment, sum
            ag = np.zeros((self.E, self.N, 5))
            ag[:, :, 1:] = ar
            dn = o["done"].astype(bool) & (o["eaten"] >= 0)
            idx = np.nonzero(dn)[0]
            ag[idx, o["eaten"][idx], 0] = float(self.predator_eat_rew)
            o["agent"] = ag
        if self.agent_terms is not None:
            o["terms"] = self.agent_terms.cpu().numpy().astype(np.int64)
        if self.agent_counts is not None:
            o["agent_cnt"] = self.agent_counts.cpu().numpy().astype(np.int64)
            o["eff"] = self.eff_action.cpu().numpy()
        return o


class GraphedRollout:
    """T steps of a ``BatchedSwarmEnv`` as one CUDA graph (see ``BatchedSwarmEnv.capture_rollout``)."""

    def __init__(self, env, actions, retarget, slab, reward_types=None):
        self.env = env
        T = actions.shape[0]
        assert actions.is_cuda and retarget.is_cuda and slab.is_cuda, "graph inputs must live on the device"
        assert actions.dtype == torch.uint8 and retarget.dtype == torch.uint8 and slab.dtype == torch.uint8
        assert tuple(actions.shape) == (T, env.E, env.N) and tuple(retarget.shape) == (T, env.E)
        assert tuple(slab.shape) == (T, env.E, env.K)
        self.T, self.actions, self.retarget, self.slab = T, actions, retarget, slab
        self.rewards = torch.zeros((T, env.E, 5), dtype=torch.float32, device=env.device)
        self.dones = torch.zeros((T, env.E), dtype=torch.uint8, device=env.device)
        rts = reward_types if isinstance(reward_types, (list, tuple)) else [reward_types] * T
        env._ensure_reset_tapes()
        self.launches = 0
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):                   # warm-up outside the capture (lazy module loading)
            snap = env.state_dict()
            env.step(actions[0], rts[0], retarget=retarget[0], slab=slab[0])
            env.load_state_dict(snap)
        torch.cuda.current_stream(env.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            for t in range(T):
                env.step(actions[t], rts[t], retarget=retarget[t], slab=slab[t])
                self.launches += env.last_launch_count()
                self.rewards[t].copy_(env.reward)
                self.dones[t].copy_(env.done)

    def replay(self):
        self.graph.replay()
        return self.rewards, self.dones


class SwarmVectorEnv:
    """Vector-env style adapter (the Gymnasium ``VectorEnv`` call shapes, no gym import needed): E copies of the swarm
    env stepped together, auto-reset on termination.

    ``reset(seed=None) -> (obs, info)``; ``step(actions) -> (obs, reward, terminated, truncated, info)`` with
    ``obs`` int16[E,N,2] (x, y of every agent), ``reward`` f32[E] (the global ``sum`` reward, SDE:366) -- the
    per-component and per-agent rewards are in ``info`` -- ``terminated`` bool[E], ``truncated`` all False (the
    reference has no time limit).  All tensors stay on the device."""

    def __init__(self, num_envs, num_agents=4, obs_space_size=20, reward_type=None, seed=0, **env_kwargs):
        env_kwargs.setdefault("auto_reset", True)
        self.env = BatchedSwarmEnv(num_envs, num_agents, obs_space_size, seed=seed, **env_kwargs)
        self.num_envs, self.num_agents, self.obs_space_size = self.env.E, self.env.N, self.env.S
        self.reward_type = dict(DEFAULT_REWARD_TYPE if reward_type is None else reward_type)
        self.single_action_space = SimpleNamespace(n=8, shape=(self.num_agents,), dtype=np.uint8)
        self.single_observation_space = SimpleNamespace(shape=(self.num_agents, 2), dtype=np.int16, low=0,
                                                        high=self.obs_space_size - 1)
        self._obs = torch.zeros((self.num_envs, self.num_agents, 2), dtype=torch.int16, device=self.env.device)

    def _observe(self):
        self._obs[:, :, 0].copy_(self.env.x)
        self._obs[:, :, 1].copy_(self.env.y)
        return self._obs

    def reset(self, seed=None):
        if seed is not None:
            env = self.env
            env._rng = np.random.default_rng(seed if env.env_offset == 0 else [int(seed), env.env_offset])
            env._place = None
            env._chunk = None
            env.seed, env.step_index = int(seed), 0     # device tapes: new Philox key, streams restart
            env._dev_reset_bound = False
        self.env.reset()
        return self._observe(), {"predator": torch.stack([self.env.px, self.env.py], 1)}

    def step(self, actions):
        _, reward, done, info = self.env.step(actions, self.reward_type)
        info = dict(info, reward_components=reward["global"], agent_rewards=reward["agents"],
                    predator=torch.stack([self.env.px, self.env.py], 1))
        terminated = done.bool()
        return self._observe(), reward["global"][:, 4], terminated, torch.zeros_like(terminated), info

    def close(self):
        self.env.close()


class SwarmEnv:
    """Drop-in for the reference ``DiscreteSwarmEnv`` (SDE:168-406) backed by one GPU environment.

    Differences (documented in INTEGRATION.md): randomness comes from explicit tapes drawn from a
    seeded ``np.random.Generator`` (``seed()``), not the process-global ``np.random``; ``render`` is
    out of scope; returned arrays are copies.
    """
    metadata = {"render.modes": ["human"]}

    def __init__(self, device="cuda:0", seed=0):
        self.num_agents = 4                          # SDE:184-195
        self.obs_space_size = 20
        self.action_space = SimpleNamespace(n=8)
        self.observation_space = np.zeros((self.obs_space_size, self.obs_space_size))
        self.random_placement = True
        self.done = None
        self.attraction_thresh = 5
        self.repulsion_thresh = 2
        self.predator_eat_rew = -10
        self._device, self._seed = device, seed
        self._rng = np.random.default_rng(seed)
        self._env = None
        self._key = None
        self.predator = None

    def seed(self, seed=None):
        self._seed = seed
        self._rng = np.random.default_rng(seed)
        return [seed]

    def set_env_parameters(self, num_agents=4, obs_space_size=20, verbose=True):   # SDE:382-392
        self.num_agents = num_agents
        self.obs_space_size = obs_space_size
        if verbose:
            print("Swarm Environment Parameters have been set to:")
            print("\t Number of Agents: {}".format(self.num_agents))
            print("\t State Space: {}x{} Grid".format(self.obs_space_size, self.obs_space_size))

    def set_reward_parameters(self, attraction_thresh=5, repulsion_thresh=2, predator_eat_rew=-10,
                              verbose=False):                                        # SDE:394-406
        self.attraction_thresh = attraction_thresh
        self.repulsion_thresh = repulsion_thresh
        self.predator_eat_rew = predator_eat_rew
        if verbose:
            print("Swarm Reward Parameters have been set to:")
            print("\t Attraction Threshold: {}".format(self.attraction_thresh))
            print("\t Repulsion Threshold: {}".format(self.repulsion_thresh))
            print("\t Predator Punishment: {}".format(self.predator_eat_rew))

    def _build(self):
        key = (self.num_agents, self.obs_space_size, self.attraction_thresh, self.repulsion_thresh,
               self.predator_eat_rew)
        if self._env is None or key != self._key:
            if self._env is not None:
                self._env.close()
            self._env = BatchedSwarmEnv(1, self.num_agents, self.obs_space_size, self.attraction_thresh,
                                        self.repulsion_thresh, self.predator_eat_rew, device=self._device,
                                        auto_reset=False, reset_slots=1, per_agent=True, debug_outputs=True)
            self._key = key

    def _state_dict(self):
        x = self._env.x[0].cpu().numpy().astype(np.int64)
        y = self._env.y[0].cpu().numpy().astype(np.int64)
        return {i: np.array([x[i], y[i]]) for i in range(self.num_agents)}

    def reset(self):                                                                 # SDE:197-221
        if not self.random_placement:
            raise NameError("name 'states_temp' is not defined")   # the reference has no fixed-placement mode
        self._build()
        env = self._env
        rt = _tapes.make_reset_tapes(self._rng, 1, 1, env.N, env.S, KR=env.KR, KP=env.KP)
        env.set_reset_tapes(rt.place, rt.ptape)
        env.reset()
        env.check_errors()
        self.current_state = self._state_dict()
        self.orientation = dict(enumerate(rt.orient[0, 0].astype(np.int64)))         # SDE:216
        self.predator = SimpleNamespace(
            obs_space_size=self.obs_space_size,
            current_state=np.array([int(env.px[0]), int(env.py[0])]), orientation=0,
            current_target=int(env.target[0]))
        self.done = False
        return self.current_state

    def step(self, action, reward_type=DEFAULT_REWARD_TYPE):                          # SDE:223-267
        if self.done:
            raise RuntimeError("Episode has finished. Call env.reset() to start a new episode.")
        env = self._env
        move = {}
        for key in action.keys():
            move[key] = action_to_move[action[key]]          # KeyError for unknown ids, as SDE:239
        act = np.array([[int(action[i]) for i in range(self.num_agents)]], dtype=np.uint8)
        roll = self._rng.random()
        slab = self._rng.integers(0, 8, size=(1, env.K), dtype=np.uint8)
        self.orientation = action                                                    # SDE:242
        env.step(act, reward_type, retarget=np.array([roll < 0.1], dtype=np.uint8), slab=slab)
        env.check_errors()
        eff = env.eff_action[0].cpu().numpy()
        self.move = {i: action_to_move[int(eff[i])] for i in range(self.num_agents)}  # SDE:252 mutates move
        self.current_state = self._state_dict()
        self.done = bool(env.done[0])
        g = env.reward[0].cpu().numpy().astype(np.float64)
        ar = env.agent_reward[0].cpu().numpy().astype(np.float64)
        eaten = int(env.eaten[0])
        reward = {}
        for i in range(self.num_agents):
            surv = float(self.predator_eat_rew) if (self.done and i == eaten) else 0
            reward[i] = {"survival": surv, "attraction": ar[i, 0], "repulsion": ar[i, 1], "alignment": ar[i, 2],
                         "sum": ar[i, 3]}
        reward["global"] = dict(zip(REWARD_KEYS, g))
        prev = env.pred_prev[0].cpu().numpy().astype(np.int64)
        nxt = env.pred_next[0].cpu().numpy().astype(np.int64)
        self.predator.current_state = nxt.copy()
        self.predator.orientation = int(env.step_orient[0])
        self.predator.current_target = int(env.step_target[0])
        info = {"predator_state": prev, "predator_orientation": self.predator.orientation,
                "predator_next_state": nxt.copy()}
        return self.current_state, reward, self.done, info

    def render(self, mode="rgb_array", close=False):
        raise NotImplementedError("rendering (SDE:408-464) is out of scope of the B200 transition path")

    def close(self):
        if self._env is not None:
            self._env.close()
            self._env = None


DiscreteSwarmEnv = SwarmEnv


class ContinuousSwarmEnv(SwarmEnv):
    """gym_swarm/envs/swarm_continuous_env.py is identical to the discrete file except for the class name."""


def register_gym():
    """Register ``Discrete-Swarm-v0`` / ``Continuous-Swarm-v0`` (gym_swarm/__init__.py:6-14) when a gym
    package is importable; returns the list of ids registered."""
    try:
        from gym.envs.registration import register
    except Exception:
        try:
            from gymnasium.envs.registration import register
        except Exception:
            return []
    register(id="Discrete-Swarm-v0", entry_point="gym_swarm_b200.envs:DiscreteSwarmEnv")
    register(id="Continuous-Swarm-v0", entry_point="gym_swarm_b200.envs:ContinuousSwarmEnv")
    return ["Discrete-Swarm-v0", "Continuous-Swarm-v0"]
