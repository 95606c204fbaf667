"""Multi-GPU host logic on CPU: world_size-2 ``gloo`` job.  Each rank owns a contiguous slice of the
environments and of every tape (gym-swarm_b200/sharding.py), steps it (the CPU tests use the oracle as the
stepper -- the product path has no CPU fallback), and the episode statistics are summed with one all-reduce.
The sharded job must equal the single-process run over all environments."""
import os
import socket

import numpy as np
import pytest

import gym_swarm_b200  # noqa: F401
from gym_swarm_b200 import sharding, tapes

E, N, S, T, R = 10, 8, 10, 60, 8          # E not divisible by 4, small grid: many episodes end


def test_shard_range_partitions():
    for E_ in (1, 7, 16, 65536, 1048576 + 3):
        for W in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(E_, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == E_
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _make_tapes():
    rng = np.random.default_rng(42)
    return tapes.make_reset_tapes(rng, R, E, N, S), tapes.make_step_tapes(rng, T, E, N)


def _run_oracle(rt, st):
    """Steps the oracle over the given (possibly sliced) tapes; returns final state + episode statistics."""
    from oracle import swarm_oracle as so
    e = rt.place.shape[1]
    ora = so.OracleBatch(e, N, S, 5, 2, -10.0, auto_reset=True)
    ora.set_reset_tapes(rt.place, rt.ptape)
    ora.reset()
    ep_len = np.zeros(e)
    ep_ret = np.zeros((e, 4))
    stats = dict.fromkeys(sharding.STAT_KEYS, 0.0)
    o = None
    for t in range(T):
        o = ora.step(st.actions[t], st.slab[t], st.retarget[t], (1.0, 0.5, 0.25), None, False)
        ep_len += 1
        ep_ret += o["glob"][:, :4]
        d = o["done"].astype(bool)
        stats["episodes"] += d.sum()
        stats["sum_length"] += ep_len[d].sum()
        for q, k in enumerate(("sum_survival", "sum_attraction", "sum_repulsion", "sum_alignment")):
            stats[k] += ep_ret[d, q].sum()
        ep_len[d] = 0
        ep_ret[d] = 0
    stats["sum_return"] = sum(stats[k] for k in ("sum_survival", "sum_attraction", "sum_repulsion", "sum_alignment"))
    stats["running_steps"] = ep_len.sum()
    return o, stats


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rt, st = _make_tapes()
        lo, hi = sharding.shard_range(E, world, rank)
        o, stats = _run_oracle(sharding.slice_reset_tapes(rt, lo, hi), sharding.slice_step_tapes(st, lo, hi))
        tot = sharding.allreduce_stats(stats)
        xs = [None] * world
        dist.all_gather_object(xs, (lo, hi, o["x"], o["y"], o["px"]))
        if rank == 0:
            q.put((tot, xs))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gloo_job_equals_single_process(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, xs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rt, st = _make_tapes()
    o, ref = _run_oracle(rt, st)
    assert ref["episodes"] > 0
    for k in sharding.STAT_KEYS:
        assert tot[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-9), k
    for lo, hi, x, y, px in xs:
        np.testing.assert_array_equal(x, o["x"][lo:hi])
        np.testing.assert_array_equal(y, o["y"][lo:hi])
        np.testing.assert_array_equal(px, o["px"][lo:hi])


def test_allreduce_stats_without_process_group_is_identity():
    s = dict(zip(sharding.STAT_KEYS, range(8)))
    assert sharding.allreduce_stats(s) == s


@pytest.mark.gpu
def test_shard_equivalence_on_one_gpu():
    """E envs in one handle == the same envs split over two handles (slices), every output tensor; the summed
    episode statistics of the slices equal the statistics of the whole."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    E_, N_, S_, T_ = 300, 16, 12, 50
    rng = np.random.default_rng(3)
    whole = BatchedSwarmEnv(E_, N_, S_, reset_slots=6, track_stats=True, debug_outputs=True)
    rt = tapes.make_reset_tapes(rng, 6, E_, N_, S_, KR=whole.KR, KP=whole.KP)
    st = tapes.make_step_tapes(rng, T_, E_, N_, K=whole.K)
    whole.set_reset_tapes(rt.place, rt.ptape)
    whole.reset()
    parts = []
    for r in range(2):
        lo, hi = sharding.shard_range(E_, 2, r)
        env = BatchedSwarmEnv(hi - lo, N_, S_, reset_slots=6, track_stats=True, debug_outputs=True)
        s = sharding.slice_reset_tapes(rt, lo, hi)
        env.set_reset_tapes(s.place, s.ptape)
        env.reset()
        parts.append((lo, hi, env))
    for t in range(T_):
        a = whole.step_numpy(st.actions[t], st.slab[t], st.retarget[t], (1.0, 0.5, 0.25), None)
        for lo, hi, env in parts:
            b = env.step_numpy(st.actions[t, lo:hi], st.slab[t, lo:hi], st.retarget[t, lo:hi], (1.0, 0.5, 0.25), None)
            for k in b:
                np.testing.assert_array_equal(a[k][lo:hi], b[k], err_msg=f"step {t} {k}")
    sw = whole.stats()
    assert sw["episodes"] > 0
    sp = [env.stats() for _, _, env in parts]
    for k in ("episodes", "sum_length", "running_steps"):
        assert sw[k] == sp[0][k] + sp[1][k], k
    for k in ("sum_return", "sum_survival", "sum_attraction", "sum_repulsion", "sum_alignment"):
        assert sw[k] == pytest.approx(sp[0][k] + sp[1][k], rel=1e-9, abs=1e-6), k
    whole.close()
    for _, _, env in parts:
        env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("E,N,S,path", [(301, 8, 9, "auto"), (150, 16, 12, "auto"), (40, 100, 40, "auto"), (5, 300, 64, "staged")])
def test_sharded_device_tapes_equal_unsharded(E, N, S, path):
    """In-kernel (Philox) tapes are keyed by the GLOBAL env index (swarm_config.env_offset = the shard's lo): three
    shards of a batch draw exactly what the whole batch draws -- resets, re-draws, retarget flags, auto-resets."""
    import torch
    from gym_swarm_b200.envs import BatchedSwarmEnv
    T_ = 40
    acts = np.random.default_rng(E + N).integers(0, 9, size=(T_, E, N), dtype=np.uint8)
    kw = dict(device_tapes=True, seed=1234, track_stats=True, debug_outputs=True, path=path, place_extra=400,
              pred_attempts=32, slab_len=256)
    whole = BatchedSwarmEnv(E, N, S, **kw)
    whole.reset()
    parts = []
    for r in range(3):
        lo, hi = sharding.shard_range(E, 3, r)
        env = BatchedSwarmEnv(hi - lo, N, S, env_offset=lo, **kw)
        env.reset()
        parts.append((lo, hi, env))
    keys = ("x", "y", "px", "py", "target", "done", "eaten", "reward", "counts", "consumed", "agent_reward", "eff_action",
            "reset_count")
    n_done = 0
    for t in range(T_):
        whole.step(acts[t])
        for lo, hi, env in parts:
            env.step(acts[t, lo:hi])
            for k in keys:
                assert torch.equal(getattr(whole, k)[lo:hi], getattr(env, k)), f"step {t} {k} shard [{lo},{hi})"
        n_done += int(whole.done.sum())
    assert n_done > 0 and int(whole.reset_count.max()) >= 1
    whole.check_errors()
    whole.close()
    for _, _, env in parts:
        env.close()


@pytest.mark.gpu
def test_sharded_env_host_tapes_differ_per_shard():
    """ShardedSwarmEnv passes env_offset = lo, so the shards' internal HOST tape generators are distinct streams (with
    one shared seed every rank used to replay rank 0's placements)."""
    a = sharding.ShardedSwarmEnv(64, 8, 16, rank=0, world_size=2, seed=5)
    b = sharding.ShardedSwarmEnv(64, 8, 16, rank=1, world_size=2, seed=5)
    assert (a.lo, a.hi, b.lo, b.hi) == (0, 32, 32, 64) and b.env.env_offset == 32
    a.reset(), b.reset()
    assert not bool((a.env.x == b.env.x).all())
    a.close(), b.close()
