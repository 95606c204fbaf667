"""Host-side features around the transition: checkpoint / resume, CUDA-graph rollouts, the vector-env adapter."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(E, N, S, **kw):
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    rng = np.random.default_rng(E + N)
    env = BatchedSwarmEnv(E, N, S, reset_slots=4, track_stats=True, debug_outputs=True, **kw)
    rt = tapes.make_reset_tapes(rng, 4, E, N, S, KR=env.KR, KP=env.KP)
    env.set_reset_tapes(rt.place, rt.ptape)
    return env, rng


@pytest.mark.parametrize("E,N,S", [(200, 8, 10), (64, 16, 12), (8, 128, 40), (3, 600, 64)])
def test_state_dict_resume_is_bit_identical(E, N, S):
    from gym_swarm_b200 import tapes
    env, rng = _mk(E, N, S)
    st = tapes.make_step_tapes(rng, 30, E, N, K=env.K)
    env.reset()
    for t in range(10):
        env.step_numpy(st.actions[t], st.slab[t], st.retarget[t])
    ckpt = env.state_dict()
    a = [env.step_numpy(st.actions[t], st.slab[t], st.retarget[t], (1.0, 0.5, 0.25)) for t in range(10, 30)]
    sa = env.stats()
    env2, _ = _mk(E, N, S)
    env2.load_state_dict(ckpt)
    b = [env2.step_numpy(st.actions[t], st.slab[t], st.retarget[t], (1.0, 0.5, 0.25)) for t in range(10, 30)]
    for t, (oa, ob) in enumerate(zip(a, b)):
        dn = oa["done"].astype(bool)
        for k in oa:
            if k in ("term_x", "term_y"):               # only written for the envs that finished in this step
                np.testing.assert_array_equal(oa[k][dn], ob[k][dn], err_msg=f"step {t} {k}")
            else:
                np.testing.assert_array_equal(oa[k], ob[k], err_msg=f"step {t} {k}")
    assert sa == env2.stats() and sa["episodes"] > 0
    env.close()
    env2.close()


def test_state_dict_covers_internal_tape_generator():
    """With the env drawing its own retarget / slab tapes, a resumed env continues the same random streams."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    acts = np.random.default_rng(0).integers(0, 8, size=(40, 50, 8), dtype=np.uint8)
    env = BatchedSwarmEnv(50, 8, 9, seed=11, tape_chunk=16)
    env.reset()
    for t in range(21):
        env.step(acts[t])
    ckpt = env.state_dict()
    ref = []
    for t in range(21, 40):
        env.step(acts[t])
        ref.append((env.x.cpu().numpy().copy(), env.reward.cpu().numpy().copy()))
    env2 = BatchedSwarmEnv(50, 8, 9, seed=999, tape_chunk=16)
    env2.load_state_dict(ckpt)
    for i, t in enumerate(range(21, 40)):
        env2.step(acts[t])
        np.testing.assert_array_equal(env2.x.cpu().numpy(), ref[i][0])
        np.testing.assert_array_equal(env2.reward.cpu().numpy(), ref[i][1])
    env.close()
    env2.close()


@pytest.mark.parametrize("E,N,S", [(4096, 16, 32), (300, 8, 16), (4, 700, 64)])
def test_graphed_rollout_equals_eager_steps(E, N, S):
    import torch
    from gym_swarm_b200 import tapes
    T = 12
    env, rng = _mk(E, N, S)
    st = tapes.make_step_tapes(rng, 2 * T, E, N, K=env.K)
    env.reset()
    snap = env.state_dict()
    eager_r, eager_d = [], []
    for t in range(2 * T):
        env.step(st.actions[t], None, retarget=st.retarget[t], slab=st.slab[t])
        eager_r.append(env.reward.cpu().numpy().copy())
        eager_d.append(env.done.cpu().numpy().copy())
    final = env.state_numpy()
    env.load_state_dict(snap)
    dev = env.device
    a = torch.from_numpy(st.actions[:T].copy()).to(dev)
    r = torch.from_numpy(st.retarget[:T].copy()).to(dev)
    s = torch.from_numpy(st.slab[:T].copy()).to(dev)
    roll = env.capture_rollout(a, r, s)
    assert roll.launches >= T
    for half in range(2):                     # replay twice with refilled tapes: 2T steps in two graph launches
        a.copy_(torch.from_numpy(st.actions[half * T:(half + 1) * T].copy()))
        r.copy_(torch.from_numpy(st.retarget[half * T:(half + 1) * T].copy()))
        s.copy_(torch.from_numpy(st.slab[half * T:(half + 1) * T].copy()))
        rew, done = roll.replay()
        torch.cuda.synchronize()
        for t in range(T):
            np.testing.assert_array_equal(rew[t].cpu().numpy(), eager_r[half * T + t], err_msg=f"reward {half} {t}")
            np.testing.assert_array_equal(done[t].cpu().numpy(), eager_d[half * T + t], err_msg=f"done {half} {t}")
    after = env.state_numpy()
    for k in final:
        np.testing.assert_array_equal(final[k], after[k], err_msg=k)
    env.close()


def test_vector_env_adapter():
    import torch
    from gym_swarm_b200.envs import SwarmVectorEnv
    venv = SwarmVectorEnv(64, num_agents=8, obs_space_size=10, seed=3)
    obs, info = venv.reset()
    assert tuple(obs.shape) == (64, 8, 2) and obs.dtype == torch.int16 and info["predator"].shape == (64, 2)
    g = torch.Generator(device="cuda").manual_seed(0)
    n_term = 0
    for _ in range(60):
        actions = torch.randint(0, 8, (64, 8), dtype=torch.uint8, device="cuda", generator=g)
        obs, reward, terminated, truncated, info = venv.step(actions)
        assert tuple(reward.shape) == (64,) and terminated.dtype == torch.bool and not bool(truncated.any())
        assert int(obs.min()) >= 0 and int(obs.max()) < 10
        # cells stay distinct inside an env (collision re-draw), also right after an auto-reset
        cell = obs[:, :, 0].int() * 10 + obs[:, :, 1].int()
        assert all(len(set(row.tolist())) == 8 for row in cell[:8])
        # a terminated env reports the eat penalty as its reward
        assert torch.all(reward[terminated] == -10)
        n_term += int(terminated.sum())
    assert n_term > 0
    venv.close()


def test_vector_env_reset_seed_restarts_device_tapes():
    """reset(seed) re-keys the in-kernel streams too: the same seed replays the same episode, another seed does not."""
    import torch
    from gym_swarm_b200.envs import SwarmVectorEnv
    venv = SwarmVectorEnv(128, num_agents=8, obs_space_size=10, seed=3, device_tapes=True)
    acts = torch.randint(0, 8, (20, 128, 8), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))

    def rollout(seed):
        obs, _ = venv.reset(seed=seed)
        out = [obs.clone()]
        for t in range(20):
            obs, *_ = venv.step(acts[t])
            out.append(obs.clone())
        return torch.stack(out)
    a, b, c = rollout(7), rollout(8), rollout(7)
    assert torch.equal(a, c) and not torch.equal(a, b)
    venv.close()


def test_two_handles_share_the_shared_memory_opt_in():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per function and device, not per handle: creating a second,
    smaller handle of the same kernel instantiation must not take the first one's shared memory away."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    acts = np.random.default_rng(0).integers(0, 8, size=(64, 16), dtype=np.uint8)
    big = BatchedSwarmEnv(64, 16, 256, path="group")            # fused<8,2> with a 256 x 256 bitmap per env (> 48 KB / CTA)
    big.reset()
    big.step(acts)
    small = BatchedSwarmEnv(64, 16, 32, path="group")
    small.reset()
    small.step(acts)
    big.step(acts)                                              # used to fail with cudaErrorInvalidValue
    big.check_errors()
    bigs = BatchedSwarmEnv(2, 512, 1024, path="staged")         # staged: 128 KB bitmap
    bigs.reset()
    smalls = BatchedSwarmEnv(2, 512, 64, path="staged")
    smalls.reset()
    a2 = np.random.default_rng(1).integers(0, 8, size=(2, 512), dtype=np.uint8)
    smalls.step(a2)
    bigs.step(a2)
    bigs.check_errors()
    import torch
    torch.cuda.synchronize()
    for e in (big, small, bigs, smalls):
        e.close()


@pytest.mark.parametrize("E,N,S,path", [(64, 8, 6, "auto"), (40, 16, 8, "auto"), (8, 100, 20, "auto"), (3, 300, 24, "staged")])
def test_frozen_envs_have_defined_outputs(E, N, S, path):
    """auto_reset off: a finished env freezes; a further step on it flags STEP_WHEN_DONE and writes DEFINED outputs:
    zero rewards / counts, done = 1, eaten = -1, the predator where it stands, action 8 (no move)."""
    import torch
    from gym_swarm_b200 import _lib
    from gym_swarm_b200.envs import BatchedSwarmEnv
    env = BatchedSwarmEnv(E, N, S, auto_reset=False, debug_outputs=True, path=path, seed=2, place_extra=600, slab_len=512,
                          pred_attempts=64)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(300):
        env.step(torch.randint(0, 8, (E, N), dtype=torch.uint8, device="cuda", generator=g))
        if bool(env.frozen.any()):
            break
    fz = env.frozen.bool().clone()
    assert bool(fz.any()), "no episode finished"
    x0, y0, px0, py0, tg0 = env.x.clone(), env.y.clone(), env.px.clone(), env.py.clone(), env.target.clone()
    for t in (env.reward, env.agent_reward, env.pred_prev, env.pred_next):
        t.fill_(7)
    for t in (env.counts, env.agent_counts, env.step_target, env.eff_action, env.step_orient, env.consumed):
        t.fill_(5)
    env.step(torch.randint(0, 8, (E, N), dtype=torch.uint8, device="cuda", generator=g))
    assert torch.equal(env.x[fz], x0[fz]) and torch.equal(env.y[fz], y0[fz])
    assert bool((env.done[fz] == 1).all()) and bool((env.eaten[fz] == -1).all())
    assert bool((env.err[fz] & _lib.ERR_STEP_WHEN_DONE).bool().all())
    for t in (env.reward, env.agent_reward, env.counts, env.agent_counts, env.consumed):
        assert not bool(t[fz].any()), "frozen env: reward / count outputs must be zero"
    pred = torch.stack([px0, py0], 1)
    assert torch.equal(env.pred_prev[fz], pred[fz]) and torch.equal(env.pred_next[fz], pred[fz])
    assert torch.equal(env.step_target[fz], tg0[fz]) and bool((env.eff_action[fz] == 8).all())
    env.close()


def test_async_statistics_match_the_synchronous_read():
    """swarm_stats_begin / swarm_stats_end: results come back in order, each equal to what a synchronous read at the
    same point of the stream returns, and collecting them never needs a host-side stream synchronisation."""
    import torch
    from gym_swarm_b200.envs import BatchedSwarmEnv
    E, N, S = 5000, 8, 8
    env = BatchedSwarmEnv(E, N, S, track_stats=True, seed=4)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.randint(0, 8, (48, E, N), dtype=torch.uint8, device="cuda", generator=g)
    sync_reads = []
    for t in range(48):
        env.step(acts[t])
        if t % 16 == 15:
            env.stats_begin()
            sync_reads.append(env.stats())       # (drains the asynchronous one first: keep a copy of it)
    # the synchronous read consumed the begun results in order; do it again purely asynchronously
    env2 = BatchedSwarmEnv(E, N, S, track_stats=True, seed=4)
    env2.reset()
    for t in range(48):
        env2.step(acts[t])
        if t % 16 == 15:
            env2.stats_begin()
    first = env2.stats_end(wait=False)                            # None while in flight; never blocks, never raises
    async_reads = [] if first is None else [first]
    while len(async_reads) < 3:
        async_reads.append(env2.stats_end(wait=True))
    assert async_reads == sync_reads
    assert sync_reads[-1]["episodes"] > 0 and sync_reads[0]["episodes"] <= sync_reads[-1]["episodes"]
    assert sync_reads[-1]["running_steps"] + sync_reads[-1]["sum_length"] == 48 * E
    with pytest.raises(Exception):
        env2.stats_end(wait=True)                                 # nothing outstanding
    env.close()
    env2.close()


@pytest.mark.parametrize("E,N,S,path,vf", [(300, 8, 8, "auto", None), (64, 16, 16, "auto", None), (16, 128, 64, "auto", None),
                                           (16, 128, 64, "auto", 3), (16, 100, 256, "auto", 9), (3, 300, 48, "staged", None),
                                           (3, 300, 48, "staged", 4)])
def test_compact_agent_terms_decode_to_agent_rewards(E, N, S, path, vf):
    """swarm_step_out.agent_terms (int16 x 4 per agent) carries exactly the integers the per-agent float rewards are
    functions of: decoded with the step's weights it equals agent_reward (f32) to f32 rounding, on every kernel path,
    with and without vf_size, terminal steps included."""
    import torch
    from gym_swarm_b200.envs import BatchedSwarmEnv
    env = BatchedSwarmEnv(E, N, S, per_agent=True, agent_terms=True, debug_outputs=True, path=path, seed=3,
                          place_extra=600, slab_len=512, pred_attempts=64)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    n_done = 0
    for t in range(40):
        rt = {"attraction": 1.0, "repulsion": 0.25 + 0.01 * t, "alignment": 0.5, "vf_size": vf}
        _, reward, done, _ = env.step(torch.randint(0, 9, (E, N), dtype=torch.uint8, device="cuda", generator=g), rt)
        live = ~done.bool()
        want = reward["agents"].double()[live]
        got = env.decode_agent_terms(reward["terms"], rt)[live]
        assert torch.allclose(got, want, rtol=2e-7, atol=0.0), f"step {t}: {float((got - want).abs().max())}"
        assert not bool(reward["terms"][done.bool()].any()), "terminal step: all four terms are 0 (SDE:321-327)"
        assert torch.equal(reward["terms"][..., 0].int(), env.agent_counts[..., 0])
        assert torch.equal(reward["terms"][..., 1].int(), env.agent_counts[..., 1])
        n_done += int(done.sum())
    assert n_done > 0 or S > 100
    env.close()


def test_fetch_host_is_one_packed_copy_of_the_step_results():
    import torch
    from gym_swarm_b200.envs import BatchedSwarmEnv
    E, N, S = 500, 16, 16
    env = BatchedSwarmEnv(E, N, S, per_agent=True, agent_terms=True, seed=9)
    env.reset()
    hin = env.host_inputs()
    rng = np.random.default_rng(0)
    for t in range(5):
        hin["actions"].copy_(torch.from_numpy(rng.integers(0, 8, size=(E, N), dtype=np.uint8)))
        hin["retarget"].copy_(torch.from_numpy((rng.random(E) < 0.1).astype(np.uint8)))
        hin["slab"].copy_(torch.from_numpy(rng.integers(0, 8, size=(E, env.K), dtype=np.uint8)))
        env.step_host()
        h = env.fetch_host(sync=True)
        for k in ("x", "y", "done", "reward", "agent_reward", "agent_terms"):
            assert torch.equal(h[k], getattr(env, k).cpu()), k
    assert env.fetch_bytes >= E * N * (4 + 16 + 8) + E * 21 and env.input_bytes >= E * (N + 1 + env.K)
    env.close()
