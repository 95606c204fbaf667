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
