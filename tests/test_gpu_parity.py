"""GPU parity tests proper: the CUDA path (through the C ABI / ctypes) against
  (1) the golden trajectories recorded from the unmodified reference (tests/golden/*.npz),
  (2) the NumPy oracle on fresh seeded inputs,
  (3) size-independent properties at BASELINE.json's full shapes (tests/test_gpu_properties.py).
Bit-exact for positions / predator / target / done / integer counts; float rewards within the
north-star tolerance 1e-6 relative (golden_utils.RTOL)."""
import numpy as np
import pytest

from golden_utils import assert_float_parity, case_names, load_case, replay

pytestmark = pytest.mark.gpu


def make_env(cfg, path="auto", **kw):
    from gym_swarm_b200.envs import BatchedSwarmEnv
    return BatchedSwarmEnv(cfg["E"], cfg["N"], cfg["S"], cfg["att"], cfg["rep"], cfg["eat"],
                           auto_reset=cfg["auto_reset"], slab_len=cfg["K"], place_extra=cfg["KR"],
                           pred_attempts=cfg["KP"], reset_slots=cfg["R"], per_agent=True, debug_outputs=True,
                           path=path, **kw)


class CudaImpl:
    def __init__(self, cfg, z, path, **kw):
        self.env = make_env(cfg, path, **kw)
        self.env.set_reset_tapes(z["place"], z["ptape"])

    def reset(self):
        self.env.reset()
        return self.env.state_numpy()

    def step(self, actions, slab, retarget, weights, vf, want_agents):
        return self.env.step_numpy(actions, slab, retarget, weights, vf)


@pytest.mark.parametrize("name", case_names())
def test_golden_default_path(name):
    cfg, z = load_case(name)
    impl = CudaImpl(cfg, z, "auto")
    replay(impl, cfg, z)
    impl.env.close()


@pytest.mark.parametrize("name", [n for n in case_names() if not n.startswith("large")])
def test_golden_staged_path(name):
    """The staged kernels (the large-N path) must reproduce the same fixtures at small N too."""
    cfg, z = load_case(name)
    impl = CudaImpl(cfg, z, "staged")
    replay(impl, cfg, z, steps=60)
    impl.env.close()


def _small_n_cases(lo, hi):
    out = []
    for n in case_names():
        cfg, _ = load_case(n)
        if lo <= cfg["N"] <= hi:
            out.append(n)
    return out


@pytest.mark.parametrize("name", _small_n_cases(2, 8))
def test_golden_lane_group_kernel_small_n(name):
    """Fixtures with N <= 8 normally run on the thread-per-env kernel; the lane-group kernel must reproduce them too."""
    cfg, z = load_case(name)
    impl = CudaImpl(cfg, z, "group")
    replay(impl, cfg, z, steps=120)
    impl.env.close()


@pytest.mark.parametrize("name", _small_n_cases(9, 16))
def test_golden_thread_per_env_kernel_16(name):
    """Fixtures with 9..16 agents through the 16-agent thread-per-env kernel (normally reserved for huge batches)."""
    cfg, z = load_case(name)
    impl = CudaImpl(cfg, z, "auto", tiny16_min_envs=1)
    assert impl.env.path_name.startswith("tiny<16>") or cfg["vf"] is not None, impl.env.path_name
    replay(impl, cfg, z, steps=150)
    impl.env.close()


def _oracle_lockstep(E, N, S, T, seed, path="auto", att=5, rep=2, eat=-10.0, vf=None, n_actions=8, K=None, KR=None,
                     weights=(1.0, 1.0, 1.0), auto_reset=True, check_agents=True, agent_out=True, flags=0,
                     tiny16_min_envs=0, KP=8):
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    from oracle import swarm_oracle as so
    rng = np.random.default_rng(seed)
    env = BatchedSwarmEnv(E, N, S, att, rep, eat, auto_reset=auto_reset, slab_len=K, place_extra=KR, reset_slots=3,
                          per_agent=agent_out, debug_outputs=agent_out, path=path, flags=flags,
                          tiny16_min_envs=tiny16_min_envs, pred_attempts=KP)
    check_agents = check_agents and agent_out
    rt = tapes.make_reset_tapes(rng, env.R, E, N, S, KR=env.KR, KP=env.KP)
    st = tapes.make_step_tapes(rng, T, E, N, K=env.K, n_actions=n_actions)
    env.set_reset_tapes(rt.place, rt.ptape)
    ora = so.OracleBatch(E, N, S, att, rep, eat, auto_reset)
    ora.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    a, b = env.state_numpy(), ora.reset()
    for k in ("x", "y", "px", "py", "target"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"reset {k}")
    for t in range(T):
        a = env.step_numpy(st.actions[t], st.slab[t], st.retarget[t], weights, vf)
        b = ora.step(st.actions[t], st.slab[t], st.retarget[t], weights, vf, check_agents)
        live = b["frozen"] == 0
        for k in ("x", "y", "px", "py", "target", "done"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"step {t} {k}")
        for k in ("eaten", "cnt", "pred_prev", "pred_next", "pred_orient", "step_target", "consumed"):
            np.testing.assert_array_equal(a[k][live], b[k][live], err_msg=f"step {t} {k}")
        assert_float_parity(a["glob"][live], b["glob"][live], err_msg=f"step {t} glob")
        if check_agents:
            np.testing.assert_array_equal(a["agent_cnt"][live], b["agent_cnt"][live], err_msg=f"step {t} agent_cnt")
            assert_float_parity(a["agent"][live], b["agent"][live], err_msg=f"step {t} agent")
            np.testing.assert_array_equal(a["eff"][live], b["eff"][live], err_msg=f"step {t} eff action")
    env.close()


@pytest.mark.parametrize("N,S", [(2, 4), (3, 5), (4, 20), (5, 9), (8, 16), (9, 16), (16, 32), (17, 12), (31, 40),
                                 (32, 32), (33, 64), (64, 64), (100, 48), (128, 128), (129, 50), (255, 64),
                                 (256, 300), (256, 128), (200, 96), (136, 32)])
def test_oracle_lockstep_fused_shapes(N, S):
    dense = S * S < 4 * N
    _oracle_lockstep(E=9, N=N, S=S, T=25, seed=N * 1000 + S, n_actions=9, K=512 if dense else None,
                     KR=2000 if dense else None)


@pytest.mark.parametrize("N,S", [(2, 4), (3, 5), (4, 20), (5, 9), (7, 6), (8, 16), (8, 5)])
def test_oracle_lockstep_tiny_and_group_kernels(N, S):
    """N <= 8: the thread-per-env kernel (default) and the lane-group kernel it replaces (path="group", still the
    kernel behind vf_size steps) both reproduce the oracle; many envs so that every lane of a warp is used."""
    dense = S * S < 4 * N
    for path in ("auto", "group"):
        _oracle_lockstep(E=70, N=N, S=S, T=30, seed=N * 77 + S, n_actions=9, path=path, K=512 if dense else None,
                         KR=2000 if dense else None, weights=(0.5, 1.0, 0.25))
    _oracle_lockstep(E=40, N=N, S=S, T=10, seed=N + S, n_actions=9, vf=2, K=512 if dense else None,
                     KR=2000 if dense else None)
    # env totals only (no per-agent outputs): the packed-counter epilogue of tiny_step_kernel<false>
    for (att, rep) in ((5, 2), (2, 5), (0, 0), (30, 3), (10, 200)):
        _oracle_lockstep(E=70, N=N, S=S, T=12, seed=N * 5 + S + att, n_actions=9, att=att, rep=rep, agent_out=False,
                         K=512 if dense else None, KR=2000 if dense else None, weights=(0.5, 1.0, 0.25))


@pytest.mark.parametrize("S", [16, 5, 4, 40, 300])
def test_oracle_lockstep_packed_tiny_kernel(S):
    """Exactly 8 agents without per-agent outputs: tiny8_packed_kernel (positions never unpacked; PRMT action table,
    packed wrap, nibble collision counters) against the oracle -- action 8, dense grids with several re-draw rounds,
    every threshold plan, frozen envs without auto-reset -- and the scalar kernel it replaces (FLAG_TINY_SCALAR)."""
    from gym_swarm_b200 import _lib
    from gym_swarm_b200.envs import BatchedSwarmEnv
    dense = S * S < 32
    kw = dict(K=512 if dense else None, KR=4000 if dense else None, KP=64 if dense else 8)
    for flags in (0, _lib.FLAG_TINY_SCALAR):
        _oracle_lockstep(E=300, N=8, S=S, T=40, seed=S * 13 + flags, n_actions=9, agent_out=False, flags=flags,
                         weights=(0.5, 1.0, 0.25), **kw)
    for (att, rep) in ((5, 2), (2, 5), (0, 0), (30, 3), (10, 200), (1, 1)):
        _oracle_lockstep(E=100, N=8, S=S, T=12, seed=S + att, n_actions=9, att=att, rep=rep, agent_out=False, **kw)
    _oracle_lockstep(E=200, N=8, S=S, T=60, seed=S + 5, n_actions=8, agent_out=False, auto_reset=False, **kw)
    env = BatchedSwarmEnv(64, 8, 16)
    assert env.path_name.startswith("tiny<8>"), env.path_name
    env.close()


@pytest.mark.parametrize("N,S", [(9, 16), (12, 7), (13, 40), (16, 32), (16, 8)])
def test_oracle_lockstep_tiny16_kernel(N, S):
    """9..16 agents: the 16-agent instantiation of the thread-per-env kernel (normally chosen only for batches of
    >= 65,536 envs; swarm_config.tiny16_min_envs lowers the bar so that the oracle can follow)."""
    from gym_swarm_b200.envs import BatchedSwarmEnv
    env = BatchedSwarmEnv(4, N, S, tiny16_min_envs=1)
    assert env.path_name.startswith("tiny<16>"), env.path_name
    env.close()
    dense = S * S < 4 * N
    _oracle_lockstep(E=70, N=N, S=S, T=25, seed=N * 7 + S, n_actions=9, K=512 if dense else None,
                     KR=2000 if dense else None, weights=(0.5, 1.0, 0.25))
    for (att, rep) in ((5, 2), (2, 5), (0, 0), (30, 3), (10, 200)):
        _oracle_lockstep(E=70, N=N, S=S, T=8, seed=N * 3 + S + att, n_actions=9, att=att, rep=rep, agent_out=False,
                         K=512 if dense else None, KR=2000 if dense else None)


def test_tiny_kernel_is_the_default_for_small_n():
    from gym_swarm_b200.envs import BatchedSwarmEnv
    for N, name in ((8, "tiny<8>"), (4, "tiny<8>"), (9, "fused<8,2>"), (16, "fused<8,2>")):
        env = BatchedSwarmEnv(4, N, 20)
        assert env.path_name.startswith(name), env.path_name
        env.close()
    env = BatchedSwarmEnv(4, 8, 20, path="group")
    assert env.path_name.startswith("fused<8,1>")
    env.close()


@pytest.mark.parametrize("N,S", [(257, 64), (300, 96), (512, 512), (1000, 100), (2048, 1024)])
def test_oracle_lockstep_staged_shapes(N, S):
    """Large-N path, reward counts by the box-count kernel (default) and by the tiled pairwise kernels."""
    _oracle_lockstep(E=3, N=N, S=S, T=4, seed=N + S, n_actions=9)
    _oracle_lockstep(E=3, N=N, S=S, T=3, seed=N + S + 1, n_actions=9, path="staged_pairwise")


@pytest.mark.parametrize("att,rep", [(5, 2), (2, 5), (0, 0), (1, 1), (9, 3), (3, 12), (10, 200), (70, 3), (200, 200),
                                     (64, 65), (5, 60)])
def test_box_count_path_thresholds(att, rep):
    """Every method of the box-count plan (zero / all / direct box / complement with histograms and corner boxes)
    against the oracle, on a grid small enough that border bands and corners are populated; thresholds the plan
    rejects fall back to the pairwise kernels inside the same handle."""
    _oracle_lockstep(E=3, N=700, S=64, T=3, seed=att * 31 + rep, n_actions=9, att=att, rep=rep, weights=(0.5, 1.0, 0.25))
    _oracle_lockstep(E=2, N=1500, S=50, T=2, seed=att + rep, n_actions=9, att=att, rep=rep, K=4096, KR=6000)


@pytest.mark.parametrize("N,S", [(16, 32), (24, 12), (33, 64), (64, 64), (100, 48), (128, 128), (255, 64)])
def test_fused_reward_by_ring_sweep(N, S):
    """The fused lane-group kernel's reward counts (symmetric ring sweep + border pass) on every ring shape, dense grids
    included."""
    dense = S * S < 4 * N
    _oracle_lockstep(E=9, N=N, S=S, T=12, seed=N * 3 + S, n_actions=9, K=512 if dense else None,
                     KR=2000 if dense else None, weights=(0.5, 1.0, 0.25))


@pytest.mark.parametrize("N,S,att,rep", [(16, 32, 5, 2), (24, 12, 5, 2), (33, 64, 5, 2), (64, 64, 9, 3), (100, 96, 16, 1),
                                         (128, 128, 5, 2), (128, 128, 2, 5), (128, 128, 0, 0), (128, 128, 17, 2),
                                         (255, 64, 5, 2), (200, 300, 5, 2), (12, 5, 3, 1), (128, 256, 16, 16),
                                         # 8 agents per lane (N = 129 .. 256): box popcounts + border pass without a pair sweep
                                         (256, 128, 5, 2), (256, 128, 16, 16), (136, 96, 9, 3), (200, 64, 2, 5), (129, 32, 5, 2),
                                         (256, 128, 0, 0), (248, 160, 17, 2)])
@pytest.mark.parametrize("ring", [False, True])
def test_two_launch_step(N, S, att, rep, ring):
    """The lane-group step as two launches (transition, then reward; the default from 2^22 agents per handle up, forced
    here by SWARM_FLAG_FUSED_SPLIT): reward counts by box popcounts on the rebuilt occupancy bitmap where they apply
    (word-aligned rows, radii <= 15, far thresholds from the border pass) and by the ring sweep otherwise / when
    SWARM_FLAG_NO_BOX_NEAR asks for it; terminal steps and auto-resets come from the transition launch."""
    from gym_swarm_b200 import _lib
    dense = S * S < 4 * N
    _oracle_lockstep(E=11, N=N, S=S, T=12, seed=N * 5 + S + att, n_actions=9, att=att, rep=rep, K=512 if dense else None,
                     KR=2000 if dense else None, KP=64 if dense else 8, weights=(0.5, 1.0, 0.25),
                     flags=_lib.FLAG_FUSED_SPLIT | (_lib.FLAG_NO_BOX_NEAR if ring else 0))


@pytest.mark.parametrize("N,S,vf", [(16, 32, 2), (60, 64, 5), (128, 128, 5), (37, 50, 4)])
@pytest.mark.parametrize("dense", [False, True])
def test_two_launch_step_with_vf(N, S, vf, dense):
    from gym_swarm_b200 import _lib
    _oracle_lockstep(E=13, N=N, S=S, T=6, seed=N + S + vf, vf=vf, weights=(0.5, 1.0, 2.0), n_actions=9,
                     flags=_lib.FLAG_FUSED_SPLIT | (_lib.FLAG_VF_DENSE if dense else 0))


def test_two_launch_step_without_debug_outputs():
    """No eff_action bound: the effective actions travel between the two launches in the library's own scratch."""
    from gym_swarm_b200 import _lib
    _oracle_lockstep(E=40, N=128, S=128, T=8, seed=3, n_actions=9, flags=_lib.FLAG_FUSED_SPLIT, agent_out=False)


@pytest.mark.parametrize("N,S", [(16, 32), (24, 12), (64, 64), (60, 11), (128, 128)])
def test_fused_four_threshold_sweep(N, S):
    """Default thresholds normally take the border pass for the two far thresholds + a 2-threshold ring sweep;
    SWARM_FLAG_NO_FAR_BORDER keeps all four thresholds in the sweep (the path other threshold combinations use)."""
    from gym_swarm_b200 import _lib
    dense = S * S < 4 * N
    _oracle_lockstep(E=9, N=N, S=S, T=10, seed=N + S, n_actions=9, K=512 if dense else None, KR=2000 if dense else None,
                     flags=_lib.FLAG_NO_FAR_BORDER)


@pytest.mark.parametrize("att,rep", [(5, 2), (2, 5), (1, 1), (9, 3), (3, 12), (12, 12), (20, 3), (31, 31), (32, 5)])
def test_fused_far_border_pass_thresholds(att, rep):
    """Far thresholds S-att+1, S-rep+1 on both sides of S/2 (S = 64, 21): the border pass is taken iff both exceed S/2;
    small grids put most agents into the border band."""
    _oracle_lockstep(E=6, N=120, S=64, T=4, seed=att * 17 + rep, n_actions=9, att=att, rep=rep, weights=(0.5, 1.0, 0.25))
    _oracle_lockstep(E=6, N=40, S=21, T=4, seed=att + rep + 1, n_actions=9, att=min(att, 21), rep=min(rep, 21))
    _oracle_lockstep(E=6, N=16, S=7, T=6, seed=att + rep + 2, n_actions=9, att=min(att, 7), rep=min(rep, 7), K=256, KR=500)


@pytest.mark.parametrize("att,rep", [(5, 2), (2, 5), (0, 0), (1, 1), (9, 3), (3, 12), (10, 200), (40, 3), (70, 70), (5, 60)])
def test_staged_box_count_thresholds(att, rep):
    """Box counts on the occupancy bitmap (the staged path's default reward counts): zero / all / direct / complement plans."""
    _oracle_lockstep(E=3, N=600, S=64, T=3, seed=att * 13 + rep, n_actions=9, att=att, rep=rep, weights=(0.5, 1.0, 0.25),
                     path="staged")
    _oracle_lockstep(E=3, N=300, S=45, T=3, seed=att + rep, n_actions=9, att=min(att, 45), rep=min(rep, 45), path="staged")


@pytest.mark.parametrize("E,N,S,path", [(300, 8, 6, "auto"), (64, 5, 9, "group"), (40, 16, 8, "auto"), (9, 128, 40, "auto"),
                                        (6, 200, 300, "auto"), (3, 300, 40, "auto"), (3, 300, 40, "staged_pairwise"),
                                        (20, 12, 5, "staged")])
def test_device_tapes_equal_host_materialised_tapes(E, N, S, path):
    """Counter-based tapes: with no tape bound the kernels draw retarget flags, re-draw actions, placements and predator
    attempts from Philox4x32-10; the NumPy twin materialises the same values, and the oracle fed with them reproduces
    the trajectory (resets included: dense grids finish episodes quickly)."""
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    from oracle import swarm_oracle as so
    T, R, seed = 25, 40, 1234 + N
    dense = S * S < 6 * N
    env = BatchedSwarmEnv(E, N, S, auto_reset=True, per_agent=True, debug_outputs=True, path=path, seed=seed,
                          device_tapes=True, reset_slots=R, slab_len=512 if dense else None,
                          place_extra=600 if dense else None, pred_attempts=32)
    rt = tapes.philox_reset_tapes(env.reset_seed, R, E, N, S, KR=env.KR, KP=env.KP)
    ora = so.OracleBatch(E, N, S, 5, 2, -10.0, True)
    ora.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    a, b = env.state_numpy(), ora.reset()
    for k in ("x", "y", "px", "py", "target"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"reset {k}")
    act_rng = np.random.default_rng(seed)
    for t in range(T):
        actions = act_rng.integers(0, 9, size=(E, N), dtype=np.uint8)
        retarget, slab = tapes.philox_step_tapes(seed, t, E, env.K)
        env.step(actions, {"attraction": 1.0, "repulsion": 0.5, "alignment": 0.25, "vf_size": None})
        a = env.state_numpy()
        b = ora.step(actions, slab, retarget, (1.0, 0.5, 0.25), None, True)
        for k in ("x", "y", "px", "py", "target"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"step {t} {k}")
        np.testing.assert_array_equal(env.done.cpu().numpy(), b["done"], err_msg=f"step {t} done")
        np.testing.assert_array_equal(env.counts.cpu().numpy(), b["cnt"], err_msg=f"step {t} counts")
        np.testing.assert_array_equal(env.consumed.cpu().numpy(), b["consumed"], err_msg=f"step {t} consumed")
    assert int(ora.reset_count.max()) <= R, "raise R: the host twin only holds R reset slots"
    if S <= 10:
        assert int(ora.reset_count.sum()) > E, "no episode ended: the reset streams were not exercised"
    assert not np.any(a["err"])
    env.close()


def test_large_n_predator_stream_philox_and_frozen():
    """predator_block_async_kernel inside swarm_step (N >= 4096, E % 4 == 0): the retarget flags drawn in the kernel
    (device tapes) and frozen envs (auto_reset off; at 25 % density episodes end within a few steps)."""
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    from oracle import swarm_oracle as so
    E, N, S, T, R, seed = 4, 4096, 128, 4, 8, 77
    env = BatchedSwarmEnv(E, N, S, auto_reset=True, per_agent=False, debug_outputs=False, seed=seed, device_tapes=True,
                          reset_slots=R, pred_attempts=32)
    assert env.path_name.startswith("staged"), env.path_name
    rt = tapes.philox_reset_tapes(env.reset_seed, R, E, N, S, KR=env.KR, KP=env.KP)
    ora = so.OracleBatch(E, N, S, 5, 2, -10.0, True)
    ora.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    ora.reset()
    act_rng = np.random.default_rng(seed)
    for t in range(T):
        actions = act_rng.integers(0, 9, size=(E, N), dtype=np.uint8)
        retarget, slab = tapes.philox_step_tapes(seed, t, E, env.K)
        env.step(actions, {"attraction": 1.0, "repulsion": 0.5, "alignment": 0.25, "vf_size": None})
        a = env.state_numpy()
        b = ora.step(actions, slab, retarget, (1.0, 0.5, 0.25), None, False)
        for k in ("x", "y", "px", "py", "target"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"step {t} {k}")
        np.testing.assert_array_equal(env.done.cpu().numpy(), b["done"], err_msg=f"step {t} done")
    env.close()
    _oracle_lockstep(E=4, N=4096, S=128, T=8, seed=9, n_actions=8, auto_reset=False, agent_out=False)


def test_staged_path_names():
    from gym_swarm_b200.envs import BatchedSwarmEnv
    for kw, name in ((dict(path="auto"), "staged+box"), (dict(path="staged_pairwise"), "staged+bitmap")):
        env = BatchedSwarmEnv(2, 1024, 256, **kw)
        assert env.path_name.startswith(name), env.path_name
        env.close()


def test_oracle_lockstep_large_n16384():
    """BASELINE configs[3] shape (16384 agents on 1024^2), 2 envs x 2 steps against the oracle."""
    _oracle_lockstep(E=2, N=16384, S=1024, T=2, seed=4, check_agents=True)


@pytest.mark.parametrize("N,S,vf", [(12, 24, 1), (16, 32, 2), (30, 40, 2), (60, 64, 5), (100, 128, 0), (128, 128, 5),
                                    (128, 200, 12), (37, 50, 4), (128, 16384, 9)])
@pytest.mark.parametrize("dense", [False, True])
def test_vf_sparse_and_dense_modes(N, S, vf, dense):
    """vf_size steps on the ring-sweep shapes: the sparse mode (visible pairs recorded by the symmetric sweep, picked when
    (2 vf + 1)^2 <= S^2 / 16) and the dense mode (every ordered pair gated; SWARM_FLAG_VF_DENSE forces it) both reproduce the
    oracle; att / rep far thresholds, action 8 and auto-reset included."""
    from gym_swarm_b200 import _lib
    for (att, rep) in ((5, 2), (2, 5)):
        _oracle_lockstep(E=37, N=N, S=S, T=6, seed=N + S + vf, vf=vf, att=att, rep=rep, weights=(0.5, 1.0, 2.0),
                         n_actions=9, flags=_lib.FLAG_VF_DENSE if dense else 0)


@pytest.mark.parametrize("N,S", [(16, 32), (64, 64), (128, 128)])
def test_vf_sparse_with_four_threshold_sweep(N, S):
    """Sparse visual-field mode on the ring sweep that carries all four thresholds (no border pass), the variant taken
    when a far threshold is at or below S/2."""
    from gym_swarm_b200 import _lib
    _oracle_lockstep(E=9, N=N, S=S, T=6, seed=N + S, vf=2, n_actions=9, weights=(0.5, 1.0, 2.0),
                     flags=_lib.FLAG_NO_FAR_BORDER)


@pytest.mark.parametrize("vf", [-1, 2.5])
def test_vf_negative_and_fractional(vf):
    """SDE:355 `unalign[dist_1 > vf_size] = 0` literally: a negative vf_size clears every pair, a fractional one acts
    like its floor (distances are integers)."""
    _oracle_lockstep(E=5, N=24, S=20, T=8, seed=3, vf=vf, weights=(0.5, 1.0, 2.0), n_actions=9)
    _oracle_lockstep(E=5, N=8, S=12, T=8, seed=4, vf=vf, weights=(0.5, 1.0, 2.0), n_actions=9)


@pytest.mark.parametrize("vf", [0, 3, 40])
def test_vf_gate(vf):
    _oracle_lockstep(E=5, N=24, S=20, T=20, seed=vf, vf=vf, weights=(0.5, 1.0, 2.0), n_actions=9)
    _oracle_lockstep(E=2, N=300, S=64, T=3, seed=vf, vf=vf, weights=(0.5, 1.0, 2.0), n_actions=9)


def test_fused_equals_staged_at_cfg2():
    """Same tapes through both kernel paths: every output tensor identical."""
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    E, N, S, T = 512, 16, 32, 30
    rng = np.random.default_rng(0)
    envs = [BatchedSwarmEnv(E, N, S, path=p, reset_slots=4, debug_outputs=True) for p in ("fused", "staged")]
    rt = tapes.make_reset_tapes(rng, 4, E, N, S, KR=envs[0].KR, KP=envs[0].KP)
    st = tapes.make_step_tapes(rng, T, E, N, K=envs[0].K)
    for e in envs:
        e.set_reset_tapes(rt.place, rt.ptape)
        e.reset()
    for t in range(T):
        outs = [e.step_numpy(st.actions[t], st.slab[t], st.retarget[t], (1.0, 0.7, 0.3), None) for e in envs]
        for k in outs[0]:
            np.testing.assert_array_equal(outs[0][k], outs[1][k], err_msg=f"step {t} {k}")
    for e in envs:
        e.close()


def test_errors_and_freeze():
    from gym_swarm_b200 import _lib, tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    E, N, S = 4, 12, 4            # 12 agents on 16 cells: collisions every step
    rng = np.random.default_rng(1)
    env = BatchedSwarmEnv(E, N, S, slab_len=2, place_extra=4000, pred_attempts=64, reset_slots=1, auto_reset=False)
    rt = tapes.make_reset_tapes(rng, 1, E, N, S, KR=4000, KP=64)
    env.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    st = tapes.make_step_tapes(rng, 30, E, N, K=2)
    for t in range(30):
        env.step(st.actions[t], retarget=st.retarget[t], slab=st.slab[t])
    err = env.state_numpy()["err"]
    assert np.all(err & (_lib.ERR_SLAB_OVERFLOW | _lib.ERR_STEP_WHEN_DONE)), err
    with pytest.raises(_lib.SwarmLibraryError):
        env.check_errors()
    env2 = BatchedSwarmEnv(2, 4, 20, auto_reset=True)
    env2.reset()
    a = np.full((2, 4), 3, np.uint8)
    a[1, 2] = 9
    env2.step(a)
    err2 = env2.state_numpy()["err"]
    assert err2[0] == 0 and err2[1] & _lib.ERR_BAD_ACTION
    env.close()
    env2.close()


def test_bad_action_ids_in_the_packed_kernel():
    """Action ids > 8 in the packed 8-agent kernel: error bit, and the step is the one with id 8 (no move) in their place
    (the scalar kernel's rule, tiny_step.cuh)."""
    from gym_swarm_b200 import _lib
    from gym_swarm_b200.envs import BatchedSwarmEnv
    rng = np.random.default_rng(5)
    E = 96
    acts = rng.integers(0, 9, size=(6, E, 8)).astype(np.uint8)
    bad = acts.copy()
    mask = rng.random(size=bad.shape) < 0.05
    mask[:, 0, :] = False                                   # env 0 stays clean
    bad[mask] = rng.integers(9, 256, size=int(mask.sum())).astype(np.uint8)
    clean = acts.copy()
    clean[mask] = 8
    envs = [BatchedSwarmEnv(E, 8, 16, auto_reset=True, per_agent=False, seed=3) for _ in range(2)]
    assert envs[0].path_name.startswith("tiny<8>")
    for e in envs:
        e.reset()
    for t in range(6):
        envs[0].step(bad[t])
        envs[1].step(clean[t])
    a, b = envs[0].state_numpy(), envs[1].state_numpy()
    for k in ("x", "y", "px", "py", "target"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert a["err"][0] == 0 and b["err"].max() == 0
    hit = mask.any(axis=(0, 2))
    assert np.all((a["err"][hit] & _lib.ERR_BAD_ACTION) != 0) and np.all(a["err"][~hit] == 0)
    for e in envs:
        e.close()


def test_compat_surface_matches_reference_tests():
    """The reference's own smoke tests (gym_swarm/tests/test_discrete_swarm.py:12-40) against SwarmEnv."""
    from gym_swarm_b200.envs import ContinuousSwarmEnv, SwarmEnv
    for cls in (SwarmEnv, ContinuousSwarmEnv):
        env = cls()
        env.reset()
        state, reward, done, info = env.step({0: 0, 1: 1, 2: 2, 3: 3})
        assert len(state) == env.num_agents
        assert set(reward.keys()) == {0, 1, 2, 3, "global"}
        assert set(reward["global"].keys()) == {"survival", "attraction", "repulsion", "alignment", "sum"}
        assert set(info.keys()) == {"predator_state", "predator_orientation", "predator_next_state"}
        env.set_env_parameters(num_agents=10, obs_space_size=200, verbose=False)
        env.reset()
        assert env.num_agents == 10 and env.obs_space_size == 200 and len(env.current_state) == 10
        env.set_reward_parameters(attraction_thresh=10, repulsion_thresh=200, predator_eat_rew=-50, verbose=False)
        env.reset()
        assert (env.attraction_thresh, env.repulsion_thresh, env.predator_eat_rew) == (10, 200, -50)
        with pytest.raises(KeyError):
            env.step({i: 11 for i in range(10)})
        env.close()


def test_gym_make_roundtrip(monkeypatch):
    """gym.make(id) -> the drop-in env (through a stand-in registry: gym itself is not in the image), one reference-style
    episode fragment through the registered entry point."""
    from test_abi import _stub_gym
    from gym_swarm_b200 import envs
    gym, _ = _stub_gym(monkeypatch)
    envs.register_gym()
    for id_ in ("Discrete-Swarm-v0", "Continuous-Swarm-v0"):
        env = gym.make(id_)
        state = env.reset()
        assert len(state) == env.num_agents == 4
        for _ in range(5):
            if env.done:
                env.reset()
            state, reward, done, info = env.step({0: 0, 1: 1, 2: 2, 3: 3})
            assert set(reward.keys()) == {0, 1, 2, 3, "global"}
        env.close()


def test_compat_env_lockstep_with_oracle():
    """SwarmEnv (dict API) reproduces the oracle step by step when fed the same tapes."""
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import SwarmEnv
    from oracle import swarm_oracle as so
    env = SwarmEnv(seed=5)
    twin = np.random.default_rng(5)            # mirrors the env's internal generator
    env.reset()
    N, S = 4, 20
    rt = tapes.make_reset_tapes(twin, 1, 1, N, S, KR=env._env.KR, KP=env._env.KP)
    ora = so.OracleEnv(N, S, 5, 2, -10.0)
    ora.reset(rt.place[0, 0], rt.ptape[0, 0])
    act_rng = np.random.default_rng(99)
    for t in range(200):
        if env.done:
            with pytest.raises(RuntimeError):
                env.step({i: 0 for i in range(N)})
            break
        action = {i: int(act_rng.integers(0, 9)) for i in range(N)}
        roll = twin.random()
        slab = twin.integers(0, 8, size=(1, env._env.K), dtype=np.uint8)
        state, reward, done, info = env.step(action)
        r = ora.step(np.array([action[i] for i in range(N)]), slab[0], roll < 0.1)
        assert all(tuple(state[i]) == (r["x"][i], r["y"][i]) for i in range(N))
        assert done == r["done"]
        assert tuple(info["predator_state"]) == tuple(r["pred_prev"])
        assert tuple(info["predator_next_state"]) == tuple(r["pred_next"])
        assert info["predator_orientation"] == r["pred_orient"]
        for q, k in enumerate(so.REWARD_KEYS):
            assert abs(reward["global"][k] - r["glob"][q]) <= 1e-6 * abs(r["glob"][q]) + 1e-9
            for i in range(N):
                assert abs(reward[i][k] - r["agent"][i, q]) <= 1e-6 * abs(r["agent"][i, q]) + 1e-9
    env.close()
