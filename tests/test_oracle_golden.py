"""Pins the oracle: the NumPy restatement (oracle/swarm_oracle.py) must reproduce every golden
trajectory recorded from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from golden_utils import case_names, load_case, replay
from oracle import swarm_oracle as so


class OracleImpl:
    def __init__(self, cfg, z):
        self.b = so.OracleBatch(cfg["E"], cfg["N"], cfg["S"], cfg["att"], cfg["rep"], cfg["eat"], cfg["auto_reset"])
        self.b.set_reset_tapes(z["place"], z["ptape"])

    def reset(self):
        return self.b.reset()

    def step(self, actions, slab, retarget, weights, vf, want_agents):
        return self.b.step(actions, slab, retarget, weights, vf, want_agents)


def test_fixtures_present():
    assert len(case_names()) >= 15


@pytest.mark.parametrize("name", case_names())
def test_numpy_oracle_reproduces_reference(name):
    cfg, z = load_case(name)
    replay(OracleImpl(cfg, z), cfg, z, float_rtol=1e-12)


def test_redraw_slot_formula():
    # SDE:269-292 + :250-252: ch_id lists agent i (c_i - 1) times; the last draw wins
    c = np.array([1, 3, 2, 3, 1, 2, 3])
    slot, L = so.redraw_slots(c)
    ch_id = [i for i in range(len(c)) for _ in range(c[i] - 1)]
    assert L == len(ch_id)
    last = {}
    for k, i in enumerate(ch_id):
        last[i] = k
    for i in range(len(c)):
        if c[i] > 1:
            assert slot[i] == last[i]


class COracleImpl:
    def __init__(self, cfg, z):
        from oracle import c_oracle
        self.b = c_oracle.COracleBatch(cfg["E"], cfg["N"], cfg["S"], cfg["att"], cfg["rep"], cfg["eat"],
                                       cfg["auto_reset"], K=cfg["K"])
        self.b.set_reset_tapes(z["place"], z["ptape"])

    def reset(self):
        return self.b.reset()

    def step(self, actions, slab, retarget, weights, vf, want_agents):
        return self.b.step(actions, slab, retarget, weights, vf, want_agents)


@pytest.mark.parametrize("name", case_names())
def test_c_oracle_reproduces_reference(name):
    cfg, z = load_case(name)
    impl = COracleImpl(cfg, z)
    replay(impl, cfg, z, float_rtol=1e-12)
    impl.b.close()


def test_c_oracle_equals_numpy_oracle_on_fresh_tapes():
    from gym_swarm_b200 import tapes
    from oracle import c_oracle
    for (E, N, S, T, vf) in [(6, 24, 20, 40, None), (3, 70, 30, 10, 4), (2, 300, 90, 3, None)]:
        rng = np.random.default_rng(N)
        rt = tapes.make_reset_tapes(rng, 3, E, N, S)
        st = tapes.make_step_tapes(rng, T, E, N, n_actions=9)
        a = so.OracleBatch(E, N, S, 5, 2, -10.0, True)
        b = c_oracle.COracleBatch(E, N, S, 5, 2, -10.0, True, K=st.slab.shape[2])
        a.set_reset_tapes(rt.place, rt.ptape)
        b.set_reset_tapes(rt.place, rt.ptape)
        sa, sb = a.reset(), b.reset()
        for k in ("x", "y", "px", "py", "target"):
            np.testing.assert_array_equal(sa[k], sb[k])
        for t in range(T):
            oa = a.step(st.actions[t], st.slab[t], st.retarget[t], (0.3, 1.0, 0.6), vf)
            ob = b.step(st.actions[t], st.slab[t], st.retarget[t], (0.3, 1.0, 0.6), vf)
            for k in ("x", "y", "px", "py", "target", "done", "eaten", "cnt", "agent_cnt", "consumed", "eff"):
                np.testing.assert_array_equal(oa[k], ob[k], err_msg=f"{k} step {t}")
            np.testing.assert_allclose(oa["glob"], ob["glob"], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(oa["agent"], ob["agent"], rtol=1e-12, atol=1e-13)
        b.close()


def test_philox_known_answers():
    """The NumPy twin of the in-kernel tape generator against Random123's published Philox4x32-10 vectors."""
    from gym_swarm_b200 import tapes
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in tapes.philox4x32_10(*ctr, *key)) == out
    r, s = tapes.philox_step_tapes(5, 3, 4000, 40)
    assert 0.07 < r.mean() < 0.13 and s.max() <= 7
    rt = tapes.philox_reset_tapes(7, 2, 50, 8, 16)
    assert rt.place.min() >= 0 and rt.place.max() < 16 and rt.ptape.shape == (2, 50, 8, 2)
