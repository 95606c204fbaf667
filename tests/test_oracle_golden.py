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
