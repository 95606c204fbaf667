"""Shared replay/compare logic: a fixture recorded from the reference
(tests/golden/*.npz) is replayed through an implementation exposing the batch
interface (``reset()`` / ``step(actions, slab, retarget, weights, vf, want_agents)``
returning dicts of numpy arrays) and every recorded quantity is compared."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# float rewards: north-star tolerance (1e-6 RELATIVE); counts/positions/flags: bit-exact
RTOL = 1e-6
ZERO = 1e-12    # the reference's float64 sums return its exact zeros as +-1e-16: below this a reference value IS zero


def assert_float_parity(got, want, rtol=RTOL, err_msg=""):
    """|got - want| <= rtol * |want| for every entry; where the reference value is (numerically) zero the result must
    be zero as well.  No absolute slack anywhere else."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{err_msg}: shape {got.shape} != {want.shape}"
    zero = np.abs(want) < ZERO
    bad = np.where(zero, np.abs(got) >= ZERO, np.abs(got - want) > rtol * np.abs(want))
    if bad.any():
        i = tuple(int(v[0]) for v in np.nonzero(bad))
        raise AssertionError(f"{err_msg}: {int(bad.sum())} of {bad.size} entries beyond rtol={rtol}; first at {i}: "
                             f"got {got[i]!r}, reference {want[i]!r}")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = {k: v for k, v in z["meta"]}
    cfg = dict(E=int(meta["E"]), N=int(meta["N"]), S=int(meta["S"]), att=int(meta["att"]), rep=int(meta["rep"]),
               eat=float(meta["eat"]), T=int(meta["T"]), R=int(meta["R"]), K=int(meta["K"]), KR=int(meta["KR"]),
               KP=int(meta["KP"]), vf=None if int(meta["vf"]) < 0 else int(meta["vf"]),
               auto_reset=bool(int(meta["auto_reset"])))
    return cfg, z


def replay(impl, cfg, z, steps=None, check_agents=True, float_rtol=RTOL):
    """impl: object with reset()/step(); raises AssertionError with context on the first mismatch."""
    T = cfg["T"] if steps is None else min(steps, cfg["T"])
    st = impl.reset()
    np.testing.assert_array_equal(st["x"], z["init_x"], err_msg="reset x")
    np.testing.assert_array_equal(st["y"], z["init_y"], err_msg="reset y")
    np.testing.assert_array_equal(np.stack([st["px"], st["py"]], 1), z["init_pred"], err_msg="reset predator")
    np.testing.assert_array_equal(st["target"], z["init_target"], err_msg="reset target")
    for t in range(T):
        w = tuple(float(v) for v in z["weights"][t])
        o = impl.step(z["actions"][t], z["slab"][t], z["retarget"][t], w, cfg["vf"], check_agents)
        live = z["frozen"][t] == 0
        ctx = f"step {t}"
        np.testing.assert_array_equal(o["done"], z["done"][t], err_msg=ctx + " done")
        np.testing.assert_array_equal(o["x"], z["x"][t], err_msg=ctx + " x")
        np.testing.assert_array_equal(o["y"], z["y"][t], err_msg=ctx + " y")
        np.testing.assert_array_equal(np.stack([o["px"], o["py"]], 1), z["pred"][t], err_msg=ctx + " predator")
        np.testing.assert_array_equal(o["target"], z["target"][t], err_msg=ctx + " target")
        np.testing.assert_array_equal(o["eaten"][live], z["eaten"][t][live], err_msg=ctx + " eaten")
        np.testing.assert_array_equal(o["pred_prev"][live], z["pred_prev"][t][live], err_msg=ctx + " pred_prev")
        np.testing.assert_array_equal(o["pred_next"][live], z["pred_next"][t][live], err_msg=ctx + " pred_next")
        np.testing.assert_array_equal(o["pred_orient"][live], z["pred_orient"][t][live], err_msg=ctx + " pred_orient")
        np.testing.assert_array_equal(o["step_target"][live], z["step_target"][t][live], err_msg=ctx + " step_target")
        np.testing.assert_array_equal(o["consumed"][live], z["consumed"][t][live], err_msg=ctx + " consumed draws")
        np.testing.assert_array_equal(o["cnt"][live], z["cnt"][t][live], err_msg=ctx + " integer reward counts")
        assert_float_parity(o["glob"][live], z["glob"][t][live], rtol=float_rtol, err_msg=ctx + " global reward")
        dn = (z["done"][t] == 1) & live
        np.testing.assert_array_equal(o["term_x"][dn], z["term_x"][t][dn], err_msg=ctx + " terminal x")
        np.testing.assert_array_equal(o["term_y"][dn], z["term_y"][t][dn], err_msg=ctx + " terminal y")
        if check_agents:
            np.testing.assert_array_equal(o["agent_cnt"][live], z["agent_cnt"][t][live],
                                          err_msg=ctx + " per-agent integer counts")
            assert_float_parity(o["agent"][live], z["agent"][t][live], rtol=float_rtol,
                                       err_msg=ctx + " per-agent reward")
        assert not np.any(o["err"] & ~np.uint32(16)), ctx + f" error flags {o['err']}"
    return T
