"""Records golden trajectories from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it drives E instances of the reference ``DiscreteSwarmEnv``
(loaded by ``oracle/reference_shim.py``; stable argsort; every ``np.random`` call
served from tapes) through T steps with auto-reset semantics (``AutoResetRef``:
when an episode ends the terminal step's outputs are recorded, then the
reference's own ``reset()`` is run from reset-tape slot ``n_resets % R``) and
stores tapes, actions and everything the reference returned in
``tests/golden/<case>.npz``.  The oracle (CPU tests) and the CUDA path (GPU tests)
must reproduce these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_shim as rs  # noqa: E402
import gym_swarm_b200  # noqa: E402,F401
from gym_swarm_b200 import tapes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: dict(E,N,S,att,rep,eat,T,R,K,KR,KP,n_actions,auto_reset,weights(schedule or tuple),vf,seed)
CASES = {
    # BASELINE configs[0]: reference defaults, seeded random actions, 1000 steps
    "default_n4_s20": dict(E=2, N=4, S=20, att=5, rep=2, eat=-10, T=1000, R=32, n_actions=8, auto_reset=True, seed=1),
    # configs[1] shape at reduced E
    "cfg2_n16_s32": dict(E=8, N=16, S=32, att=5, rep=2, eat=-10, T=250, R=16, n_actions=8, auto_reset=True, seed=2),
    # configs[2] shape: per-agent rewards + float curriculum weights
    "cfg3_n128_s128_curriculum": dict(E=2, N=128, S=128, att=5, rep=2, eat=-10, T=40, R=4, n_actions=8,
                                      auto_reset=True, seed=3, curriculum=True),
    "cfg3_n128_s128_vf5": dict(E=1, N=128, S=128, att=5, rep=2, eat=-10, T=20, R=4, n_actions=8,
                               auto_reset=True, seed=4, vf=5, weights=(1.0, 1.0, 0.5)),
    # configs[4] shape: tiny envs, many episodes, auto-reset
    "cfg5_n8_s16_autoreset": dict(E=16, N=8, S=16, att=5, rep=2, eat=-10, T=300, R=24, n_actions=8,
                                  auto_reset=True, seed=5),
    # dense grids: many collision re-draw rounds, action 8, vf gate
    "dense_n12_s6": dict(E=4, N=12, S=6, att=3, rep=1, eat=-5, T=150, R=64, K=256, KR=2000, KP=64, n_actions=9,
                         auto_reset=True, seed=6, vf=2),
    "dense_n40_s9": dict(E=2, N=40, S=9, att=5, rep=2, eat=-10, T=40, R=32, K=2048, KR=4000, KP=64, n_actions=9,
                         auto_reset=True, seed=7, weights=(True, False, True)),
    # odd thresholds (the reference's own test uses att=10, rep=200: XNOR branch, rep > S)
    "thresh_att10_rep200": dict(E=2, N=10, S=20, att=10, rep=200, eat=-50, T=120, R=8, n_actions=8,
                                auto_reset=True, seed=8),
    "thresh_rep_gt_att": dict(E=2, N=10, S=20, att=2, rep=5, eat=-50, T=120, R=8, n_actions=9, auto_reset=True, seed=9),
    "thresh_zero": dict(E=2, N=10, S=20, att=0, rep=0, eat=-1, T=80, R=8, n_actions=8, auto_reset=True, seed=10),
    "thresh_att_gt_s_vf7": dict(E=2, N=10, S=20, att=30, rep=3, eat=-50, T=80, R=8, n_actions=8, auto_reset=True,
                                seed=11, vf=7),
    # no auto-reset: envs freeze after termination
    "freeze_n8_s10": dict(E=8, N=8, S=10, att=5, rep=2, eat=-10, T=120, R=1, n_actions=8, auto_reset=False, seed=12),
    # odd N (not a multiple of the lane/vector widths)
    "odd_n37_s50": dict(E=3, N=37, S=50, att=6, rep=3, eat=-10, T=60, R=8, n_actions=9, auto_reset=True, seed=13,
                        weights=(0.7, 1.3, 0.2)),
    "odd_n200_s64": dict(E=1, N=200, S=64, att=5, rep=2, eat=-10, T=12, R=4, n_actions=8, auto_reset=True, seed=14),
    # large-N (CTA-per-env + tiled pairwise path): as large as the reference's O(N^2) Python loop allows
    "large_n600_s256": dict(E=1, N=600, S=256, att=5, rep=2, eat=-10, T=3, R=2, n_actions=8, auto_reset=True, seed=15),
}


def weights_at(case, t):
    if case.get("curriculum"):
        # reward curriculum (README.md:23): attraction on, repulsion ramps 0->1, alignment 0.5
        return (1.0, min(1.0, t / 20.0), 0.5)
    return tuple(case.get("weights", (1.0, 1.0, 1.0)))


def record(name, case):
    E, N, S, T, R = case["E"], case["N"], case["S"], case["T"], case["R"]
    K = case.get("K", tapes.default_slab_len(N))
    # pinned to the default that was in force when the fixtures were recorded (tapes.default_place_extra has since
    # grown to max(32, N // 4)); the value is stored in every fixture's meta, which is what the tests read
    KR = case.get("KR", int(max(16, N // 4)))
    KP = case.get("KP", 8)
    vf = case.get("vf", None)
    rng = np.random.default_rng(case["seed"])
    rt = tapes.make_reset_tapes(rng, R, E, N, S, KR=KR, KP=KP)
    st = tapes.make_step_tapes(rng, T, E, N, K=K, n_actions=case["n_actions"])
    envs = [rs.TapedReferenceEnv(N, S, case["att"], case["rep"], case["eat"]) for _ in range(E)]
    nreset = np.zeros(E, dtype=np.int64)

    g = dict(
        init_x=np.zeros((E, N), np.int16), init_y=np.zeros((E, N), np.int16), init_pred=np.zeros((E, 2), np.int16),
        init_target=np.zeros(E, np.int32),
        x=np.zeros((T, E, N), np.int16), y=np.zeros((T, E, N), np.int16),                 # state after step (post auto-reset)
        term_x=np.zeros((T, E, N), np.int16), term_y=np.zeros((T, E, N), np.int16),       # positions the step produced
        pred=np.zeros((T, E, 2), np.int16), target=np.zeros((T, E), np.int32),            # state after step (post auto-reset)
        pred_prev=np.zeros((T, E, 2), np.int16), pred_next=np.zeros((T, E, 2), np.int16),
        pred_orient=np.zeros((T, E), np.uint8), step_target=np.zeros((T, E), np.int32),
        done=np.zeros((T, E), np.uint8), eaten=np.full((T, E), -1, np.int32), frozen=np.zeros((T, E), np.uint8),
        cnt=np.zeros((T, E, 2), np.int64), glob=np.zeros((T, E, 5), np.float64),
        agent=np.zeros((T, E, N, 5), np.float64), agent_cnt=np.zeros((T, E, N, 2), np.int32),
        consumed=np.zeros((T, E), np.int32), weights=np.zeros((T, 3), np.float64),
    )

    def do_reset(e):
        slot = int(nreset[e] % R)
        nreset[e] += 1
        return envs[e].reset(rt.place[slot, e], rt.ptape[slot, e], rt.orient[slot, e])

    for e in range(E):
        r = do_reset(e)
        g["init_x"][e], g["init_y"][e], g["init_pred"][e], g["init_target"][e] = r["x"], r["y"], r["pred"], r["target"]
    frozen = np.zeros(E, bool)
    last = [None] * E
    for t in range(T):
        w = weights_at(case, t)
        g["weights"][t] = [float(v) for v in w]
        rtp = {"attraction": w[0], "repulsion": w[1], "alignment": w[2], "indiv_rewards": True, "vf_size": vf}
        for e in range(E):
            if frozen[e]:
                # reference raises RuntimeError (swarm_discrete_env.py:233-234); batch semantics: frozen
                try:
                    envs[e].step(st.actions[t, e], st.slab[t, e], st.roll[t, e], rtp)
                    raise AssertionError("reference did not raise on a finished episode")
                except RuntimeError:
                    pass
                g["frozen"][t, e] = 1
                g["done"][t, e] = 1
                for k in ("x", "y", "pred", "target"):
                    g[k][t, e] = g[k][t - 1, e]
                continue
            r = envs[e].step(st.actions[t, e], st.slab[t, e], st.roll[t, e], rtp)
            g["term_x"][t, e], g["term_y"][t, e] = r["x"], r["y"]
            g["pred_prev"][t, e], g["pred_next"][t, e] = r["pred_prev"], r["pred_next"]
            g["pred_orient"][t, e], g["step_target"][t, e] = r["pred_orient"], r["target"]
            g["done"][t, e], g["eaten"][t, e] = r["done"], r["eaten"]
            g["cnt"][t, e] = (r["cnt_rep"], r["cnt_att"])
            g["glob"][t, e], g["agent"][t, e], g["agent_cnt"][t, e] = r["glob"], r["agent"], r["agent_cnt"]
            g["consumed"][t, e] = r["consumed"]
            if r["done"] and case["auto_reset"]:
                rr = do_reset(e)
                g["x"][t, e], g["y"][t, e], g["pred"][t, e], g["target"][t, e] = rr["x"], rr["y"], rr["pred"], rr["target"]
            else:
                g["x"][t, e], g["y"][t, e], g["pred"][t, e], g["target"][t, e] = r["x"], r["y"], r["pred_next"], r["target"]
                if r["done"]:
                    frozen[e] = True
    meta = dict(E=E, N=N, S=S, att=case["att"], rep=case["rep"], eat=case["eat"], T=T, R=R, K=K, KR=KR, KP=KP,
                vf=-1 if vf is None else vf, auto_reset=int(case["auto_reset"]))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta=np.array([[k, str(v)] for k, v in meta.items()]),
        place=rt.place, ptape=rt.ptape, orient=rt.orient,
        actions=st.actions, retarget=st.retarget, slab=st.slab, **g)
    print(f"{name}: E={E} N={N} S={S} T={T} dones={int(g['done'].sum())} "
          f"redraws={int(g['consumed'].sum())} max_consumed={int(g['consumed'].max())}")


def main(argv):
    if not rs.reference_available():
        raise SystemExit("needs /root/reference (authoring container only)")
    names = argv[1:] or list(CASES)
    for n in names:
        record(n, CASES[n])


if __name__ == "__main__":
    main(sys.argv)
