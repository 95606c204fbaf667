"""Standalone C-ABI kernels (swarm_move / swarm_predator / swarm_pairwise) against the oracle's functions."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("E,N,S", [(1, 1, 2), (7, 5, 3), (64, 8, 16), (33, 37, 50), (1024, 128, 128), (3, 16384, 1024)])
def test_swarm_move_matches_reference_rule(E, N, S):
    import torch
    from gym_swarm_b200 import _lib
    from oracle import swarm_oracle as so
    lib = _lib.load()
    rng = np.random.default_rng(E * 31 + N)
    x = rng.integers(0, S, size=(E, N), dtype=np.int16)
    y = rng.integers(0, S, size=(E, N), dtype=np.int16)
    a = rng.integers(0, 9, size=(E, N), dtype=np.uint8)
    dx, dy, da = (torch.from_numpy(v).cuda() for v in (x, y, a))
    ox, oy = torch.empty_like(dx), torch.empty_like(dy)
    _lib.check(lib.swarm_move(E, N, S, _ptr(dx), _ptr(dy), _ptr(da), _ptr(ox), _ptr(oy), _stream()))
    mv = so.ACTION_MOVES[a.astype(np.int64)]
    np.testing.assert_array_equal(ox.cpu().numpy(), so.wrap(x.astype(np.int64) + mv[..., 0], S))
    np.testing.assert_array_equal(oy.cpu().numpy(), so.wrap(y.astype(np.int64) + mv[..., 1], S))


@pytest.mark.parametrize("E,N,S", [(200, 2, 4), (300, 8, 16), (100, 16, 8), (50, 37, 50), (64, 40, 9), (40, 128, 128),
                                   (9, 200, 64), (5, 512, 100), (3, 4096, 256), (2, 16384, 1024), (700, 16384, 300),
                                   # the cp.async stream (N <= 256, E % 4 == 0): every <LPE, V> shape, ragged last tiles,
                                   # several tiles per warp, more retargeting envs than row slots (50 % retarget here)
                                   (3004, 8, 16), (1500, 16, 12), (900, 24, 16), (644, 32, 20), (420, 48, 30),
                                   (372, 64, 33), (204, 104, 64), (260, 128, 128), (124, 136, 90), (100, 256, 200),
                                   (52, 248, 2048),
                                   # the CTA-per-env cp.async stream (N >= 4096, E % 4 == 0): one chunk per pass, several,
                                   # a ragged last chunk, more envs than CTAs in flight
                                   (8, 4096, 256), (12, 16384, 1024), (4, 40000, 2048), (900, 4096, 100)])
def test_swarm_predator_matches_oracle(E, N, S):
    """Every dispatch of the predator kernel (lane groups 1..32, CTA per env, scalar fallback): retarget tie rule,
    pursuit on the OLD positions, wrap, termination test on the NEW positions."""
    import torch
    from gym_swarm_b200 import _lib
    from oracle import swarm_oracle as so
    if E * N > 4_000_000:
        E = max(1, 4_000_000 // N)
    lib = _lib.load()
    rng = np.random.default_rng(E + N + S)
    xo = rng.integers(0, S, size=(E, N), dtype=np.int16)
    yo = rng.integers(0, S, size=(E, N), dtype=np.int16)
    act = rng.integers(0, 9, size=(E, N))
    mv = so.ACTION_MOVES[act]
    xn = so.wrap(xo.astype(np.int64) + mv[..., 0], S).astype(np.int16)
    yn = so.wrap(yo.astype(np.int64) + mv[..., 1], S).astype(np.int16)
    ret = (rng.random(E) < 0.5).astype(np.uint8)
    tgt = rng.integers(0, N, size=E, dtype=np.int32)
    # predator next to its target in half of the envs so that terminations happen
    px = np.where(rng.random(E) < 0.5, np.clip(xo[np.arange(E), tgt] + rng.integers(-2, 3, size=E), 0, S - 1),
                  rng.integers(0, S, size=E)).astype(np.int16)
    py = np.where(rng.random(E) < 0.5, np.clip(yo[np.arange(E), tgt] + rng.integers(-2, 3, size=E), 0, S - 1),
                  rng.integers(0, S, size=E)).astype(np.int16)
    # precondition of closest_target (SDE:111-131): the predator never stands on an agent's (old) cell
    for _ in range(1000):
        clash = ((xo == px[:, None]) & (yo == py[:, None])).any(axis=1)
        if not clash.any():
            break
        px[clash] = rng.integers(0, S, size=int(clash.sum()))
        py[clash] = rng.integers(0, S, size=int(clash.sum()))
    assert not clash.any()
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(xo=xo, yo=yo, xn=xn, yn=yn, ret=ret, px=px.copy(), py=py.copy(),
                                                        tgt=tgt.copy()).items()}
    ori = torch.zeros(E, dtype=torch.uint8, device="cuda")
    prev = torch.zeros((E, 2), dtype=torch.int16, device="cuda")
    done = torch.zeros(E, dtype=torch.uint8, device="cuda")
    eaten = torch.zeros(E, dtype=torch.int32, device="cuda")
    # second run with the cached target cell (swarm_state.tcell) instead of the dependent read of the old positions
    d2 = {k: v.clone() for k, v in d.items()}
    ori2, prev2, done2, eaten2 = torch.zeros_like(ori), torch.zeros_like(prev), torch.zeros_like(done), torch.zeros_like(eaten)
    ar = np.arange(E)
    tcell = torch.from_numpy(np.stack([xo[ar, tgt], yo[ar, tgt]], 1).astype(np.int16)).cuda()
    _lib.check(lib.swarm_predator(E, N, S, _ptr(d["xo"]), _ptr(d["yo"]), _ptr(d["xn"]), _ptr(d["yn"]), _ptr(d["ret"]),
                                  _ptr(d["px"]), _ptr(d["py"]), _ptr(d["tgt"]), _ptr(ori), _ptr(prev), _ptr(done),
                                  _ptr(eaten), None, _stream()))
    _lib.check(lib.swarm_predator(E, N, S, _ptr(d2["xo"]), _ptr(d2["yo"]), _ptr(d2["xn"]), _ptr(d2["yn"]), _ptr(d2["ret"]),
                                  _ptr(d2["px"]), _ptr(d2["py"]), _ptr(d2["tgt"]), _ptr(ori2), _ptr(prev2), _ptr(done2),
                                  _ptr(eaten2), _ptr(tcell), _stream()))
    for a, b in ((d["px"], d2["px"]), (d["py"], d2["py"]), (d["tgt"], d2["tgt"]), (ori, ori2), (prev, prev2), (done, done2),
                 (eaten, eaten2)):
        assert torch.equal(a, b), "target-cell cache changed the predator step"
    t_new = d2["tgt"].long().cpu().numpy()
    np.testing.assert_array_equal(tcell.cpu().numpy(), np.stack([xn[ar, t_new], yn[ar, t_new]], 1),
                                  err_msg="tcell must hold the target's NEW cell after the step")
    n_done = 0
    for e in range(E):
        npx, npy, t, o = so.follow_target(int(px[e]), int(py[e]), int(tgt[e]), 0, xo[e].astype(np.int64),
                                          yo[e].astype(np.int64), S, bool(ret[e]))
        hit = np.nonzero((xn[e] == npx) & (yn[e] == npy))[0]
        assert (int(d["px"][e]), int(d["py"][e]), int(d["tgt"][e]), int(ori[e])) == (npx, npy, t, o), f"env {e}"
        assert tuple(prev[e].tolist()) == (int(px[e]), int(py[e]))
        assert int(done[e]) == (1 if len(hit) else 0) and int(eaten[e]) == (int(hit[0]) if len(hit) else -1), f"env {e}"
        n_done += len(hit) > 0
    if S * S < 4 * N:
        assert n_done > 0


@pytest.mark.parametrize("E,N,S,att,rep", [(6, 2, 4, 5, 2), (5, 37, 50, 5, 2), (4, 128, 128, 5, 2), (3, 600, 64, 9, 3),
                                           (2, 2048, 1024, 5, 2), (3, 100, 20, 10, 200), (3, 100, 20, 2, 5),
                                           (2, 50, 20, 0, 0), (2, 50, 20, 30, 3)])
def test_swarm_pairwise_matches_oracle(E, N, S, att, rep):
    import torch
    from gym_swarm_b200 import _lib
    from oracle import swarm_oracle as so
    lib = _lib.load()
    rng = np.random.default_rng(N * 7 + S)
    x = rng.integers(0, S, size=(E, N), dtype=np.int16)
    y = rng.integers(0, S, size=(E, N), dtype=np.int16)
    skip = np.zeros(E, np.uint8)
    skip[E - 1] = 1
    dx, dy, ds = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(skip).cuda()
    out = torch.full((E, N, 2), -7, dtype=torch.int32, device="cuda")
    nbytes = int(lib.swarm_pairwise_scratch_bytes(E, N))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(lib.swarm_pairwise(E, N, S, att, rep, _ptr(dx), _ptr(dy), _ptr(ds), _ptr(out), _ptr(scratch), nbytes,
                                  _stream()))
    got = out.cpu().numpy()
    eff = np.zeros(N, dtype=np.int64)
    for e in range(E):
        if skip[e]:
            assert not got[e].any()
            continue
        r = so.swarm_reward(x[e].astype(np.int64), y[e].astype(np.int64), eff, -5, -5, N, S, att, rep, -10.0,
                            (1.0, 1.0, 1.0), None)
        np.testing.assert_array_equal(got[e], r["agent_cnt"], err_msg=f"env {e}")


@pytest.mark.parametrize("E,N,S,vf", [(5, 9, 12, 2), (3, 4, 20, 0), (7, 8, 6, 1), (4, 128, 40, 2), (2, 300, 30, 3), (2, 130, 16, 7),
                                      (1, 2048, 128, 2), (2, 50, 1500, 3), (3, 37, 50, 4),
                                      # warp-per-env kernel (N <= 256, small bitmap): several envs per warp, every vf
                                      (70, 128, 128, 2), (41, 256, 64, 5), (33, 100, 128, 6), (50, 16, 32, 3), (90, 64, 33, 1)])
def test_visual_field_observation(E, N, S, vf):
    """swarm_observe_vf / BatchedSwarmEnv.observe_visual_field against the oracle's definition (parity unpinned: the
    reference tree has only the README figure for this output)."""
    from gym_swarm_b200 import tapes
    from gym_swarm_b200.envs import BatchedSwarmEnv
    from oracle import swarm_oracle as so
    rng = np.random.default_rng(N + vf)
    env = BatchedSwarmEnv(E, N, S, reset_slots=2, place_extra=max(64, 4 * N), pred_attempts=32)
    rt = tapes.make_reset_tapes(rng, 2, E, N, S, KR=env.KR, KP=env.KP)
    env.set_reset_tapes(rt.place, rt.ptape)
    env.reset()
    for t in range(3):
        if t:
            env.step(rng.integers(0, 8, size=(E, N), dtype=np.uint8))
        obs = env.observe_visual_field(vf).cpu().numpy()
        st = env.state_numpy()
        for e in range(E):
            ref = so.visual_field(st["x"][e], st["y"][e], st["px"][e], st["py"][e], S, vf)
            np.testing.assert_array_equal(obs[e], ref, err_msg=f"t={t} env {e}")
        assert (obs[:, :, vf, vf] >= 1).all()          # the centre is the agent itself (or the predator on top of it)
    env.close()
