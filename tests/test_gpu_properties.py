"""Size-independent properties at BASELINE.json's FULL shapes (where the oracle cannot run in seconds):
invariants of the reference transition that must hold for every environment, checked on the device with torch
reductions, plus cross-path checksums (two different kernel paths must produce identical tensors).

Properties (SDE = swarm_discrete_env.py):
  P1  positions stay on the grid and every env's cells are pairwise distinct (SDE:246-255 re-draw loop)
  P2  every agent moved by at most one cell per axis on the torus (SDE:60-80), unless its env was just reset
  P3  the predator moved by at most one cell per axis on the torus (SDE:142-165); its target is a valid index
  P4  done <=> the predator's new cell holds an agent; `eaten` is that agent (SDE:307-309, 321-327)
  P5  reward on termination is exactly (eat, 0, 0, 0, eat); otherwise survival == 0 (SDE:321-327, 363-366)
  P6  integer identities: global counts == sum of the per-agent counts (SDE:338/372, 343/373)
  P7  episode statistics: finished episodes == number of done flags seen; lengths add up to the steps taken
  P8  no sticky error bit
  P10 integer reward counts, exactly, on a strided subset of envs (first and last included): global and per-agent
      repulsion / attraction counts recomputed from the new positions with the literal N x N formulas of SDE:330-343
  P9  predator, exactly, for every env: on a retarget step the target is the canonical nearest
      agent of the PRE-move positions (SDE:111-131: lowest index at the minimum Chebyshev distance, the highest one when
      lowest != highest and 2 d_min >= S), otherwise it is unchanged; the predator then moves one cell towards the
      target's PRE-move cell, both axes negated when the signed max offset exceeds S // 2 (SDE:142-165)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torus_step_ok(a, b, S):
    d = (b.int() - a.int()).abs()
    return (d <= 1) | (d == S - 1)


def _literal_counts(nx, ny, S, rep, att, rows=1024):
    """SDE:330-343 / 371-373 on [n, N] position tensors -> per-agent [n, N, 2] and global [n, 2] integer counts."""
    import torch
    n, N = nx.shape
    per_agent = torch.zeros((n, N, 2), dtype=torch.int64, device=nx.device)
    for r0 in range(0, N, rows):
        xr, yr = nx[:, r0:r0 + rows, None].int(), ny[:, r0:r0 + rows, None].int()
        D = torch.maximum((xr - nx[:, None, :].int()).abs(), (yr - ny[:, None, :].int()).abs())
        D2 = S - D
        R = (D < rep).int() + (D2 < rep).int()
        A = ((D < att) == (D >= rep)).int() + ((D2 < att) == (D2 >= rep)).int()
        per_agent[:, r0:r0 + rows, 0] = -(R.sum(2, dtype=torch.int64) - 1)
        per_agent[:, r0:r0 + rows, 1] = A.sum(2, dtype=torch.int64)
    glob = per_agent.sum(1)
    return per_agent, glob


def _run_properties(E, N, S, T, per_agent, path="auto", seed=0):
    import torch
    from gym_swarm_b200.envs import BatchedSwarmEnv
    dev = torch.device("cuda:0")
    env = BatchedSwarmEnv(E, N, S, auto_reset=True, track_stats=True, reset_slots=2, per_agent=per_agent,
                          debug_outputs=per_agent, path=path, seed=seed)
    g = torch.Generator(device=dev).manual_seed(seed)
    place = torch.randint(0, S, (env.R, E, env.PL), dtype=torch.int16, device=dev, generator=g)
    ptape = torch.randint(0, S, (env.R, E, env.KP, 2), dtype=torch.int16, device=dev, generator=g)
    env.set_reset_tapes(place, ptape)
    env.reset()
    dones_seen = torch.zeros((), dtype=torch.int64, device=dev)
    big = N > 2048
    for t in range(T):
        x0, y0 = env.x.clone(), env.y.clone()
        px0, py0 = env.px.clone(), env.py.clone()
        tgt0 = env.target.clone()
        actions = torch.randint(0, 8, (E, N), dtype=torch.uint8, device=dev, generator=g)
        retarget = (torch.rand((E,), device=dev, generator=g) < 0.1).to(torch.uint8)
        slab = torch.randint(0, 8, (E, env.K), dtype=torch.uint8, device=dev, generator=g)
        w = {"attraction": 1.0, "repulsion": 0.25 * t, "alignment": 0.5, "vf_size": None}
        (x, y), reward, done, info = env.step(actions, w, retarget=retarget, slab=slab)
        done_b = done.bool()
        live = ~done_b
        # P8
        assert int((env.err != 0).sum()) == 0, "sticky error bits"
        # P1 (post auto-reset state as well)
        assert int(x.min()) >= 0 and int(x.max()) < S and int(y.min()) >= 0 and int(y.max()) < S
        cell = x.int() * S + y.int()
        srt = cell.sort(dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "two agents share a cell"
        # P2 / P3 on the envs that were not reset in this step; terminal positions for the others
        tx, ty = info["terminal_state"]
        nx = torch.where(done_b[:, None], tx, x)
        ny = torch.where(done_b[:, None], ty, y)
        assert bool(_torus_step_ok(x0, nx, S).all()) and bool(_torus_step_ok(y0, ny, S).all())
        pn = info["predator_next_state"]
        assert bool(torch.equal(info["predator_state"][:, 0], px0)) and bool(torch.equal(info["predator_state"][:, 1], py0))
        assert bool(_torus_step_ok(px0, pn[:, 0], S).all()) and bool(_torus_step_ok(py0, pn[:, 1], S).all())
        tgt = info["predator_target"]
        assert int(tgt.min()) >= 0 and int(tgt.max()) < N
        # P9
        d0 = torch.maximum((px0[:, None].int() - x0.int()).abs(), (py0[:, None].int() - y0.int()).abs())
        dmin = d0.min(dim=1).values
        at_min = d0 == dmin[:, None]
        lo = at_min.int().argmax(dim=1)
        hi = (N - 1) - at_min.flip(1).int().argmax(dim=1)
        nearest = torch.where((lo == hi) | (2 * dmin < S), lo, hi)
        want_tgt = torch.where(retarget.bool(), nearest, tgt0.long())
        assert bool(torch.equal(tgt.long(), want_tgt)), "predator target"     # (the step's target, also for done envs)
        step_tgt = tgt.long()
        cdx = px0.int() - x0.gather(1, step_tgt[:, None]).squeeze(1).int()
        cdy = py0.int() - y0.gather(1, step_tgt[:, None]).squeeze(1).int()
        flip = torch.where(torch.maximum(cdx, cdy) > S // 2, -1, 1)
        want_px = torch.remainder(px0.int() - torch.sign(cdx) * flip, S)
        want_py = torch.remainder(py0.int() - torch.sign(cdy) * flip, S)
        assert bool(torch.equal(pn[:, 0].int(), want_px)) and bool(torch.equal(pn[:, 1].int(), want_py)), "predator move"
        del d0, at_min
        # P4
        hit = (nx == pn[:, 0:1]) & (ny == pn[:, 1:2])
        assert bool(torch.equal(hit.any(dim=1), done_b))
        eaten = info["eaten"]
        assert bool((eaten[live] == -1).all())
        if int(done_b.sum()) > 0:
            rows = done_b.nonzero().squeeze(1)
            assert bool(hit[rows, eaten[rows].long()].all())
            first = hit[rows].int().argmax(dim=1)
            assert bool(torch.equal(first.int(), eaten[rows]))
        # P5
        rg = reward["global"]
        assert bool((rg[done_b] == torch.tensor([-10.0, 0, 0, 0, -10.0], device=dev)).all())
        assert bool((rg[live][:, 0] == 0).all())
        assert bool(torch.isfinite(rg).all())
        if not big:
            assert bool((rg[live][:, 4] - rg[live][:, 1:4].sum(dim=1)).abs().max() < 1e-5)
        # P6
        if per_agent:
            ac = env.agent_counts
            assert bool(torch.equal(ac.sum(dim=1, dtype=torch.int64), reward["counts"].long()))
            ar = reward["agents"]
            assert bool((ar[done_b][:, :, :3] == 0).all())
        # P10
        sel = torch.unique(torch.linspace(0, E - 1, 256 if N <= 256 else 2, device=dev).long())
        sel = sel[live[sel]]
        if sel.numel() > 0:
            pa, gl = _literal_counts(nx[sel], ny[sel], S, 2, 5)
            assert bool(torch.equal(reward["counts"][sel].long(), gl)), "global counts"
            if per_agent:
                assert bool(torch.equal(env.agent_counts[sel].long(), pa)), "per-agent counts"
        dones_seen += done_b.sum()
    # P7
    st = env.stats()
    assert st["episodes"] == float(dones_seen)
    assert st["sum_length"] + st["running_steps"] == float(E) * T
    assert st["sum_survival"] == -10.0 * st["episodes"]
    out = {k: v.clone() for k, v in dict(x=env.x, y=env.y, px=env.px, py=env.py, target=env.target, reward=env.reward,
                                         counts=env.counts, done=env.done).items()}
    if per_agent:
        out["agent_counts"] = env.agent_counts.clone()
        out["agent_reward"] = env.agent_reward.clone()
    env.close()
    return out


def test_properties_cfg3_full_size():
    """configs[2]: 65,536 envs x 128 agents on 128^2, per-agent rewards + curriculum weights."""
    _run_properties(65536, 128, 128, T=6, per_agent=True)


def test_properties_cfg5_full_size():
    """configs[4]: 1,048,576 envs x 8 agents on 16^2, auto-reset + episode statistics (thread-per-env kernel)."""
    _run_properties(1 << 20, 8, 16, T=8, per_agent=False)


def test_properties_cfg4_full_size():
    """configs[3]: 16 envs x 16,384 agents on 1024^2 (staged kernels, box-count reward)."""
    _run_properties(16, 16384, 1024, T=4, per_agent=True)


def test_properties_cfg2_full_size():
    """configs[1]: 4,096 envs x 16 agents on 32^2, global reward."""
    _run_properties(4096, 16, 32, T=40, per_agent=False)


@pytest.mark.parametrize("E,N,S,paths", [
    (65536, 128, 128, ("auto", "staged", "staged_pairwise")),     # cfg3: fused ring sweep == staged box counts == staged pairwise
    (1 << 18, 8, 16, ("auto", "group", "staged")),                # cfg5 shape: thread per env == lane group == staged
    (16, 16384, 1024, ("auto", "staged_pairwise")),               # cfg4: box counts == symmetric tiled pairwise
    (4096, 16, 32, ("auto", "staged")),                           # cfg2
    (1 << 17, 16, 32, ("auto", "group")),                         # big batch of 16-agent envs: tiny<16> == lane group
    (1 << 16, 11, 20, ("auto", "group")),
])
def test_kernel_paths_agree_at_full_size(E, N, S, paths):
    """Same seeds through different kernel paths: every integer state / output tensor identical; float rewards (the same
    integers combined in fp64, rounded to f32) within the north-star tolerance 1e-6 relative."""
    import torch
    ref = None
    for path in paths:
        out = _run_properties(E, N, S, T=3, per_agent=True, path=path, seed=5)
        if ref is None:
            ref = out
            continue
        for k in ref:
            if ref[k].is_floating_point():
                # two kernel paths form the same integers and combine them in fp64: identical up to the f32 store
                assert torch.allclose(ref[k], out[k], rtol=1e-6, atol=0.0), f"{k}: path {paths[0]} != {path}"
            else:
                assert torch.equal(ref[k], out[k]), f"{k}: path {paths[0]} != {path}"
