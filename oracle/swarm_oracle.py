"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the swarm transition.

A NumPy restatement of the reference algorithm in
``/root/reference/gym_swarm/envs/swarm_discrete_env.py`` (abbreviated ``SDE``
below).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline may import this module; the product path (``gym-swarm_b200``) never
does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference's own tests hold no golden vectors
(SURVEY.md 4), so the oracle is pinned against trajectories recorded from the
unmodified reference file executed under ``oracle/reference_shim.py`` with the
same tapes (``tests/golden/*.npz``, generator ``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` replays every fixture through this module.

Randomness is never drawn here: every random quantity the reference draws from
``np.random`` arrives as an explicit tape (DESIGN.md "Tapes"):

* step:  ``slab``  u8[K]   -- k-th collision re-draw action of this step (SDE:250)
         ``retarget`` bool -- ``np.random.random() < 0.1`` (SDE:138-139)
* reset: ``place`` int[2N+2KR] -- values in consumption order: N x's, N y's, then
         per re-draw round L x's followed by L y's (SDE:202-203, 208-209)
         ``ptape`` int[KP,2]   -- predator placement attempts (SDE:92, 101)
         (the N render-only orientation draws of SDE:216 are not modelled)

Third-party arithmetic restated here (not under /root/reference; versions
installed when the fixtures were recorded: numpy 2.3.5, scikit-learn 1.9.0):
``DistanceMetric('chebyshev').pairwise`` = max(|dx|,|dy|);
``cosine_distances(X)`` = clip(1 - normalize(X) normalize(X)^T, 0, 2) with a zero
diagonal; ``np.argsort(kind='stable')``.
"""
from __future__ import annotations

import numpy as np

# SDE:467-475  action id -> (dx, dy); id 8 = no move (accepted by step()).
ACTION_MOVES = np.array([[-1, 0], [-1, -1], [0, -1], [1, -1], [1, 0], [1, 1], [0, 1], [-1, 1], [0, 0]],
                        dtype=np.int64)

# error bits (shared with include/swarm_abi.h)
ERR_SLAB_OVERFLOW = 1
ERR_PLACE_OVERFLOW = 2
ERR_PRED_OVERFLOW = 4
ERR_BAD_ACTION = 8
ERR_STEP_WHEN_DONE = 16

REWARD_KEYS = ("survival", "attraction", "repulsion", "alignment", "sum")


def wrap(p, S):
    """SDE:45-57 / SDE:60-80: one-step torus wrap (valid because |move| <= 1)."""
    p = np.where(p > S - 1, 0, p)
    return np.where(p < 0, S - 1, p)


def occupancy_counts(x, y, S):
    """c_i = number of agents (including i) standing on agent i's cell.

    ``change_position_id`` (SDE:269-292) returns ``None`` iff all c_i == 1, else
    the list with agent i repeated (c_i - 1) times in ascending i.
    """
    key = x.astype(np.int64) * int(S) + y.astype(np.int64)
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    return cnt[inv]


def redraw_slots(c):
    """Index (within the round's draw vector of length L = sum(c-1)) of the draw
    agent i ends up with: the LAST of its (c_i - 1) consecutive entries
    (SDE:250-252: list order, later writes win)."""
    extra = c - 1
    start = np.cumsum(extra) - extra          # exclusive scan
    return start + c - 2, int(extra.sum())


def resolve_collisions(x_old, y_old, actions, S, slab):
    """SDE:237-255: move, then re-draw colliding agents until all cells differ.

    Returns (eff_action, new_x, new_y, consumed, err)."""
    eff = np.array(actions, dtype=np.int64).copy()
    err = 0
    if np.any(eff > 8) or np.any(eff < 0):
        err |= ERR_BAD_ACTION
        eff = np.clip(eff, 0, 8)
    cur = 0
    while True:
        mv = ACTION_MOVES[eff]
        nx = wrap(x_old + mv[:, 0], S)
        ny = wrap(y_old + mv[:, 1], S)
        c = occupancy_counts(nx, ny, S)
        if np.all(c == 1):
            break
        slot, L = redraw_slots(c)
        if cur + L > len(slab):
            err |= ERR_SLAB_OVERFLOW
            break
        coll = c > 1
        eff[coll] = np.asarray(slab, dtype=np.int64)[cur + slot[coll]]
        cur += L
    return eff, nx, ny, cur, err


def closest_target(px, py, x, y, S):
    """SDE:111-131 with stable argsort (canonical tie rule).

    Literal restatement: distances row 0 of the (N+1) stacked points."""
    d = np.maximum(np.abs(np.concatenate(([px], x)) - px), np.abs(np.concatenate(([py], y)) - py)).astype(np.float64)
    id1 = np.argsort(d, kind="stable")[1]
    id2 = np.flip(np.argsort(S - d, kind="stable"))[1]
    if id1 == id2:
        t = id1
    else:
        d1 = d[id1]
        d2 = S - d[id2]
        t = id1 if d1 < d2 else id2
    return int(t) - 1


_MOVE_TO_ORIENT = {tuple(m): a for a, m in enumerate(ACTION_MOVES.tolist())}


def follow_target(px, py, target, orient, x_old, y_old, S, retarget):
    """SDE:133-165. ``x_old,y_old`` are the agents' PRE-move positions."""
    if retarget:
        target = closest_target(px, py, x_old, y_old, S)
    cd = (px - int(x_old[target]), py - int(y_old[target]))
    mv = [0, 0]
    for i in range(2):
        if cd[i] >= 1:
            mv[i] = -1
        elif cd[i] <= -1:
            mv[i] = 1
    if max(cd) > S // 2:            # SDE:152-154: SIGNED max, floor(S/2)
        mv = [-mv[0], -mv[1]]
    orient = _MOVE_TO_ORIENT[(mv[0], mv[1])]
    npx = int(wrap(np.int64(px + mv[0]), S))
    npy = int(wrap(np.int64(py + mv[1]), S))
    return npx, npy, target, orient


def _cosine_half(eff):
    """cosine_distances(moves)/2 (SDE:346-347), restated from scikit-learn:
    normalize rows (zero rows stay zero), 1 - X X^T, clip to [0,2], zero diagonal."""
    mv = ACTION_MOVES[eff].astype(np.float64)
    nrm = np.sqrt((mv * mv).sum(axis=1))
    nrm[nrm == 0.0] = 1.0
    xn = mv / nrm[:, None]
    return xn


def swarm_reward(x, y, eff, px, py, N, S, att, rep, eat, weights, vf, want_agents=True, block=1024):
    """SDE:294-380.  Returns dict(done, eaten, glob[5], cnt_rep, cnt_att,
    agent[N,5], agent_cnt[N,2]) -- integer counts are the un-weighted
    ``rew_rep`` / ``rew_attr`` (and their per-agent rows)."""
    w_att, w_rep, w_ali = weights
    out = dict(done=False, eaten=-1, glob=np.zeros(5), cnt_rep=0, cnt_att=0,
               agent=np.zeros((N, 5)) if want_agents else None,
               agent_cnt=np.zeros((N, 2), dtype=np.int64) if want_agents else None)
    hit = (x == px) & (y == py)
    if hit.sum() > 0:                                   # SDE:321-327
        a = int(np.argwhere(hit)[0][0])
        out["done"], out["eaten"] = True, a
        out["glob"][0] = eat
        out["glob"][4] = eat
        if want_agents:
            out["agent"][a, 0] = eat
            out["agent"][a, 4] = eat
        return out
    nn = N * (N - 1)
    xn = _cosine_half(eff)
    rep_rows = np.zeros(N, dtype=np.int64)
    att_rows = np.zeros(N, dtype=np.int64)
    una_rows = np.zeros(N, dtype=np.float64)
    tot_r1 = tot_r2 = tot_att = 0
    tot_una = 0.0
    xi64, yi64 = x.astype(np.int64), y.astype(np.int64)
    for lo in range(0, N, block):                       # row blocks keep N=16384 in memory
        hi = min(N, lo + block)
        d1 = np.maximum(np.abs(xi64[lo:hi, None] - xi64[None, :]), np.abs(yi64[lo:hi, None] - yi64[None, :]))
        d2 = S - d1                                     # SDE:333 (NOT a torus distance)
        r1 = d1 < rep                                   # SDE:336-337
        r2 = d2 < rep
        a1 = (d1 < att) == (d1 >= rep)                  # SDE:341-342 boolean XNOR
        a2 = (d2 < att) == (d2 >= rep)
        un = 1.0 - xn[lo:hi] @ xn.T
        np.clip(un, 0.0, 2.0, out=un)
        un /= 2.0
        if vf is not None:
            un[d1 > vf] = 0                             # SDE:349-355
        un[np.arange(hi - lo), np.arange(lo, hi)] = 0   # zero diagonal (sklearn + SDE:358)
        rep_rows[lo:hi] = r1.sum(axis=1) + r2.sum(axis=1)
        att_rows[lo:hi] = a1.sum(axis=1) + a2.sum(axis=1)
        una_rows[lo:hi] = un.sum(axis=1)
        tot_r1 += int(r1.sum())
        tot_r2 += int(r2.sum())
        tot_att += int(a1.sum()) + int(a2.sum())
        tot_una += float(un.sum())
    rew_rep = -(tot_r1 - N + tot_r2)                    # SDE:338
    rew_attr = tot_att                                  # SDE:343
    rew_unalign = -tot_una                              # SDE:359
    g = out["glob"]
    g[2] = w_rep * rew_rep / nn                         # SDE:363-366
    g[1] = w_att * rew_attr / nn
    g[3] = w_ali * rew_unalign / nn
    g[4] = g[2] + g[1] + g[3]
    out["cnt_rep"], out["cnt_att"] = int(rew_rep), int(rew_attr)
    if want_agents:                                     # SDE:371-379
        rep_i = -(rep_rows - 1)
        ag = out["agent"]
        ag[:, 2] = w_rep * rep_i / nn
        ag[:, 1] = w_att * att_rows / nn
        ag[:, 3] = w_ali * (-una_rows) / nn
        ag[:, 4] = ag[:, 2] + ag[:, 1] + ag[:, 3]
        out["agent_cnt"][:, 0] = rep_i
        out["agent_cnt"][:, 1] = att_rows
    return out


def reset_env(N, S, place, ptape):
    """SDE:197-221 + SDE:90-109.  Returns dict(x, y, px, py, target, err,
    place_used, pred_used)."""
    place = np.asarray(place, dtype=np.int64)
    ptape = np.asarray(ptape, dtype=np.int64).reshape(-1, 2)
    err = 0
    x = place[0:N].copy()
    y = place[N:2 * N].copy()
    cur = 2 * N
    while True:
        c = occupancy_counts(x, y, S)
        if np.all(c == 1):
            break
        slot, L = redraw_slots(c)
        if cur + 2 * L > len(place):
            err |= ERR_PLACE_OVERFLOW
            break
        coll = c > 1
        x[coll] = place[cur + slot[coll]]               # (2, L): L x's then L y's
        y[coll] = place[cur + L + slot[coll]]
        cur += 2 * L
    a = 0
    px = py = 0
    while True:
        if a >= len(ptape):
            err |= ERR_PRED_OVERFLOW
            break
        px, py = int(ptape[a, 0]), int(ptape[a, 1])
        a += 1
        if not np.any((x == px) & (y == py)):
            break
    target = closest_target(px, py, x, y, S) if not (err & ERR_PRED_OVERFLOW) else 0
    return dict(x=x, y=y, px=px, py=py, target=target, err=err, place_used=cur, pred_used=a)


class OracleEnv:
    """One environment; mirrors the reference object state."""

    def __init__(self, N=4, S=20, att=5, rep=2, eat=-10.0):
        self.N, self.S, self.att, self.rep, self.eat = N, S, att, rep, eat
        self.done = None
        self.err = 0

    def reset(self, place, ptape):
        r = reset_env(self.N, self.S, place, ptape)
        self.x, self.y = r["x"], r["y"]
        self.px, self.py, self.target = r["px"], r["py"], r["target"]
        self.orient = 0                                  # SDE:107
        self.done = False
        self.err = r["err"]
        return r

    def step(self, actions, slab, retarget, weights=(1.0, 1.0, 1.0), vf=None, want_agents=True):
        if self.done:
            raise RuntimeError("Episode has finished. Call env.reset() to start a new episode.")  # SDE:233-234
        N, S = self.N, self.S
        eff, nx, ny, consumed, err = resolve_collisions(self.x, self.y, actions, S, slab)
        pred_prev = (self.px, self.py)                   # SDE:258
        npx, npy, target, orient = follow_target(self.px, self.py, self.target, self.orient,
                                                 self.x, self.y, S, retarget)   # OLD agent positions, SDE:259
        self.x, self.y = nx, ny                          # SDE:260
        self.px, self.py, self.target, self.orient = npx, npy, target, orient
        rw = swarm_reward(nx, ny, eff, npx, npy, N, S, self.att, self.rep, self.eat, weights, vf, want_agents)
        self.done = rw["done"]
        self.err |= err
        rw.update(x=nx.copy(), y=ny.copy(), pred_prev=pred_prev, pred_next=(npx, npy), pred_orient=orient,
                  target=target, consumed=consumed, eff=eff, err=self.err)
        return rw


class OracleBatch:
    """E independent environments with the auto-reset semantics the CUDA path
    implements (SURVEY.md 8b "Auto-reset semantics"): on the step where ``done``
    becomes true the step's terminal reward/done/info are returned, the terminal
    positions go to ``term_x/term_y``, then the env is reset from reset-tape slot
    ``reset_count % R`` and the returned ``x/y/px/py/target`` are post-reset.
    With auto-reset off a finished env is frozen and stepping it sets
    ERR_STEP_WHEN_DONE."""

    def __init__(self, E, N, S, att=5, rep=2, eat=-10.0, auto_reset=False):
        self.E, self.N, self.S = E, N, S
        self.envs = [OracleEnv(N, S, att, rep, eat) for _ in range(E)]
        self.auto_reset = auto_reset
        self.reset_count = np.zeros(E, dtype=np.int64)
        self.place = self.ptape = None

    def set_reset_tapes(self, place, ptape):
        """place int[R,E,PL], ptape int[R,E,KP,2]"""
        self.place, self.ptape = np.asarray(place), np.asarray(ptape)

    def _reset_one(self, e):
        R = self.place.shape[0]
        slot = int(self.reset_count[e] % R)
        self.reset_count[e] += 1
        return self.envs[e].reset(self.place[slot, e], self.ptape[slot, e])

    def reset(self):
        self.reset_count[:] = 0
        for e in range(self.E):
            self._reset_one(e)
        return self.state()

    def state(self):
        E, N = self.E, self.N
        st = dict(x=np.zeros((E, N), np.int16), y=np.zeros((E, N), np.int16), px=np.zeros(E, np.int16),
                  py=np.zeros(E, np.int16), target=np.zeros(E, np.int32), err=np.zeros(E, np.uint32),
                  frozen=np.zeros(E, np.uint8))
        for e, env in enumerate(self.envs):
            st["x"][e], st["y"][e] = env.x, env.y
            st["px"][e], st["py"][e], st["target"][e] = env.px, env.py, env.target
            st["err"][e] = env.err
            st["frozen"][e] = bool(env.done)
        return st

    def step(self, actions, slab, retarget, weights=(1.0, 1.0, 1.0), vf=None, want_agents=True):
        E, N = self.E, self.N
        o = dict(done=np.zeros(E, np.uint8), eaten=np.full(E, -1, np.int32), glob=np.zeros((E, 5)),
                 cnt=np.zeros((E, 2), np.int64), pred_prev=np.zeros((E, 2), np.int16),
                 pred_next=np.zeros((E, 2), np.int16), pred_orient=np.zeros(E, np.uint8),
                 step_target=np.zeros(E, np.int32), consumed=np.zeros(E, np.int32),
                 term_x=np.zeros((E, N), np.int16), term_y=np.zeros((E, N), np.int16),
                 eff=np.zeros((E, N), np.uint8))
        if want_agents:
            o["agent"] = np.zeros((E, N, 5))
            o["agent_cnt"] = np.zeros((E, N, 2), np.int64)
        for e, env in enumerate(self.envs):
            if env.done:
                env.err |= ERR_STEP_WHEN_DONE
                o["done"][e] = 1
                continue
            r = env.step(actions[e], slab[e], bool(retarget[e]), weights, vf, want_agents)
            o["done"][e] = r["done"]
            o["eaten"][e] = r["eaten"]
            o["glob"][e] = r["glob"]
            o["cnt"][e] = (r["cnt_rep"], r["cnt_att"])
            o["pred_prev"][e] = r["pred_prev"]
            o["pred_next"][e] = r["pred_next"]
            o["pred_orient"][e] = r["pred_orient"]
            o["step_target"][e] = r["target"]
            o["consumed"][e] = r["consumed"]
            o["eff"][e] = r["eff"]
            o["term_x"][e], o["term_y"][e] = r["x"], r["y"]
            if want_agents:
                o["agent"][e] = r["agent"]
                o["agent_cnt"][e] = r["agent_cnt"]
            if r["done"] and self.auto_reset:
                keep = env.err
                self._reset_one(e)
                env.err |= keep
        o.update(self.state())
        return o


def visual_field(x, y, px, py, S, vf):
    """Egocentric (2vf+1)^2 windows, uint8[N, 2vf+1, 2vf+1] indexed [dy+vf][dx+vf]: 0 empty / outside the grid, 1 self,
    2 other agent, 3 predator.  PARITY UNPINNED: the reference tree holds no code for this (only the README figure
    gym_swarm/images/env_illustration.png, "Visual Field 5x5"; its ``vf_size`` gate at SDE:349-355 fixes the metric:
    plain, non-periodic Chebyshev distance).  This function is the definition the CUDA kernel is tested against."""
    N, W = len(x), 2 * vf + 1
    out = np.zeros((N, W, W), dtype=np.uint8)
    for i in range(N):
        for j in range(N):
            dx, dy = int(x[j]) - int(x[i]), int(y[j]) - int(y[i])
            if abs(dx) <= vf and abs(dy) <= vf:
                out[i, dy + vf, dx + vf] = 2
        out[i, vf, vf] = 1                               # the centre is the agent itself
        dx, dy = int(px) - int(x[i]), int(py) - int(y[i])
        if abs(dx) <= vf and abs(dy) <= vf:
            out[i, dy + vf, dx + vf] = 3
    return out
