"""The integer identities the CUDA kernels rely on, restated in NumPy and checked against the LITERAL N x N formulas of
the reference's ``swarm_reward`` (SDE:330-379) -- host-side documentation of the math in
``gym-swarm_b200/csrc/swarm_common.cuh`` (``agent_terms``, ``align_PQ``, ``box_count``), ``csrc/swarm_abi.cu``
(``fill_thresholds``) and ``csrc/fused_step.cuh`` (``far_border_counts``, the sparse visual-field sums).  The kernels
themselves are compared with the oracle in the ``-m gpu`` tests; these run on any CPU.

Third-party arithmetic: ``cosine_distances`` is scikit-learn's (imported here, as the reference does).
"""
import numpy as np
import pytest

MOVES = np.array([(-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1), (0, 1), (-1, 1), (0, 0)])   # SDE:467-475


def _distinct_cells(rng, N, S):
    cells = rng.choice(S * S, size=N, replace=False)
    return cells // S, cells % S


def _cheb(x, y):
    return np.maximum(np.abs(x[:, None] - x[None, :]), np.abs(y[:, None] - y[None, :]))


def _literal_counts(D, S, rep, att):
    """SDE:330-343, 371-373: per-agent (rew_rep_i, rew_attr_i) with the diagonal included literally."""
    D2 = S - D
    R = (D < rep).astype(int) + (D2 < rep).astype(int)
    A = ((D < att) == (D >= rep)).astype(int) + ((D2 < att) == (D2 >= rep)).astype(int)
    return -(R.sum(1) - 1), A.sum(1)


def _clamp(v, lo, hi):
    return max(lo, min(hi, v))


def _thresholds(S, rep, att):
    """fill_thresholds (swarm_abi.cu): four thresholds on D and the literal diagonal terms."""
    t_rep, t_att = _clamp(rep, 0, S), _clamp(att, 0, S)
    t_repf, t_attf = _clamp(S - rep + 1, 0, S), _clamp(S - att + 1, 0, S)
    diag_rep = int(0 < rep) + int(S < rep)
    diag_att = int((0 < att) == (0 >= rep)) + int((S < att) == (S >= rep))
    return t_rep, t_att, t_repf, t_attf, diag_rep, diag_att


def _c(D, t):
    """off-diagonal threshold count c(t)_i = #{j != i : D_ij < t}"""
    return (D < t).sum(1) - (np.diag(D) < t).astype(int)


@pytest.mark.parametrize("S", [6, 9, 20, 33])
def test_threshold_count_identity(S):
    """agent_terms: rep_i = -(cR + (N-1) - cRf + diag_rep - 1), att_i = |cA - cR| + |cRf - cAf| + diag_att."""
    rng = np.random.default_rng(S)
    for _ in range(60):
        N = int(rng.integers(2, min(S * S - 1, 40)))
        rep, att = int(rng.integers(0, S + 4)), int(rng.integers(0, S + 4))
        if rng.random() < 0.1:
            rep = 200                                       # the reference test-suite's repulsion threshold
        x, y = _distinct_cells(rng, N, S)
        D = _cheb(x, y)
        want_rep, want_att = _literal_counts(D, S, rep, att)
        t_rep, t_att, t_repf, t_attf, diag_rep, diag_att = _thresholds(S, rep, att)
        cR, cA, cRf, cAf = _c(D, t_rep), _c(D, t_att), _c(D, t_repf), _c(D, t_attf)
        got_rep = -(cR + (N - 1) - cRf + diag_rep - 1)
        got_att = np.abs(cA - cR) + np.abs(cRf - cAf) + diag_att
        np.testing.assert_array_equal(got_rep, want_rep, err_msg=f"rep S={S} N={N} rep={rep} att={att}")
        np.testing.assert_array_equal(got_att, want_att, err_msg=f"att S={S} N={N} rep={rep} att={att}")


@pytest.mark.parametrize("S,t", [(16, 9), (16, 12), (16, 16), (33, 17), (33, 30), (128, 124)])
def test_far_border_lemma(S, t):
    """far_border_counts: for t > S/2 a pair with D >= t has its agents in opposite border bands of width b = S - t, so
    counting among the band agents only gives #{j : D_ij >= t} for everyone (0 outside the band)."""
    assert 2 * t > S
    rng = np.random.default_rng(S * 100 + t)
    b = S - t
    for _ in range(30):
        N = int(rng.integers(2, min(S * S - 1, 150)))
        x, y = _distinct_cells(rng, N, S)
        D = _cheb(x, y)
        want = (D >= t).sum(1)
        band = (x < b) | (x >= S - b) | (y < b) | (y >= S - b)
        got = np.zeros(N, dtype=int)
        idx = np.nonzero(band)[0]
        if idx.size:
            got[idx] = (_cheb(x[idx], y[idx]) >= t).sum(1)
        np.testing.assert_array_equal(got, want)


def _bitmap(x, y, S):
    occ = np.zeros((S, S), dtype=int)                      # occ[y, x]
    occ[y, x] = 1
    return occ


def _box(occ, S, x0, x1, y0, y1):
    if x0 > x1 or y0 > y1:
        return 0
    return int(occ[y0:y1 + 1, x0:x1 + 1].sum())


@pytest.mark.parametrize("S", [8, 21, 64])
def test_box_counts_equal_pairwise_counts(S):
    """box_count: #{j : D_ij < t} (self included) from the occupancy bitmap -- directly (clipped box of radius t - 1) and
    as the complement N - (|dx| >= t) - (|dy| >= t) + (both) from coordinate histograms and four corner boxes."""
    rng = np.random.default_rng(S)
    for _ in range(25):
        N = int(rng.integers(2, min(S * S - 1, 120)))
        x, y = _distinct_cells(rng, N, S)
        D = _cheb(x, y)
        occ = _bitmap(x, y, S)
        px = np.cumsum(np.bincount(x, minlength=S))        # inclusive prefix histograms
        py = np.cumsum(np.bincount(y, minlength=S))

        def le(pre, a):
            return 0 if a < 0 else (N if a >= S - 1 else int(pre[a]))
        for t in sorted({1, 2, 5, S // 2, S - 3, S - 1} & set(range(1, S))):
            want = (D < t).sum(1)
            for i in range(N):
                xi, yi = int(x[i]), int(y[i])
                r = t - 1
                direct = _box(occ, S, max(0, xi - r), min(S - 1, xi + r), max(0, yi - r), min(S - 1, yi + r))
                nxs = le(px, xi - t) + N - le(px, xi + t - 1)
                nys = le(py, yi - t) + N - le(py, yi + t - 1)
                both = (_box(occ, S, 0, xi - t, 0, yi - t) + _box(occ, S, 0, xi - t, yi + t, S - 1) +
                        _box(occ, S, xi + t, S - 1, 0, yi - t) + _box(occ, S, xi + t, S - 1, yi + t, S - 1))
                assert direct == want[i], (S, N, t, i)
                assert N - (nxs + nys - both) == want[i], (S, N, t, i)


def _align_PQ(a, Ax, Ay, Dx, Dy, self_gate):
    """align_PQ (swarm_common.cuh): cos sums as integers, cosSum_i = P/2 + (sqrt2/2) Q minus the own contribution."""
    if a >= 8:
        return 0, 0
    mx, my = MOVES[a]
    sa, sd = mx * Ax + my * Ay, mx * Dx + my * Dy
    return (sd - 2 * self_gate, sa) if a & 1 else (2 * sa - 2 * self_gate, sd)


@pytest.mark.parametrize("vf", [None, 0, 1, 3, 50])
def test_alignment_decomposition(vf):
    """una_i = vis/2 - P/4 - (sqrt2/4) Q equals sum_j cosine_distances(moves)/2 with the diagonal cleared and, with
    vf_size, the entries at D > vf cleared (SDE:346-359, 374); actions 0..8, the zero move included."""
    from sklearn.metrics.pairwise import cosine_distances
    rng = np.random.default_rng(7 if vf is None else vf)
    S = 12
    for _ in range(40):
        N = int(rng.integers(2, 40))
        x, y = _distinct_cells(rng, N, S)
        D = _cheb(x, y)
        acts = rng.integers(0, 9, size=N)
        un = cosine_distances(MOVES[acts].astype(float)) / 2
        if vf is not None:
            un[D > vf] = 0
        np.fill_diagonal(un, 0)
        want = un.sum(1)
        axis = (acts < 8) & (acts % 2 == 0)
        diag = (acts < 8) & (acts % 2 == 1)
        got = np.zeros(N)
        for i in range(N):
            gate = np.ones(N, bool) if vf is None else (D[i] <= vf)    # includes i itself (D = 0) unless vf < 0
            Ax, Ay = MOVES[acts[gate & axis]].sum(0) if (gate & axis).any() else (0, 0)
            Dx, Dy = MOVES[acts[gate & diag]].sum(0) if (gate & diag).any() else (0, 0)
            self_gate = int(gate[i])
            vis = int(gate.sum()) - self_gate
            P, Q = _align_PQ(int(acts[i]), Ax, Ay, Dx, Dy, self_gate)
            got[i] = 0.5 * vis - 0.25 * P - 0.35355339059327376220 * Q
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_sparse_visual_field_sums_equal_dense_gating():
    """Sparse visual-field mode (fused_step.cuh): walking the list of visible UNORDERED pairs and crediting both agents
    with the other's move gives the same gated sums as gating every ordered pair, self excluded."""
    rng = np.random.default_rng(11)
    S, vf = 24, 3
    for _ in range(30):
        N = int(rng.integers(2, 60))
        x, y = _distinct_cells(rng, N, S)
        D = _cheb(x, y)
        acts = rng.integers(0, 9, size=N)
        mv = MOVES[acts]
        dense_cnt = (D <= vf).sum(1) - 1
        dense_sum = ((D <= vf)[:, :, None] * mv[None, :, :]).sum(1) - mv
        cnt = np.zeros(N, int)
        acc = np.zeros((N, 2), int)
        for i in range(N):
            for j in range(i + 1, N):
                if D[i, j] <= vf:
                    cnt[i] += 1; cnt[j] += 1
                    acc[i] += mv[j]; acc[j] += mv[i]
        np.testing.assert_array_equal(cnt, dense_cnt)
        np.testing.assert_array_equal(acc, dense_sum)
