"""The packed-integer identities of the round-2 kernels, restated in NumPy and checked against the plain definitions
(the oracle's helpers, citing SDE = swarm_discrete_env.py) -- host-side documentation of the tricks in
``csrc/tiny_packed.cuh`` (PRMT action tables, two-step unsigned-min torus wrap, 4-bit collision counters),
``csrc/predator_async.cuh`` / ``csrc/swarm_common.cuh`` (packed nearest-neighbour keys, zero-halfword termination test),
``csrc/staged.cuh`` (incremental collision rounds with a hit list, overlapping-word bitmap and byte spreading of the
visual-field kernel).  The kernels themselves are compared with the oracle in the ``-m gpu`` tests; these run on any CPU.
"""
import numpy as np
import pytest

from oracle import swarm_oracle as so

U16 = np.uint16


def _s8(v):
    return v - 256 if v & 0x80 else v


def _s16(v):
    return v - 65536 if v & 0x8000 else v


def _prmt(a, b, sel):
    """PTX prmt.b32 (generic mode): nibble k of `sel` picks byte (nibble & 7) of {a (bytes 0-3), b (bytes 4-7)};
    bit 3 of the nibble replicates the sign of that byte instead."""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for k in range(4):
        n = (sel >> (4 * k)) & 0xF
        byte = src[n & 7]
        if n & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * k)
    return out


def test_action_table_by_byte_permute():
    """tiny8_deltas: two 8-entry byte tables per axis indexed by the action id, id 8 folded onto an entry that holds 0;
    a second permute sign-extends the bytes into the 16x2 layout (SDE:467-475)."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        acts = rng.integers(0, 9, size=4)
        aw = int(sum(int(a) << (8 * i) for i, a in enumerate(acts)))
        is8 = (aw >> 3) & 0x01010101
        ax = aw - 6 * is8
        ay = aw & 0x07070707
        tx, ty = (ax | (ax >> 4)) & 0x00FF00FF, (ay | (ay >> 4)) & 0x00FF00FF
        sx, sy = _prmt(tx, 0, 0x4420), _prmt(ty, 0, 0x4420)
        dxb = _prmt(0x0100FFFF, 0xFF000101, sx & 0xFFFF)
        dyb = _prmt(0xFFFFFF00, 0x01010100, sy & 0xFFFF)
        for k, a in enumerate(acts):
            want = so.ACTION_MOVES[a]
            assert _s8((dxb >> (8 * k)) & 0xFF) == want[0] and _s8((dyb >> (8 * k)) & 0xFF) == want[1]
        lo, hi = _prmt(dxb, 0, 0x9180), _prmt(dxb, 0, 0xB3A2)          # bytes -> sign-extended halfwords
        halves = [lo & 0xFFFF, lo >> 16, hi & 0xFFFF, hi >> 16]
        assert [_s16(h) for h in halves] == [int(so.ACTION_MOVES[a][0]) for a in acts]


@pytest.mark.parametrize("S", [2, 3, 16, 128, 1024, 32767])
def test_torus_wrap_by_two_unsigned_minima(S):
    """wrap2: for t = x + d in [-1, S] as a 16-bit pattern, min_u(t, t + S) then min_u(r, r - S) is the wrap of SDE:45-57."""
    t = np.arange(-1, S + 1, dtype=np.int64)
    u = t.astype(np.int16).view(U16).astype(np.uint32)
    r = np.minimum(u, (u + S) & 0xFFFF)
    r = np.minimum(r, (r - S) & 0xFFFF)
    np.testing.assert_array_equal(r.astype(np.int64), so.wrap(t, S))


def test_excess_counters_are_pair_equalities():
    """tiny8_excess: adding 1 into the 4-bit counters of BOTH agents of every equal pair leaves c_i - 1 in nibble i
    (SDE:269-292: change_position_id lists agent i once per other agent on its cell)."""
    rng = np.random.default_rng(1)
    for _ in range(300):
        cells = rng.integers(0, 6, size=8)
        packed = 0
        for i in range(8):
            for j in range(i + 1, 8):
                if cells[i] == cells[j]:
                    packed += (1 << (4 * i)) + (1 << (4 * j))
        got = [(packed >> (4 * k)) & 15 for k in range(8)]
        want = [(cells == cells[k]).sum() - 1 for k in range(8)]
        assert got == want
        b = (packed & 0x0F0F0F0F) + ((packed >> 4) & 0x0F0F0F0F)
        assert ((b * 0x01010101) & 0xFFFFFFFF) >> 24 == sum(want)      # nibble_sum


@pytest.mark.parametrize("S,N", [(3, 8), (16, 8), (9, 32), (128, 128), (2048, 32)])
def test_packed_nearest_neighbour_keys(S, N):
    """nn_keys8_packed + nn_pick: per lane the minimum of (d << 5) | local index picks the lowest index at the minimum
    Chebyshev distance, (d << 5) | (31 - local) the highest; the highest is only needed when 2 d_min >= S
    (SDE:111-131 under the stable-sort rule: oracle.closest_target)."""
    rng = np.random.default_rng(S + N)
    for _ in range(300):
        x, y = rng.integers(0, S, size=N), rng.integers(0, S, size=N)
        px, py = int(rng.integers(0, S)), int(rng.integers(0, S))
        d = np.maximum(np.abs(x - px), np.abs(y - py))
        lanes = N // 32 if N >= 32 else 1
        per = N // lanes
        klo, khi = [], []
        for l in range(lanes):                       # blocked ownership: local index = position inside the lane's block
            loc = np.arange(per)
            dl = d[l * per:(l + 1) * per]
            a, b = int(((dl << 5) | loc).min()), int(((dl << 5) | (31 - loc)).min())
            klo.append(((a >> 5) << 16) | (l * per + (a & 31)))
            khi.append(((b >> 5) << 16) | (0xFFFF - (l * per + 31 - (b & 31))))
        lo, hi = min(klo), min(khi)
        dmin, ilo, ihi = lo >> 16, lo & 0xFFFF, 0xFFFF - (hi & 0xFFFF)
        target = ilo if (ilo == ihi or 2 * dmin < S) else ihi
        lazy = ilo if 2 * dmin < S else target          # the kernels skip the second pass unless 2 d_min >= S
        assert target == lazy == so.closest_target(px, py, x, y, S)


def test_zero_halfword_test_and_its_one_false_positive():
    """zero_half(z) = (z - 0x00010001) & ~z & 0x80008000 is non-zero iff a halfword of z is zero; the only false flag is
    the HIGH half when the low half is zero (borrow) -- a lower index, so the minimum index is unaffected."""
    rng = np.random.default_rng(2)
    vals = np.concatenate([rng.integers(0, 1 << 32, size=20000, dtype=np.uint64),
                           (rng.integers(0, 4, size=4000, dtype=np.uint64) << 16) | rng.integers(0, 4, size=4000, dtype=np.uint64)])
    for z in vals.astype(np.uint64):
        z = int(z)
        zh = ((z - 0x00010001) & 0xFFFFFFFF) & (~z & 0xFFFFFFFF) & 0x80008000
        lo0, hi0 = (z & 0xFFFF) == 0, (z >> 16) == 0
        assert (zh != 0) == (lo0 or hi0)
        assert bool(zh & 0x8000) == lo0
        if zh & 0x80000000 and not hi0:
            assert lo0 and (z >> 16) == 1


def test_four_cells_to_four_code_bytes():
    """observe_vf_warp_kernel: ((bits & 15) * 0x00408102) & 0x02020202 puts code 2 into byte k for every set bit k."""
    for bits in range(16):
        e = ((bits & 0xF) * 0x00408102) & 0x02020202
        assert [(e >> (8 * k)) & 0xFF for k in range(4)] == [2 * ((bits >> k) & 1) for k in range(4)]


@pytest.mark.parametrize("S,vf", [(16, 1), (128, 2), (40, 5), (33, 3)])
def test_overlapping_word_bitmap_windows(S, vf):
    """observe_vf_warp_kernel: with the row-padded bitmap stored as 50 %-overlapping words (word v = bits [16 v - 16,
    16 v + 16)) every window row (<= 11 cells) is (word[(lo >> 4) + 1] >> (lo & 15)) & mask -- checked against the oracle's
    visual_field for the agent codes."""
    rng = np.random.default_rng(S * 7 + vf)
    N = min(64, S * S // 4)
    cells = rng.choice(S * S, size=N, replace=False)
    x, y = cells % S, cells // S
    nbits = (S + 2 * vf) * S
    words = np.zeros((nbits + 15) // 16 + 2, dtype=np.uint64)
    for xi, yi in zip(x, y):
        c = (yi + vf) * S + xi
        words[(c >> 4) + 1] |= np.uint64(1 << (c & 15))
        words[c >> 4] |= np.uint64(1 << ((c & 15) + 16))
    ref = so.visual_field(x.astype(np.int16), y.astype(np.int16), -100, -100, S, vf)      # predator far away
    W = 2 * vf + 1
    for i in range(N):
        x0, x1 = max(x[i] - vf, 0), min(x[i] + vf, S - 1)
        cmask, cshift = (1 << (x1 - x0 + 1)) - 1, x0 - (x[i] - vf)
        for r in range(W):
            lo = (y[i] + r) * S + x0                                     # padded rows: grid row y - vf is row y here
            bits = ((int(words[(lo >> 4) + 1]) >> (lo & 15)) & cmask) << cshift
            row = [(bits >> c) & 1 for c in range(W)]
            want = [1 if v in (1, 2) else 0 for v in ref[i, r]]
            assert row == want, (i, r)


def test_incremental_collision_rounds_count_like_the_full_rebuild():
    """collide_bitmap_kernel: keep the occupancy bitmap across the rounds of an env -- a re-drawn agent takes its old bit
    out, only moved agents are inserted, 'bit already set' is a hit and c_i - 1 = hits on agent i's cell -- and compare
    with the occupancy counts of a full rebuild (SDE:246-255, 269-292) over several random re-draw rounds."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        S, N = 6, 20
        old = rng.choice(S * S, size=N, replace=False)
        pos = np.array([rng.choice([c, (c + 1) % (S * S), (c + S) % (S * S)]) for c in old])
        bitmap = set()
        moved = np.ones(N, dtype=bool)
        for _ in range(6):
            hits = []
            for i in rng.permutation(np.nonzero(moved)[0]):             # insertion order is arbitrary on the GPU
                if pos[i] in bitmap:
                    hits.append(pos[i])
                bitmap.add(pos[i])
            extra = np.array([hits.count(p) for p in pos])
            full = np.array([(pos == p).sum() - 1 for p in pos])
            np.testing.assert_array_equal(extra, full)
            if not hits:
                break
            moved = extra > 0
            for i in np.nonzero(moved)[0]:
                bitmap.discard(pos[i])                                   # every occupant of a shared cell leaves it
            pos = pos.copy()
            pos[moved] = rng.integers(0, S * S, size=int(moved.sum()))
