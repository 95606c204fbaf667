// tiny_packed.cuh -- the thread-per-env transition for EXACTLY 8 agents with the positions kept packed (sm_100a).
//
// Same step as tiny_step_kernel<8, false> (tiny_step.cuh; reference citations there), same results bit for bit.  What
// changes is the representation: an env's 8 x's are the four 32-bit words of ONE LDG.128 (two agents per word) and they
// are never unpacked.
//   * action table (SDE:467-475): the 8 action bytes index two byte tables with PRMT (one per axis; id 8 is folded onto
//     an entry that holds 0), a second PRMT sign-extends the looked-up bytes into the 16x2 layout;
//   * move + torus wrap (SDE:45-57, 60-80): t = x + d per halfword, then min_u(t, t + S) and min_u(r, r - S) as two
//     VIADDMNMX.U16x2 -- t = -1 wraps to S - 1, t = S to 0, everything else is unchanged;
//   * collision counts (SDE:269-292): the 28 cell comparisons add into ONE register of 4-bit counters;
//   * predator arg-min and termination test: the packed forms of predator_async.cuh;
//   * reward: the symmetric pair sweep of tiny_step.cuh, its pair words being the packed positions themselves;
//   * alignment sums (SDE:346-359): signed byte dot products (IDP.4A) of the looked-up move bytes under the odd-id mask.
// ~40 % fewer issued instructions per env than the scalar kernel and half its registers.
// Chosen by tiny_step_launch for N == 8 without per-agent outputs (cfg5); SWARM_FLAG_TINY_SCALAR keeps the scalar kernel.
#pragma once
#include "tiny_step.cuh"

namespace swarm {

#ifndef TP_MINB
#define TP_MINB 8      // 64 registers: measured at cfg5 0.0787 ms (4 CTAs, 85 regs), 0.0745 (6), 0.0708 (8)
#endif

// PRMT in its generic mode: a selector nibble with bit 3 set replicates the SIGN of the selected byte (sign extension of a
// byte into a halfword); __byte_perm documents only the low three bits, so the instruction is written out
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// per pair word: packed (dx, dy) of its two agents, sign-extended halves; dxb/dyb: the same as signed bytes per 4 agents
__device__ __forceinline__ void tiny8_deltas(const uint2 a, unsigned (&dxw)[4], unsigned (&dyw)[4], unsigned (&dxb)[2],
                                             unsigned (&dyb)[2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const unsigned aw = h ? a.y : a.x;                       // bytes 0..8
    const unsigned is8 = (aw >> 3) & 0x01010101u;
    const unsigned ax = aw - 6u * is8;                       // id 8 -> entry 2 (dx = 0); no borrow: 8 >= 6
    const unsigned ay = aw & 0x07070707u;                    // id 8 -> entry 0 (dy = 0)
    const unsigned tx = (ax | (ax >> 4)) & 0x00FF00FFu, ty = (ay | (ay >> 4)) & 0x00FF00FFu;
    const unsigned sx = __byte_perm(tx, 0u, 0x4420), sy = __byte_perm(ty, 0u, 0x4420);   // four selector nibbles
    dxb[h] = __byte_perm(0x0100FFFFu, 0xFF000101u, sx);      // dx: -1 -1  0  1 | 1  1  0 -1
    dyb[h] = __byte_perm(0xFFFFFF00u, 0x01010100u, sy);      // dy:  0 -1 -1 -1 | 0  1  1  1
    dxw[2 * h] = prmt(dxb[h], 0u, 0x9180u);                  // byte -> halfword, sign replicated
    dxw[2 * h + 1] = prmt(dxb[h], 0u, 0xB3A2u);
    dyw[2 * h] = prmt(dyb[h], 0u, 0x9180u);
    dyw[2 * h + 1] = prmt(dyb[h], 0u, 0xB3A2u);
  }
}
// one-step torus wrap of two packed coordinates in [-1, S]
__device__ __forceinline__ unsigned wrap2(unsigned t, unsigned S2, unsigned NS2) {
  const unsigned r = __viaddmin_u16x2(t, S2, t);
  return __viaddmin_u16x2(r, NS2, r);
}
// 4-bit excess counters: nibble k = (number of agents on agent k's cell) - 1
__device__ __forceinline__ unsigned tiny8_excess(const unsigned (&xw)[4], const unsigned (&yw)[4]) {
  unsigned key[8];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    key[2 * w] = __byte_perm(xw[w], yw[w], 0x5410);
    key[2 * w + 1] = __byte_perm(xw[w], yw[w], 0x7632);
  }
  unsigned c = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = i + 1; j < 8; ++j)
      if (key[i] == key[j]) c += (1u << (4 * i)) + (1u << (4 * j));
  return c;
}
__device__ __forceinline__ int nibble_sum(unsigned c) {
  const unsigned b = (c & 0x0F0F0F0Fu) + ((c >> 4) & 0x0F0F0F0Fu);
  return (int)((b * 0x01010101u) >> 24);
}
__device__ __forceinline__ unsigned neg2(unsigned v) { return __vadd2(~v, 0x00010001u); }

__global__ void __launch_bounds__(TINY_THREADS, TP_MINB) tiny8_packed_kernel(const __grid_constant__ StepParams p) {
  const int env = blockIdx.x * TINY_THREADS + threadIdx.x;
  if (env >= p.E) return;
  constexpr int N = 8;
  const int S = p.S;
  const size_t base = (size_t)env * N;

  if (p.frozen[env]) {                      // SDE:233-234 (see tiny_step_kernel)
    p.err[env] |= SWARM_ERR_STEP_WHEN_DONE;
    p.done[env] = 1;
    p.eaten[env] = -1;
    p.counts[2 * env] = 0; p.counts[2 * env + 1] = 0;
    p.consumed[env] = 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) p.reward[5 * env + q] = 0.f;
    const unsigned pp = pack2(p.px[env], p.py[env]);
    reinterpret_cast<unsigned*>(p.pred_prev)[env] = pp;
    reinterpret_cast<unsigned*>(p.pred_next)[env] = pp;
    p.step_orient[env] = p.orient[env];
    p.step_target[env] = p.target[env];
    if (p.eff_action) *reinterpret_cast<uint2*>(p.eff_action + base) = make_uint2(0x08080808u, 0x08080808u);
    return;
  }

  // ---- load: one DRAM round trip for everything the step reads ---------------------------------------
  const uint4 xv = reinterpret_cast<const uint4*>(p.x)[env];
  const uint4 yv = reinterpret_cast<const uint4*>(p.y)[env];
  uint2 eff = reinterpret_cast<const uint2*>(p.actions)[env];
  int px = p.px[env], py = p.py[env];
  int target = p.target[env];
  const int retarget = tape_retarget(p, env);
  uint4 ep0 = make_uint4(0u, 0u, 0u, 0u);
  if (p.track_stats) ep0 = reinterpret_cast<const uint4*>(p.ep_rec)[env];
  unsigned err = 0;
  {                                          // ids > 8: error bit, clamped to 8 (rare)
    const unsigned bx = (((eff.x & 0x7F7F7F7Fu) + 0x77777777u) | eff.x) & 0x80808080u;
    const unsigned by = (((eff.y & 0x7F7F7F7Fu) + 0x77777777u) | eff.y) & 0x80808080u;
    if (bx | by) {
      err |= SWARM_ERR_BAD_ACTION;
      unsigned w[2] = {eff.x, eff.y};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned o = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) o |= min((w[h] >> (8 * q)) & 0xFFu, 8u) << (8 * q);
        w[h] = o;
      }
      eff = make_uint2(w[0], w[1]);
    }
  }
  const unsigned xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
  const unsigned S2 = rep2(S), NS2 = rep2(-S);

  // ---- move (SDE:237-244, 60-80) + first collision count (SDE:246, 269-292) ---------------------------
  unsigned nxw[4], nyw[4], dxb[2], dyb[2];
  {
    unsigned dxw[4], dyw[4];
    tiny8_deltas(eff, dxw, dyw, dxb, dyb);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      nxw[w] = wrap2(__vadd2(xw[w], dxw[w]), S2, NS2);
      nyw[w] = wrap2(__vadd2(yw[w], dyw[w]), S2, NS2);
    }
  }
  unsigned cpk = tiny8_excess(nxw, nyw);
  const uint8_t* __restrict__ slab = p.slab ? p.slab + (size_t)env * (size_t)p.K : nullptr;
  unsigned long long slab8 = 0ull;
  const bool slab_vec = p.slab != nullptr && p.K >= 8 && (((uintptr_t)slab) & 7) == 0;
  if (cpk != 0u && slab_vec) {               // prefetched behind the predator computation
    const uint2 v = *reinterpret_cast<const uint2*>(slab);
    slab8 = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
  }

  // ---- predator (SDE:258-259, 133-165) on the OLD agent positions ------------------------------
  const int prev_px = px, prev_py = py;
  if (retarget) {                            // SDE:111-131
    const unsigned NP = rep2(-px), P1 = rep2(px + 1), NQ = rep2(-py), Q1 = rep2(py + 1);
    unsigned lo = 0xFFFFFFFFu;
    nn_keys8_packed<0, false>(xv, yv, NP, P1, NQ, Q1, lo);
    const unsigned l16 = min(lo & 0xFFFFu, lo >> 16);
    target = (int)(l16 & 7u);
    if (2 * (int)(l16 >> 5) >= S) {          // ties go to the highest index only at d_min >= S/2
      unsigned hi = 0xFFFFFFFFu;
      nn_keys8_packed<0, true>(xv, yv, NP, P1, NQ, Q1, hi);
      const unsigned h16 = min(hi & 0xFFFFu, hi >> 16);
      target = nn_pick(((l16 >> 5) << 16) | (l16 & 7u), ((h16 >> 5) << 16) | (0xFFFFu - (7u - (h16 & 7u))), S);
    }
  }
  target = min(max(target, 0), N - 1);
  int orient;
  {
    const int w = target >> 1;
    const unsigned xs = w == 0 ? xv.x : (w == 1 ? xv.y : (w == 2 ? xv.z : xv.w));
    const unsigned ys = w == 0 ? yv.x : (w == 1 ? yv.y : (w == 2 ? yv.z : yv.w));
    const int sh = 16 * (target & 1);
    pursue(px, py, (int)(int16_t)((xs >> sh) & 0xFFFFu), (int)(int16_t)((ys >> sh) & 0xFFFFu), S, px, py, orient);
  }
  const int step_target = target;

  // ---- collision re-draw rounds (SDE:246-255) ---------------------------------------------------
  int cur = 0;
  while (cpk != 0u) {
    const int L = nibble_sum(cpk);
    if (cur + L > p.K) { err |= SWARM_ERR_SLAB_OVERFLOW; break; }
    int off = 0;
    unsigned w[2] = {eff.x, eff.y};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = (int)((cpk >> (4 * k)) & 15u);
      if (e > 0) {                                          // SDE:250-252: last of its c-1 draws wins
        const int d = cur + off + e - 1;
        const unsigned a = (slab_vec && d < 8) ? (unsigned)((slab8 >> (8 * d)) & 7ull) : (unsigned)tape_slab(p, slab, env, d);
        w[k >> 2] = (w[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (a << (8 * (k & 3)));
      }
      off += e;
    }
    cur += L;
    eff = make_uint2(w[0], w[1]);
    unsigned dxw[4], dyw[4];
    tiny8_deltas(eff, dxw, dyw, dxb, dyb);                  // agents that kept their action land where they landed before
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      nxw[q] = wrap2(__vadd2(xw[q], dxw[q]), S2, NS2);
      nyw[q] = wrap2(__vadd2(yw[q], dyw[q]), S2, NS2);
    }
    cpk = tiny8_excess(nxw, nyw);                           // change_position_id again
  }

  // ---- termination (SDE:307-309, 321-327) on the NEW positions ---------------------------------
  int eaten = -1;
  {
    const uint4 nxv = make_uint4(nxw[0], nxw[1], nxw[2], nxw[3]), nyv = make_uint4(nyw[0], nyw[1], nyw[2], nyw[3]);
    const unsigned kx = rep2(px), ky = rep2(py);
    unsigned acc = 0xFFFFFFFFu;
    hit8_packed(nxv, nyv, kx, ky, acc);
    if (zero_half(acc) != 0u) {
      unsigned hit = 0xFFFFFFFFu;
      hit8(nxv, nyv, kx, ky, 0, hit);
      eaten = (int)hit;
    }
  }
  const bool done = eaten >= 0;

  // ---- reward (SDE:330-379) ---------------------------------------------------------------------
  double r_surv = 0.0, r_att = 0.0, r_rep = 0.0, r_ali = 0.0, r_sum = 0.0;
  int g_rep = 0, g_att = 0;
  if (!done) {
    unsigned mxw[4], myw[4];                                // -x, -y
#pragma unroll
    for (int w = 0; w < 4; ++w) { mxw[w] = neg2(nxw[w]); myw[w] = neg2(nyw[w]); }
    const unsigned TR = rep2(p.t_rep), TA = rep2(p.t_att), TRf = rep2(p.t_repf), TAf = rep2(p.t_attf);
    const unsigned one = (unsigned)p.one, c256 = (unsigned)p.one << 8;
    unsigned row[8][2], col[4][2];                          // byte-packed counters as in tiny_step_kernel
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k][0] = row[k][1] = 0u;
#pragma unroll
    for (int w = 0; w < 4; ++w) col[w][0] = col[w][1] = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned sel = (i & 1) ? 0x3232u : 0x1010u;     // agent i replicated into both halves
      const unsigned XI = __byte_perm(nxw[i >> 1], 0u, sel), YI = __byte_perm(nyw[i >> 1], 0u, sel);
      const unsigned NXI = __byte_perm(mxw[i >> 1], 0u, sel), NYI = __byte_perm(myw[i >> 1], 0u, sel);
#pragma unroll
      for (int w = (i + 1) / 2; w < 4; ++w) {
        const unsigned m = neg_cheb2(XI, YI, NXI, NYI, make_uint4(nxw[w], nyw[w], mxw[w], myw[w]));
        const unsigned n0 = lt2(TA, m) * c256 + lt2(TR, m);
        const unsigned n1 = lt2(TAf, m) * c256 + lt2(TRf, m);
        row[i][0] = n0 * one + row[i][0];
        row[i][1] = n1 * one + row[i][1];
        if (2 * w == i) {                                   // own word of an even agent: no column credit for self
          col[w][0] = (n0 & 0xFFFF0000u) * one + col[w][0];
          col[w][1] = (n1 & 0xFFFF0000u) * one + col[w][1];
        } else {
          col[w][0] = n0 * one + col[w][0];
          col[w][1] = n1 * one + col[w][1];
        }
      }
    }
    // alignment sums (SDE:346-359), vf_size == None: A = summed moves of the axis movers (even ids < 8), D = diagonal
    int Ax = 0, Ay = 0, Dx = 0, Dy = 0, n_diag = 0, n8 = 0;
    {
      const unsigned w[2] = {eff.x, eff.y};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned odd = w[h] & 0x01010101u, mo = odd * 0xFFu;
        Dx = __dp4a((int)(dxb[h] & mo), 0x01010101, Dx);
        Dy = __dp4a((int)(dyb[h] & mo), 0x01010101, Dy);
        Ax = __dp4a((int)(dxb[h] & ~mo), 0x01010101, Ax);
        Ay = __dp4a((int)(dyb[h] & ~mo), 0x01010101, Ay);
        n_diag += __popc(odd);
        n8 += __popc(w[h] & 0x08080808u);
      }
    }
    const int n_axis = N - n_diag - n8;
    const int selfR = p.t_rep >= 1, selfA = p.t_att >= 1, selfRf = p.t_repf >= 1, selfAf = p.t_attf >= 1;
    unsigned sR = 0u, sRf = 0u, sAbsN = 0u, sAbsF = 0u;
    const unsigned selfN = (unsigned)(selfR | (selfA << 8)), selfF = (unsigned)(selfRf | (selfAf << 8));
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const unsigned f0 = row[2 * w][f] + (row[2 * w][f] >> 16), f1 = row[2 * w + 1][f] + (row[2 * w + 1][f] >> 16);
        const unsigned T = __byte_perm(f0, f1, 0x5410) + col[w][f] - (f ? selfF : selfN);
        const unsigned ad = __vabsdiffu4(T & 0x00FF00FFu, (T >> 8) & 0x00FF00FFu);
        if (f == 0) { sR = __dp4a(T, 0x00010001u, sR); sAbsN = __dp4a(ad, 0x00010001u, sAbsN); }
        else { sRf = __dp4a(T, 0x00010001u, sRf); sAbsF = __dp4a(ad, 0x00010001u, sAbsF); }
      }
    }
    g_rep = -((int)sR - (int)sRf + N * ((N - 1) + p.diag_rep - 1));
    g_att = (int)sAbsN + (int)sAbsF + N * p.diag_att;
    const int s_P = 2 * (Ax * Ax + Ay * Ay) - 2 * n_axis + (Dx * Dx + Dy * Dy) - 2 * n_diag;
    const int s_Q = 2 * (Ax * Dx + Ay * Dy);
    const double g_una = 0.5 * (double)(N * (N - 1)) - 0.25 * (double)s_P - 0.35355339059327376220 * (double)s_Q;
    r_rep = p.w_rep * (double)g_rep * p.inv_nn;             // SDE:363-366
    r_att = p.w_att * (double)g_att * p.inv_nn;
    r_ali = p.w_ali * (-g_una) * p.inv_nn;
    r_sum = r_rep + r_att + r_ali;
  } else {
    r_surv = (double)p.eat;
    r_sum = (double)p.eat;
  }
  if (p.eff_action) *reinterpret_cast<uint2*>(p.eff_action + base) = eff;

  // ---- per-env outputs ---------------------------------------------------------------------------
  p.done[env] = done ? 1 : 0;
  p.eaten[env] = eaten;
  {
    float* rw = p.reward + 5 * (size_t)env;
    rw[0] = (float)r_surv; rw[1] = (float)r_att; rw[2] = (float)r_rep; rw[3] = (float)r_ali; rw[4] = (float)r_sum;
  }
  reinterpret_cast<int2*>(p.counts)[env] = make_int2(g_rep, g_att);
  reinterpret_cast<unsigned*>(p.pred_prev)[env] = pack2(prev_px, prev_py);
  reinterpret_cast<unsigned*>(p.pred_next)[env] = pack2(px, py);
  p.step_orient[env] = (uint8_t)orient;
  p.step_target[env] = step_target;
  p.consumed[env] = cur;
  if (p.track_stats) episode_update(p, env, ep0, done, (float)r_surv, (float)r_att, (float)r_rep, (float)r_ali);

  // ---- commit / auto-reset (SDE:260; reset semantics SDE:197-221) -----------------------------
  if (done) {
    if (p.term_x) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        p.term_x[base + k] = (int16_t)((nxw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
        p.term_y[base + k] = (int16_t)((nyw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
      }
    }
    if (!p.auto_reset) p.frozen[env] = 1;               // auto-reset: tiny_reset_kernel follows, masked by done
  }
  reinterpret_cast<uint4*>(p.x)[env] = make_uint4(nxw[0], nxw[1], nxw[2], nxw[3]);
  reinterpret_cast<uint4*>(p.y)[env] = make_uint4(nyw[0], nyw[1], nyw[2], nyw[3]);
  p.px[env] = (int16_t)px;
  p.py[env] = (int16_t)py;
  p.target[env] = target;
  p.orient[env] = (uint8_t)orient;
  if (err) p.err[env] |= err;
}

}  // namespace swarm
