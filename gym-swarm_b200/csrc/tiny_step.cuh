// tiny_step.cuh -- the fused transition for tiny environments (N <= 8): ONE THREAD PER ENVIRONMENT (sm_100a).
//
// With 8 agents an env's whole state is 16 B of x, 16 B of y and 8 B of actions: one LDG.128 + LDG.128 +
// LDG.64 per thread, perfectly coalesced across the warp's 32 consecutive envs.  Everything the lane-group
// kernel does with shuffles (collision counts, re-draw prefix sum, predator arg-min, reward reductions) is a
// fully unrolled loop over registers here, and every per-env scalar (predator pursuit, outputs, statistics)
// is executed once per thread instead of once per lane group: ~10x fewer issued instructions per env than
// fused_step_kernel<8,1>.  Same semantics, same reference citations (SDE = swarm_discrete_env.py):
//   move + wrap SDE:237-244, 60-80; collision re-draw SDE:246-255, 269-292; predator SDE:258-259, 111-165;
//   termination SDE:307-309, 321-327; reward SDE:330-379; reset SDE:197-221, 90-109.
// The pairwise reward uses the symmetric evaluation: agent i (replicated into both 16-bit halves) meets only
// the pair words holding agents j > i; each DPX indicator is credited to i's row counter and to the word's
// column counter (byte-packed two-threshold counters as in the ring sweeps).
// vf_size != None is not handled here (the host dispatches those steps to fused_step_kernel<8,1>).
// Auto-reset is a second, masked launch (tiny_reset_kernel, thread per env, exits at once unless done): the step
// kernel is one long straight-line instruction stream per warp, and keeping the rarely taken reset code out of
// it is what keeps it inside the instruction cache (measured: 30 % of the stall samples were instruction
// fetches with the reset inlined).
#pragma once
#include "launchers.h"   // TINY_N, TINY_N_BIG
#include "swarm_common.cuh"

namespace swarm {

constexpr int TINY_THREADS = 128;
#ifndef TINY_MINB
#define TINY_MINB 1      // minimum resident CTAs per SM the 8-agent step kernel is compiled for (register cap)
#endif

// c[k] = number of agents on agent k's cell (SDE:269-292); keys of padding agents are unique. Returns sum(c-1).
template <int TN>
__device__ __forceinline__ int tiny_counts(const unsigned (&key)[TN], int (&c)[TN]) {
#pragma unroll
  for (int k = 0; k < TN; ++k) c[k] = 1;
#pragma unroll
  for (int i = 0; i < TN; ++i)
#pragma unroll
    for (int j = i + 1; j < TN; ++j) {
      const int eq = key[i] == key[j];
      c[i] += eq;
      c[j] += eq;
    }
  int L = 0;
#pragma unroll
  for (int k = 0; k < TN; ++k) L += c[k] - 1;
  return L;
}

template <int TN>
__device__ __forceinline__ int tiny_closest(int N, int S, int qx, int qy, const int (&ax)[TN], const int (&ay)[TN]) {
  unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < TN; ++k) {
    if (k < N) {
      const int d = max(iabs(qx - ax[k]), iabs(qy - ay[k]));
      klo = min(klo, nn_key_lo(d, k));
      khi = min(khi, nn_key_hi(d, k));
    }
  }
  return nn_pick(klo, khi, S);
}

// reference reset (SDE:197-221 + SDE:90-109) of one env by one thread
template <int TN>
__device__ __forceinline__ void tiny_reset(const StepParams& p, int env, unsigned rc, int (&x)[TN], int (&y)[TN],
                                           int& px, int& py, int& target, unsigned& err) {
  const int N = p.N, S = p.S;
  const int16_t* __restrict__ tape = place_row(p, env, rc);
  if (tape && N == TN && (((uintptr_t)tape) & 15) == 0) {   // the first placement as 128-bit loads
    unsigned xs[TN / 2], ys[TN / 2];
#pragma unroll
    for (int v = 0; v < TN / 8; ++v) {
      const uint4 xv = *reinterpret_cast<const uint4*>(tape + 8 * v);
      const uint4 yv = *reinterpret_cast<const uint4*>(tape + TN + 8 * v);
      xs[4 * v] = xv.x; xs[4 * v + 1] = xv.y; xs[4 * v + 2] = xv.z; xs[4 * v + 3] = xv.w;
      ys[4 * v] = yv.x; ys[4 * v + 1] = yv.y; ys[4 * v + 2] = yv.z; ys[4 * v + 3] = yv.w;
    }
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      x[k] = (int)(int16_t)((xs[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
      y[k] = (int)(int16_t)((ys[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
    }
  } else {
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      x[k] = k < N ? tape_place(p, tape, env, rc, k) : 0;       // SDE:202-203: all x's then all y's
      y[k] = k < N ? tape_place(p, tape, env, rc, N + k) : 0;
    }
  }
  int cur = 2 * N;
  for (;;) {
    unsigned key[TN];
    int c[TN];
#pragma unroll
    for (int k = 0; k < TN; ++k) key[k] = k < N ? pack2(x[k], y[k]) : (0xFFFF0000u | (unsigned)k);
    const int L = tiny_counts<TN>(key, c);
    if (L == 0) break;
    if (cur + 2 * L > p.PL) { err |= SWARM_ERR_PLACE_OVERFLOW; break; }
    int off = 0;
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      const int e = c[k] - 1;
      if (e > 0) {                                    // SDE:208-209: (2,L) block, last duplicate wins
        const int s = off + e - 1;
        x[k] = tape_place(p, tape, env, rc, cur + s);
        y[k] = tape_place(p, tape, env, rc, cur + L + s);
      }
      off += e;
    }
    cur += 2 * L;
  }
  const int16_t* __restrict__ pt = ptape_row(p, env, rc);
  int a = 0;
  px = 0; py = 0;
  for (;;) {                                          // SDE:92-105
    if (a >= p.KP) { err |= SWARM_ERR_PRED_OVERFLOW; break; }
    tape_pred(p, pt, env, rc, a, px, py);
    ++a;
    bool hit = false;
#pragma unroll
    for (int k = 0; k < TN; ++k) hit |= (k < N) && x[k] == px && y[k] == py;
    if (!hit) break;
  }
  target = tiny_closest<TN>(N, S, px, py, x, y);          // SDE:109
}

// AGENT_OUT: per-agent rewards / counts are written (otherwise only the env totals are formed, from the packed
// counters, without ever unpacking a per-agent value).
template <int TN, bool AGENT_OUT>
__global__ void __launch_bounds__(TINY_THREADS, TN == 8 ? TINY_MINB : 1) tiny_step_kernel(const __grid_constant__ StepParams p) {
  const int env = blockIdx.x * TINY_THREADS + threadIdx.x;
  if (env >= p.E) return;
  const int N = p.N, S = p.S;
  const size_t base = (size_t)env * (size_t)N;

  if (p.frozen[env]) {                      // SDE:233-234: finished episode, auto-reset off: nothing moves, every output
    p.err[env] |= SWARM_ERR_STEP_WHEN_DONE; // of the step is defined (zero reward, the predator where it stands)
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
    for (int k = 0; k < N; ++k) {
      if (AGENT_OUT && p.agent_reward) reinterpret_cast<float4*>(p.agent_reward)[base + k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (AGENT_OUT && p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[base + k] = make_int2(0, 0);
      if (AGENT_OUT && p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[base + k] = make_uint2(0u, 0u);
      if (p.eff_action) p.eff_action[base + k] = 8;
    }
    return;
  }

  // ---- load ---------------------------------------------------------------------------------------
  int x[TN], y[TN], eff[TN];
  if (N == TN) {
    unsigned xs[TN / 2], ys[TN / 2], as[TN / 4];
#pragma unroll
    for (int v = 0; v < TN / 8; ++v) {
      const uint4 xv = *reinterpret_cast<const uint4*>(p.x + base + 8 * v);
      const uint4 yv = *reinterpret_cast<const uint4*>(p.y + base + 8 * v);
      const uint2 av = *reinterpret_cast<const uint2*>(p.actions + base + 8 * v);
      xs[4 * v] = xv.x; xs[4 * v + 1] = xv.y; xs[4 * v + 2] = xv.z; xs[4 * v + 3] = xv.w;
      ys[4 * v] = yv.x; ys[4 * v + 1] = yv.y; ys[4 * v + 2] = yv.z; ys[4 * v + 3] = yv.w;
      as[2 * v] = av.x; as[2 * v + 1] = av.y;
    }
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      x[k] = (int)(int16_t)((xs[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
      y[k] = (int)(int16_t)((ys[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
      eff[k] = (int)((as[k >> 2] >> (8 * (k & 3))) & 0xFFu);
    }
  } else {
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      x[k] = k < N ? (int)p.x[base + k] : 0;
      y[k] = k < N ? (int)p.y[base + k] : 0;
      eff[k] = k < N ? (int)p.actions[base + k] : 8;
    }
  }
  int px = p.px[env], py = p.py[env];
  int target = p.target[env];
  const int retarget = tape_retarget(p, env);
  // everything else the step will read, issued now so that one DRAM round trip covers all of it
  uint4 ep0 = make_uint4(0u, 0u, 0u, 0u);
  if (p.track_stats) ep0 = reinterpret_cast<const uint4*>(p.ep_rec)[env];
  unsigned err = 0;
  {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < TN; ++k) { bad |= eff[k] > 8; eff[k] = min(eff[k], 8); }
    if (bad) err |= SWARM_ERR_BAD_ACTION;
  }

  // ---- move (SDE:237-244, 60-80) + first collision count (SDE:246, 269-292) ---------------------------
  int nx[TN], ny[TN], c[TN];
  unsigned key[TN];
#pragma unroll
  for (int k = 0; k < TN; ++k) {
    nx[k] = wrap1(x[k] + move_dx(eff[k]), S);
    ny[k] = wrap1(y[k] + move_dy(eff[k]), S);
    key[k] = k < N ? pack2(nx[k], ny[k]) : (0xFFFF0000u | (unsigned)k);
  }
  int L = tiny_counts<TN>(key, c);
  // the first 8 re-draw actions of this env are fetched now (only by envs that collided) and consumed after the
  // predator has been computed, so the load's latency is hidden
  const uint8_t* __restrict__ slab = p.slab ? p.slab + (size_t)env * (size_t)p.K : nullptr;
  unsigned long long slab8 = 0ull;
  const bool slab_vec = p.slab != nullptr && p.K >= 8 && (((uintptr_t)slab) & 7) == 0;
  if (L > 0 && slab_vec) {
    const uint2 v = *reinterpret_cast<const uint2*>(slab);
    slab8 = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
  }

  // ---- predator (SDE:258-259, 133-165) on the OLD agent positions ------------------------------
  const int prev_px = px, prev_py = py;
  if (retarget) target = tiny_closest<TN>(N, S, px, py, x, y);
  target = min(max(target, 0), N - 1);
  int orient;
  {
    int tx = 0, ty = 0;
#pragma unroll
    for (int k = 0; k < TN; ++k)
      if (k == target) { tx = x[k]; ty = y[k]; }
    pursue(px, py, tx, ty, S, px, py, orient);
  }
  const int step_target = target;

  // ---- collision re-draw rounds (SDE:246-255) ---------------------------------------------------
  int cur = 0;
  while (L > 0) {
    if (cur + L > p.K) { err |= SWARM_ERR_SLAB_OVERFLOW; break; }
    int off = 0;
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      const int e = c[k] - 1;
      if (e > 0) {                                        // SDE:250-252: last of its c-1 draws wins
        const int d = cur + off + e - 1;
        const int a = (slab_vec && d < 8) ? (int)((slab8 >> (8 * d)) & 7ull) : tape_slab(p, slab, env, d);
        eff[k] = a;
        nx[k] = wrap1(x[k] + move_dx(a), S);
        ny[k] = wrap1(y[k] + move_dy(a), S);
        key[k] = pack2(nx[k], ny[k]);
      }
      off += e;
    }
    cur += L;
    L = tiny_counts<TN>(key, c);                              // change_position_id again
  }

  // ---- termination (SDE:307-309, 321-327) on the NEW positions ---------------------------------
  int eaten = -1;
#pragma unroll
  for (int k = TN - 1; k >= 0; --k)
    if (k < N && nx[k] == px && ny[k] == py) eaten = k;
  const bool done = eaten >= 0;

  // ---- reward (SDE:330-379) ---------------------------------------------------------------------
  double r_surv = 0.0, r_att = 0.0, r_rep = 0.0, r_ali = 0.0, r_sum = 0.0;
  int g_rep = 0, g_att = 0;
  if (!done) {
    int sx[TN], sy[TN];
#pragma unroll
    for (int k = 0; k < TN; ++k) { sx[k] = k < N ? nx[k] : -S; sy[k] = k < N ? ny[k] : 0; }
    uint4 pw[TN / 2];
#pragma unroll
    for (int w = 0; w < TN / 2; ++w)
      pw[w] = make_uint4(pack2(sx[2 * w], sx[2 * w + 1]), pack2(sy[2 * w], sy[2 * w + 1]),
                         pack2(-sx[2 * w], -sx[2 * w + 1]), pack2(-sy[2 * w], -sy[2 * w + 1]));
    const unsigned TR = rep2(p.t_rep), TA = rep2(p.t_att), TRf = rep2(p.t_repf), TAf = rep2(p.t_attf);
    const unsigned one = (unsigned)p.one, c256 = (unsigned)p.one << 8;
    // byte-packed counters: [.][0] = {rep, att}, [.][1] = {rep_far, att_far}; a byte meets at most 8 pairs
    unsigned row[TN][2], col[TN / 2][2];
#pragma unroll
    for (int k = 0; k < TN; ++k) row[k][0] = row[k][1] = 0u;
#pragma unroll
    for (int w = 0; w < TN / 2; ++w) col[w][0] = col[w][1] = 0u;
#pragma unroll
    for (int i = 0; i < TN; ++i) {
      const unsigned XI = rep2(sx[i]), YI = rep2(sy[i]), NXI = rep2(-sx[i]), NYI = rep2(-sy[i]);
      // even i: its own word holds (i, i+1): low half = the self pair (row only), high half = pair (i, i+1);
      // odd i: its own word holds only pairs already met; the self pair is added analytically below
#pragma unroll
      for (int w = (i + 1) / 2; w < TN / 2; ++w) {
        const unsigned m = neg_cheb2(XI, YI, NXI, NYI, pw[w]);
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
    // alignment sums (SDE:346-359), vf_size == None
    int Ax = 0, Ay = 0, Dx = 0, Dy = 0, n_axis = 0, n_diag = 0;
#pragma unroll
    for (int k = 0; k < TN; ++k) {
      const int a = eff[k];
      if (k < N && a < 8) {
        const int mx = move_dx(a), my = move_dy(a);
        if (a & 1) { Dx += mx; Dy += my; ++n_diag; } else { Ax += mx; Ay += my; ++n_axis; }
      }
    }
    const int selfR = p.t_rep >= 1, selfA = p.t_att >= 1, selfRf = p.t_repf >= 1, selfAf = p.t_attf >= 1;
    int s_P, s_Q;
    if constexpr (AGENT_OUT) {
      s_P = 0; s_Q = 0;
      float4* __restrict__ ar = p.agent_reward ? reinterpret_cast<float4*>(p.agent_reward) + base : nullptr;
      int2* __restrict__ ac = p.agent_counts ? reinterpret_cast<int2*>(p.agent_counts) + base : nullptr;
      uint2* __restrict__ at = p.agent_terms ? reinterpret_cast<uint2*>(p.agent_terms) + base : nullptr;
#pragma unroll
      for (int k = 0; k < TN; ++k) {
        if (k < N) {
          const int sh = 16 * (k & 1);
          // counts over j != k: row (pairs met as the row agent; includes the self pair for even k) + column
          int cR = (int)__dp4a(row[k][0], 0x00010001u, 0u) + (int)((col[k >> 1][0] >> sh) & 0xFFu);
          int cA = (int)__dp4a(row[k][0], 0x01000100u, 0u) + (int)((col[k >> 1][0] >> (sh + 8)) & 0xFFu);
          int cRf = (int)__dp4a(row[k][1], 0x00010001u, 0u) + (int)((col[k >> 1][1] >> sh) & 0xFFu);
          int cAf = (int)__dp4a(row[k][1], 0x01000100u, 0u) + (int)((col[k >> 1][1] >> (sh + 8)) & 0xFFu);
          if ((k & 1) == 0) { cR -= selfR; cA -= selfA; cRf -= selfRf; cAf -= selfAf; }
          int P, Q;
          align_PQ(eff[k], Ax, Ay, Dx, Dy, 1, P, Q);
          const AgentTerms t = agent_terms(p, cR, cA, cRf, cAf, N - 1, P, Q);
          g_rep += t.rep_i; g_att += t.att_i; s_P += P; s_Q += Q;
          if (ar) {                                           // SDE:371-379
            const double a_rep = p.w_rep * (double)t.rep_i * p.inv_nn;
            const double a_att = p.w_att * (double)t.att_i * p.inv_nn;
            const double a_ali = p.w_ali * (-t.una) * p.inv_nn;
            ar[k] = make_float4((float)a_att, (float)a_rep, (float)a_ali, (float)(a_rep + a_att + a_ali));
          }
          if (ac) ac[k] = make_int2(t.rep_i, t.att_i);
          if (at) at[k] = pack_terms(t.rep_i, t.att_i, N - 1, P, Q);
        }
      }
    } else {
      // Env totals straight from the byte-packed counters.  Per pair word w the four bytes of T are
      // (c1, c2) of agent 2w and (c1, c2) of agent 2w+1 with (c1, c2) = (cR, cA) or (cRf, cAf):
      //   sum cR, sum cRf: one dp4a per word; sum |c2 - c1|: VABSDIFF4.U8 of the even / odd bytes + dp4a.
      unsigned sR = 0u, sRf = 0u, sAbsN = 0u, sAbsF = 0u;
      const unsigned selfN = (unsigned)(selfR | (selfA << 8)), selfF = (unsigned)(selfRf | (selfAf << 8));
#pragma unroll
      for (int w = 0; w < TN / 2; ++w) {
        const unsigned vm = (2 * w < N ? 0x0000FFFFu : 0u) | (2 * w + 1 < N ? 0xFFFF0000u : 0u);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const unsigned f0 = row[2 * w][f] + (row[2 * w][f] >> 16), f1 = row[2 * w + 1][f] + (row[2 * w + 1][f] >> 16);
          const unsigned T = (__byte_perm(f0, f1, 0x5410) + col[w][f] - (f ? selfF : selfN)) & vm;
          const unsigned ad = __vabsdiffu4(T & 0x00FF00FFu, (T >> 8) & 0x00FF00FFu);
          if (f == 0) { sR = __dp4a(T, 0x00010001u, sR); sAbsN = __dp4a(ad, 0x00010001u, sAbsN); }
          else { sRf = __dp4a(T, 0x00010001u, sRf); sAbsF = __dp4a(ad, 0x00010001u, sAbsF); }
        }
      }
      g_rep = -((int)sR - (int)sRf + N * ((N - 1) + p.diag_rep - 1));   // sum_i -(cR + (N-1) - cRf + diag - 1)
      g_att = (int)sAbsN + (int)sAbsF + N * p.diag_att;                // sum_i |cA-cR| + |cRf-cAf| + diag
      // sum_i P_i = 2 A.A - 2 n_axis + D.D - 2 n_diag,  sum_i Q_i = 2 A.D  (align_PQ summed over the agents)
      s_P = 2 * (Ax * Ax + Ay * Ay) - 2 * n_axis + (Dx * Dx + Dy * Dy) - 2 * n_diag;
      s_Q = 2 * (Ax * Dx + Ay * Dy);
    }
    const double g_una = 0.5 * (double)(N * (N - 1)) - 0.25 * (double)s_P - 0.35355339059327376220 * (double)s_Q;
    r_rep = p.w_rep * (double)g_rep * p.inv_nn;             // SDE:363-366
    r_att = p.w_att * (double)g_att * p.inv_nn;
    r_ali = p.w_ali * (-g_una) * p.inv_nn;
    r_sum = r_rep + r_att + r_ali;
  } else {
    r_surv = (double)p.eat;
    r_sum = (double)p.eat;
    if (AGENT_OUT && p.agent_reward) {
      float4* __restrict__ ar = reinterpret_cast<float4*>(p.agent_reward) + base;
#pragma unroll
      for (int k = 0; k < TN; ++k)
        if (k < N) ar[k] = make_float4(0.f, 0.f, 0.f, k == eaten ? p.eat : 0.f);
    }
    if (AGENT_OUT && p.agent_counts) {
      int2* __restrict__ ac = reinterpret_cast<int2*>(p.agent_counts) + base;
#pragma unroll
      for (int k = 0; k < TN; ++k)
        if (k < N) ac[k] = make_int2(0, 0);
    }
    if (AGENT_OUT && p.agent_terms) {
      uint2* __restrict__ at = reinterpret_cast<uint2*>(p.agent_terms) + base;
#pragma unroll
      for (int k = 0; k < TN; ++k)
        if (k < N) at[k] = make_uint2(0u, 0u);
    }
  }
  if (p.eff_action) {
    if (N == TN) {
#pragma unroll
      for (int v = 0; v < TN / 8; ++v)
        *reinterpret_cast<uint2*>(p.eff_action + base + 8 * v) =
            make_uint2((unsigned)(eff[8 * v] | (eff[8 * v + 1] << 8) | (eff[8 * v + 2] << 16) | (eff[8 * v + 3] << 24)),
                       (unsigned)(eff[8 * v + 4] | (eff[8 * v + 5] << 8) | (eff[8 * v + 6] << 16) | (eff[8 * v + 7] << 24)));
    } else {
#pragma unroll
      for (int k = 0; k < TN; ++k)
        if (k < N) p.eff_action[base + k] = (uint8_t)eff[k];
    }
  }

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
      for (int k = 0; k < TN; ++k)
        if (k < N) { p.term_x[base + k] = (int16_t)nx[k]; p.term_y[base + k] = (int16_t)ny[k]; }
    }
    if (!p.auto_reset) p.frozen[env] = 1;               // auto-reset: tiny_reset_kernel follows, masked by done
  }
  if (N == TN) {
#pragma unroll
    for (int v = 0; v < TN / 8; ++v) {
      *reinterpret_cast<uint4*>(p.x + base + 8 * v) = make_uint4(pack2(nx[8 * v], nx[8 * v + 1]), pack2(nx[8 * v + 2], nx[8 * v + 3]),
                                                                 pack2(nx[8 * v + 4], nx[8 * v + 5]), pack2(nx[8 * v + 6], nx[8 * v + 7]));
      *reinterpret_cast<uint4*>(p.y + base + 8 * v) = make_uint4(pack2(ny[8 * v], ny[8 * v + 1]), pack2(ny[8 * v + 2], ny[8 * v + 3]),
                                                                 pack2(ny[8 * v + 4], ny[8 * v + 5]), pack2(ny[8 * v + 6], ny[8 * v + 7]));
    }
  } else {
#pragma unroll
    for (int k = 0; k < TN; ++k)
      if (k < N) { p.x[base + k] = (int16_t)nx[k]; p.y[base + k] = (int16_t)ny[k]; }
  }
  p.px[env] = (int16_t)px;
  p.py[env] = (int16_t)py;
  p.target[env] = target;
  p.orient[env] = (uint8_t)orient;
  if (err) p.err[env] |= err;
}

// Masked reset of finished envs (auto-reset after tiny_step_kernel).  Finished envs are rare (a few per cent) and
// scattered, so each CTA first compacts the finished envs among its TINY_RESET_SPAN envs into a shared list and then
// resets them with consecutive threads: a couple of full warps run the (long, divergent) reset code, with all their
// dependent tape loads in flight together, instead of every warp running it for one lane.
constexpr int TINY_RESET_SPAN = TINY_THREADS * 8;     // envs scanned per CTA (8 mask bytes = one 64-bit load per thread)

template <int TN>
__global__ void __launch_bounds__(TINY_THREADS) tiny_reset_kernel(const __grid_constant__ StepParams p,
                                                                   const uint8_t* __restrict__ mask) {
  __shared__ int s_list[TINY_RESET_SPAN];
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  {
    const long long e0 = (long long)blockIdx.x * TINY_RESET_SPAN + (long long)threadIdx.x * 8;
    if (e0 + 8 <= p.E && ((((uintptr_t)mask) + (uintptr_t)e0) & 7) == 0) {
      const uint2 m = *reinterpret_cast<const uint2*>(mask + e0);
      if (m.x | m.y) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (((q < 4 ? m.x : m.y) >> (8 * (q & 3))) & 0xFFu) s_list[atomicAdd(&s_count, 1)] = (int)e0 + q;
      }
    } else {
      for (int q = 0; q < 8; ++q)
        if (e0 + q < p.E && mask[e0 + q]) s_list[atomicAdd(&s_count, 1)] = (int)e0 + q;
    }
  }
  __syncthreads();
  const int count = s_count;
  const int N = p.N;
  for (int q = threadIdx.x; q < count; q += TINY_THREADS) {
    const int env = s_list[q];
    const size_t base = (size_t)env * (size_t)N;
    const unsigned rc = p.reset_count[env];
    p.reset_count[env] = rc + 1u;
    int x[TN], y[TN], px, py, target;
    unsigned err = 0;
    tiny_reset<TN>(p, env, rc, x, y, px, py, target, err);
#pragma unroll
    for (int k = 0; k < TN; ++k)
      if (k < N) { p.x[base + k] = (int16_t)x[k]; p.y[base + k] = (int16_t)y[k]; }
    p.px[env] = (int16_t)px;
    p.py[env] = (int16_t)py;
    p.target[env] = target;
    p.orient[env] = 0;                                    // SDE:107
    if (err) p.err[env] |= err;
  }
}

}  // namespace swarm
