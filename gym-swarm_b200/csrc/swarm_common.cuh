// swarm_common.cuh -- device helpers shared by the fused and the staged kernels (sm_100a).
//
// Reference semantics restated here (citations into /root/reference/gym_swarm/envs/swarm_discrete_env.py,
// "SDE"): action table SDE:467-475, wrap rule SDE:45-57 / 60-80, predator pursuit SDE:142-165,
// nearest-neighbour tie rule SDE:121-129 (stable argsort), reward formulas SDE:330-379.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swarm_abi.h"

namespace swarm {

// ------------------------------------------------------------------------------------------------
// Box counts: the sub-quadratic exact path for the reward's threshold counts.
// After the last collision round the shared bitmap holds exactly the N (distinct) final cells, so
//   #{j : D_ij < t}  (self included, D = Chebyshev)  =  popcount of the bitmap over the (2t-1)^2 box around i,
// and for a large threshold the complement is cheaper:
//   #{j : D_ij >= t} = #{|dx| >= t} + #{|dy| >= t} - #{both}
// with the first two from prefix sums of the per-column / per-row agent histograms and the last from the (at
// most four) corner rectangles of side <= S - t.  All integers: results are identical to the pairwise sweep.
// The host picks a method per threshold (BOX_*) and only enables the path when its cost per agent is far below
// the pairwise sweep's; otherwise (and for every vf_size step) the tiled pairwise kernel runs.
// ------------------------------------------------------------------------------------------------
enum { BOX_ZERO = 0, BOX_ALL = 1, BOX_DIRECT = 2, BOX_COMPLEMENT = 3 };
struct BoxPlan {
  int enabled;       // 0: pairwise kernels compute the counts
  int need_hist;     // some threshold uses BOX_COMPLEMENT
  int mode[4];       // for t_rep, t_att, t_repf, t_attf
  int t[4];
  int fast_rmax;     // >= 0: word-aligned rows (S % 32 == 0) and every BOX_DIRECT radius t-1 <= fast_rmax <= 15:
                     // all direct thresholds of an agent are counted in ONE branch-free sweep of 2*fast_rmax+1 rows
                     // (32-bit window per row by funnel shift); -1: generic box_count per threshold
};


// words of the padded per-env bitmap the fused reward launch counts boxes on (fused_step.cuh, box_near_counts):
// rB zero rows above and below the grid, one zero word after every row
__host__ __device__ inline int box_near_words(int S, int rB) { return ((S + 2 * rB) * ((S >> 5) + 1) + 3) / 4 * 4; }

// ----------------------------------------------------------------------------------------------
// Kernel parameter block (by-value kernel argument).
// ----------------------------------------------------------------------------------------------
struct StepParams {
  int E, N, S;
  // Integer thresholds, all clamped to [0,S].  With c(t) = #{j != i : D_ij < t} (D = plain Chebyshev):
  //   #{D < rep}              = c(t_rep)            #{S-D < rep}  = (N-1) - c(t_repf),  t_repf = S-rep+1
  //   #{(D<att)==(D>=rep)}    = |c(t_att)-c(t_rep)| #{XNOR on S-D} = |c(t_repf)-c(t_attf)|, t_attf = S-att+1
  int t_rep, t_att, t_repf, t_attf;
  int diag_rep, diag_att;  // contribution of the diagonal (D=0, S-D=S) to the literal N x N sums
  int t_vf;                // D <= vf  <=>  D < t_vf ; -1: no visual-field gate
  int far_border, far_band;  // both far thresholds > S/2: far pairs only among agents within far_band of a border
  double w_att, w_rep, w_ali, inv_nn;  // reward_type weights, 1/(N(N-1))
  float eat;
  int K, PL, KP, R;
  int auto_reset, track_stats;
  int bitmap_words;  // per-env smem occupancy bitmap words (0: not used)
  // counter-based tapes: when a tape pointer is NULL the draw is computed in the kernel by
  // Philox4x32-10 keyed with rng_seed; the host can materialise the identical tapes (tapes.philox_*).
  unsigned rng_seed_lo, rng_seed_hi;   // step streams (retarget, slab): counter (env, rng_step, block, 1)
  unsigned rng_step;
  unsigned rst_seed_lo, rst_seed_hi;   // reset streams (place: stream 2, ptape: stream 3): counter (env, reset index, block, stream)
  unsigned env_offset;                 // added to the env index in every Philox counter (swarm_config.env_offset)
  int one;           // runtime 1: `v * one + acc` compiles to IMAD (FMA pipe) instead of IADD3 (ALU pipe)
  int vf_dense;      // SWARM_FLAG_VF_DENSE
  int split;         // lane-group kernels: the step as two launches (transition, reward); see fused_step_kernel
  int tiny_scalar;   // SWARM_FLAG_TINY_SCALAR: the scalar thread-per-env kernel even where the packed one applies
  int box_near;      // reward launch: {rep, att} counts from box popcounts on the occupancy bitmap (fused_step.cuh)
  // state
  int16_t *x, *y, *px, *py;
  int32_t* target;
  uint8_t *orient, *frozen;
  uint32_t *err, *reset_count;
  int16_t* tcell;                      // [E,2] cell of the predator's target (may be NULL)
  uint32_t *ep_rec, *fin_count, *fin_len;   // ep_rec [E,4]: {length, f32 bits of attraction / repulsion / alignment}
  float* fin_ret;
  // tapes
  const int16_t *place, *ptape;
  const uint8_t *actions, *retarget, *slab;
  // outputs
  uint8_t* done;
  int32_t* eaten;
  float* reward;
  int32_t* counts;
  int16_t *pred_prev, *pred_next;
  uint8_t* step_orient;
  int32_t *step_target, *consumed;
  int16_t *term_x, *term_y;
  float* agent_reward;
  int32_t* agent_counts;
  uint8_t* eff_action;
  int16_t* agent_terms;   // [E,N,4] (rew_rep_i, rew_attr_i, U_i = 2 vis - P, Q_i)
  uint8_t* eff_buf;       // two-launch fused step: effective actions from the transition launch to the reward launch
                          // (= eff_action when the caller binds it, else library scratch)
};

// ----------------------------------------------------------------------------------------------
// Counter-based tapes: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// A draw is a pure function of (seed, stream, env, index), so the kernels need no tape memory and no RNG state;
// gym-swarm_b200/tapes.py holds the NumPy twin that materialises the same values for the oracle.
//   stream 1 (step n):    block 0 word 0 -> retarget flag (w < 0.1 * 2^32);  block 1 + k/16, byte k%16 -> slab[k] & 7
//   stream 2 (reset r):   value i of the placement tape = (w * S) >> 32 with w = word i%4 of block i/4
//   stream 3 (reset r):   predator attempt a = (x, y) from words 2(a%2), 2(a%2)+1 of block a/2
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ unsigned word_of(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

struct StepParams;
__device__ __forceinline__ int tape_retarget(const StepParams& p, int env);
__device__ __forceinline__ int tape_slab(const StepParams& p, const uint8_t* __restrict__ row, int env, int d);
__device__ __forceinline__ int tape_place(const StepParams& p, const int16_t* __restrict__ row, int env, unsigned rc, int i);
__device__ __forceinline__ void tape_pred(const StepParams& p, const int16_t* __restrict__ row, int env, unsigned rc, int a,
                                          int& px, int& py);

// ----------------------------------------------------------------------------------------------
// Action table SDE:467-475 as 2-bit LUTs: (d+1) for action a sits at bits [2a, 2a+2).
//   dx: -1 -1  0  1  1  1  0 -1  0      dy:  0 -1 -1 -1  0  1  1  1  0
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int move_dx(int a) { return (int)((0x11A90u >> (2 * a)) & 3u) - 1; }
__device__ __forceinline__ int move_dy(int a) { return (int)((0x1A901u >> (2 * a)) & 3u) - 1; }

// One-step torus wrap, valid because |move| <= 1 (SDE:52-56, 70-79).
__device__ __forceinline__ int wrap1(int p, int S) { return p > S - 1 ? 0 : (p < 0 ? S - 1 : p); }

// Inverse action table (SDE:161-163): nibble index (mx+1)*3 + (my+1) -> action id.
__device__ __forceinline__ int move_to_action(int mx, int my) {
  return (int)((0x543682701ULL >> (4 * ((mx + 1) * 3 + (my + 1)))) & 0xFULL);
}

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// Predator pursuit SDE:142-165 given the (old) target position.  Returns new position + orientation.
__device__ __forceinline__ void pursue(int px, int py, int tx, int ty, int S, int& npx, int& npy, int& orient) {
  const int cx = px - tx, cy = py - ty;                    // SDE:142 signed coordinate distance
  int mx = cx >= 1 ? -1 : (cx <= -1 ? 1 : 0);              // SDE:145-149
  int my = cy >= 1 ? -1 : (cy <= -1 ? 1 : 0);
  if (max(cx, cy) > (S >> 1)) { mx = -mx; my = -my; }      // SDE:152-159 SIGNED max, floor(S/2)
  orient = move_to_action(mx, my);                         // SDE:161-163
  npx = wrap1(px + mx, S);                                 // SDE:164
  npy = wrap1(py + my, S);
}

// Nearest-neighbour keys (SDE:111-131, canonical stable-sort tie rule):
//   lo = lowest index at minimum distance, hi = highest; target = lo if lo==hi or 2*dmin < S else hi.
__device__ __forceinline__ unsigned nn_key_lo(int d, int i) { return ((unsigned)d << 16) | (unsigned)i; }
__device__ __forceinline__ unsigned nn_key_hi(int d, int i) { return ((unsigned)d << 16) | (unsigned)(0xFFFF - i); }
__device__ __forceinline__ int nn_pick(unsigned klo, unsigned khi, int S) {
  const int dmin = (int)(klo >> 16);
  const int lo = (int)(klo & 0xFFFFu);
  const int hi = 0xFFFF - (int)(khi & 0xFFFFu);
  return (lo == hi || 2 * dmin < S) ? lo : hi;
}

// ----------------------------------------------------------------------------------------------
// Packed 16x2 Chebyshev / threshold machinery (Blackwell DPX: VIADDMNMX.S16x2[.RELU]).
// A "pair word" carries two agents j0 (low half) and j1 (high half):
//   {xj2, yj2, nxj2, nyj2} = packed x, y, -x, -y.     Own agent i: XI = x_i in both halves, etc.
//   m = min(xj-xi, xi-xj, yj-yi, yi-yj) = -D per half          (1 VIADD.16x2 + 3 VIADDMNMX.S16x2)
//   [D < t] = clamp(t + m, 0, 1) per half                       (1 VIADDMNMX.S16x2.RELU)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned rep2(int v) { return __byte_perm((unsigned)v, 0u, 0x1010); }   // one PRMT
__device__ __forceinline__ unsigned pack2(int lo, int hi) { return ((unsigned)lo & 0xFFFFu) | ((unsigned)hi << 16); }

__device__ __forceinline__ unsigned neg_cheb2(unsigned XI, unsigned YI, unsigned NXI, unsigned NYI, const uint4& w) {
  unsigned m = __vadd2(w.x, NXI);
  m = __viaddmin_s16x2(XI, w.z, m);
  m = __viaddmin_s16x2(w.y, NYI, m);
  m = __viaddmin_s16x2(YI, w.w, m);
  return m;
}
__device__ __forceinline__ unsigned lt2(unsigned T2, unsigned m) { return __viaddmin_s16x2_relu(T2, m, 0x00010001u); }
__device__ __forceinline__ int hsum2(unsigned v) { return (int)(v & 0xFFFFu) + (int)(v >> 16); }

// Unit-vector decomposition of an action for the alignment term (SDE:346-347: cosine_distances/2).
//   even ids are axis moves, odd ids diagonal moves, 8 = no move (zero vector: cos = 0 with everyone).
//   cos(a_i,a_j) = u_i . u_j  with u = m/|m|;   the sqrt(2)/2 factors are kept symbolic:
//   cosSum_i = sum_j cos = P_i/2 + (sqrt2/2) Q_i  with integers
//     axis i:  P_i = 2 (mx Ax + my Ay),  Q_i = mx Dx + my Dy
//     diag i:  P_i =    mx Dx + my Dy,   Q_i = mx Ax + my Ay
//   (Ax,Ay) / (Dx,Dy) = summed moves of the axis / diagonal movers that are counted for agent i.
__device__ __forceinline__ void align_PQ(int a, int Ax, int Ay, int Dx, int Dy, int self_gate, int& P, int& Q) {
  if (a >= 8) { P = 0; Q = 0; return; }
  const int mx = move_dx(a), my = move_dy(a);
  const int sa = mx * Ax + my * Ay, sd = mx * Dx + my * Dy;
  if (a & 1) { P = sd - 2 * self_gate; Q = sa; }      // minus own contribution (u_i.u_i = 1 -> P += -2)
  else       { P = 2 * sa - 2 * self_gate; Q = sd; }
}

// Per-agent reward epilogue (SDE:371-379) from integer pieces.
//   cR=c(t_rep) cA=c(t_att) cRf=c(t_repf) cAf=c(t_attf) (off-diagonal counts),
//   vis = number of j != i counted for alignment, P,Q as above.
struct AgentTerms {
  int rep_i;    // rew_rep_i  = -(sum R1[i,:] + sum R2[i,:] - 1)      SDE:372
  int att_i;    // rew_attr_i =   sum A1[i,:] + sum A2[i,:]          SDE:373
  double una;   // sum_j unalign[i,j]                                SDE:374
};
// compact per-agent credit (swarm_step_out.agent_terms): una = 0.25 * U - (sqrt2/4) * Q with U = 2 vis - P
__device__ __forceinline__ uint2 pack_terms(int rep_i, int att_i, int vis, int P, int Q) {
  return make_uint2(pack2(rep_i, att_i), pack2(2 * vis - P, Q));
}
__device__ __forceinline__ AgentTerms agent_terms(const StepParams& p, int cR, int cA, int cRf, int cAf, int vis,
                                                  int P, int Q) {
  AgentTerms t;
  const int offR = cR + (p.N - 1) - cRf;
  const int offA = iabs(cA - cR) + iabs(cRf - cAf);
  t.rep_i = -(offR + p.diag_rep - 1);
  t.att_i = offA + p.diag_att;
  t.una = 0.5 * (double)vis - 0.25 * (double)P - 0.35355339059327376220 * (double)Q;
  return t;
}

// ----------------------------------------------------------------------------------------------
// Packed predator arithmetic shared by the predator kernels (staged.cuh, predator_async.cuh) and the packed
// thread-per-env step (tiny_packed.cuh): positions stay two agents per 32-bit word.
// ----------------------------------------------------------------------------------------------
// 8 packed agents against the predator's NEW cell (kx, ky = cell replicated in both halves): lowest hit index.
// z has a zero halfword iff that agent stands on the cell; the exact index is only resolved on the rare hit.
__device__ __forceinline__ unsigned zero_half(unsigned z) { return (z - 0x00010001u) & ~z & 0x80008000u; }
__device__ __forceinline__ void hit8(const uint4& xv, const uint4& yv, unsigned kx, unsigned ky, int i0, unsigned& hit) {
  const unsigned z[4] = {(xv.x ^ kx) | (yv.x ^ ky), (xv.y ^ kx) | (yv.y ^ ky), (xv.z ^ kx) | (yv.z ^ ky),
                         (xv.w ^ kx) | (yv.w ^ ky)};
  if ((zero_half(z[0]) | zero_half(z[1]) | zero_half(z[2]) | zero_half(z[3])) == 0u) return;
#pragma unroll
  for (int hh = 3; hh >= 0; --hh) {
    if ((z[hh] >> 16) == 0u) hit = min(hit, (unsigned)(i0 + 2 * hh + 1));
    if ((z[hh] & 0xFFFFu) == 0u) hit = min(hit, (unsigned)(i0 + 2 * hh));
  }
}

// Chebyshev distance of two packed agents to the predator, per halfword: max(x-px, px-x, y-py, py-y).
//   NP = rep2(-px), P1 = rep2(px+1) (so that ~x + P1 = px - x), same for y.   4 DPX + 2 LOP per word.
__device__ __forceinline__ unsigned cheb2_pred(unsigned xw, unsigned yw, unsigned NP, unsigned P1, unsigned NQ, unsigned Q1) {
  unsigned d = __vadd2(xw, NP);
  d = __viaddmax_s16x2(~xw, P1, d);
  d = __viaddmax_s16x2(yw, NQ, d);
  d = __viaddmax_s16x2(~yw, Q1, d);
  return d;
}
// Nearest-neighbour keys of one vector (8 agents) in packed 16-bit form, (d << 5) | c with d < 2048 and the local index
// l = 8 q + e < 32 of the agent in its lane: c = l selects the LOWEST index at the minimum distance (HI = false),
// c = 31 - l the HIGHEST (SDE:121-129 under the stable-sort rule, see nn_pick).
template <int Q, bool HI>
__device__ __forceinline__ void nn_keys8_packed(const uint4& xv, const uint4& yv, unsigned NP, unsigned P1, unsigned NQ,
                                                unsigned Q1, unsigned& acc) {
  const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
  unsigned k[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const unsigned d = cheb2_pred(xs[h], ys[h], NP, P1, NQ, Q1);
    const unsigned l0 = (unsigned)(Q * 8 + 2 * h);
    k[h] = (d << 5) + (HI ? (((31u - (l0 + 1u)) << 16) | (31u - l0)) : (((l0 + 1u) << 16) | l0));
  }
  acc = __vimin3_u16x2(acc, k[0], k[1]);
  acc = __vimin3_u16x2(acc, k[2], k[3]);
}
// running packed minimum of (x ^ kx) | (y ^ ky): a zero halfword <=> an agent of the vector stands on the cell
__device__ __forceinline__ void hit8_packed(const uint4& xv, const uint4& yv, unsigned kx, unsigned ky, unsigned& acc) {
  acc = __vimin3_u16x2(acc, (xv.x ^ kx) | (yv.x ^ ky), (xv.y ^ kx) | (yv.y ^ ky));
  acc = __vimin3_u16x2(acc, (xv.z ^ kx) | (yv.z ^ ky), (xv.w ^ kx) | (yv.w ^ ky));
}


// number of set bits in the bit range [lo, hi] (inclusive, lo <= hi) of a word array
__device__ __forceinline__ int range_popc(const unsigned* __restrict__ bits, int lo, int hi) {
  const int w0 = lo >> 5, w1 = hi >> 5;
  const unsigned m0 = 0xFFFFFFFFu << (lo & 31), m1 = 0xFFFFFFFFu >> (31 - (hi & 31));
  if (w0 == w1) return __popc(bits[w0] & m0 & m1);
  int c = __popc(bits[w0] & m0) + __popc(bits[w1] & m1);
  for (int w = w0 + 1; w < w1; ++w) c += __popc(bits[w]);
  return c;
}
// agents in the rectangle [x0,x1] x [y0,y1] (empty when a bound is inverted)
__device__ __forceinline__ int box_popc(const unsigned* __restrict__ bits, int S, int x0, int x1, int y0, int y1) {
  if (x0 > x1 || y0 > y1) return 0;
  int c = 0;
  for (int y = y0; y <= y1; ++y) c += range_popc(bits, y * S + x0, y * S + x1);
  return c;
}
// inclusive prefix of a histogram with the out-of-range conventions P(a < 0) = 0, P(a >= S-1) = N
__device__ __forceinline__ int hist_le(const int* __restrict__ pre, int a, int S, int N) {
  return a < 0 ? 0 : (a >= S - 1 ? N : pre[a]);
}
__device__ __forceinline__ int box_count(const BoxPlan& bp, int s, const unsigned* __restrict__ bits, const int* __restrict__ hx,
                                         const int* __restrict__ hy, int xi, int yi, int S, int N) {
  const int t = bp.t[s];
  switch (bp.mode[s]) {
    case BOX_ZERO: return 0;
    case BOX_ALL: return N;
    case BOX_DIRECT: {
      const int r = t - 1;
      return box_popc(bits, S, max(0, xi - r), min(S - 1, xi + r), max(0, yi - r), min(S - 1, yi + r));
    }
    default: {
      const int nxs = hist_le(hx, xi - t, S, N) + N - hist_le(hx, xi + t - 1, S, N);
      const int nys = hist_le(hy, yi - t, S, N) + N - hist_le(hy, yi + t - 1, S, N);
      int both = 0;
      if (nxs > 0 && nys > 0) {
        both = box_popc(bits, S, 0, xi - t, 0, yi - t) + box_popc(bits, S, 0, xi - t, yi + t, S - 1) +
               box_popc(bits, S, xi + t, S - 1, 0, yi - t) + box_popc(bits, S, xi + t, S - 1, yi + t, S - 1);
      }
      return N - (nxs + nys - both);
    }
  }
}


// All BOX_DIRECT thresholds of one agent in a single sweep over the rows yi-r .. yi+r (r = bp.fast_rmax, uniform over
// the warp, so the loop and the per-threshold row tests do not diverge).  Per row: two shared-memory words, one funnel
// shift to a 32-bit window that starts at the leftmost in-grid column of the widest box, one AND + POPC per threshold.
__device__ __forceinline__ void box_direct_sweep(const BoxPlan& bp, const unsigned* __restrict__ bits, int xi, int yi, int S,
                                                 int (&cnt)[4]) {
  const int r = bp.fast_rmax;
  const int wpr = S >> 5;                                   // words per row
  const int lo_c = max(xi - r, 0);
  const int sh = lo_c & 31, w0 = lo_c >> 5, w1 = min(w0 + 1, wpr - 1);
  unsigned mq[4];
  int rq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    rq[q] = bp.mode[q] == BOX_DIRECT ? bp.t[q] - 1 : -1;
    const int lq = max(xi - rq[q], 0), hq = min(xi + rq[q], S - 1);
    mq[q] = rq[q] >= 0 ? ((0xFFFFFFFFu >> (31 - (hq - lq))) << (lq - lo_c)) : 0u;   // box columns inside the window
    cnt[q] = 0;
  }
  for (int dy = -r; dy <= r; ++dy) {
    const int y = yi + dy;
    const bool valid = y >= 0 && y < S;
    const int yb = min(max(y, 0), S - 1) * wpr;
    const unsigned a = bits[yb + w0], b = bits[yb + w1];
    const unsigned win = valid ? __funnelshift_r(a, b, sh) : 0u;
    const int ady = dy < 0 ? -dy : dy;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (ady <= rq[q]) cnt[q] += __popc(win & mq[q]);      // uniform branch: rq and dy are the same for every lane
  }
}

// ---- episode statistics: one 16-byte running record per env (swarm_state.ep_rec), finished sums updated on done ------
// Called by the single writer of the env.  rec = the record loaded at the top of the step (so that the load shares the
// step's first DRAM round trip).  The survival term only exists on the terminal step (SDE:321-327).
__device__ __forceinline__ void episode_update(const StepParams& p, int env, uint4 rec, bool done, float r_surv, float r_att,
                                               float r_rep, float r_ali) {
  const unsigned len = rec.x + 1u;
  const float att = __uint_as_float(rec.y) + r_att, rep = __uint_as_float(rec.z) + r_rep, ali = __uint_as_float(rec.w) + r_ali;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(p.ep_rec) + env;
  if (done) {                                    // fire-and-forget reductions: no read round trip
    atomicAdd(&p.fin_count[env], 1u);
    atomicAdd(&p.fin_len[env], len);
    atomicAdd(&p.fin_ret[4 * env + 0], r_surv);
    atomicAdd(&p.fin_ret[4 * env + 1], att);
    atomicAdd(&p.fin_ret[4 * env + 2], rep);
    atomicAdd(&p.fin_ret[4 * env + 3], ali);
    *out = make_uint4(0u, 0u, 0u, 0u);
  } else {
    *out = make_uint4(len, __float_as_uint(att), __float_as_uint(rep), __float_as_uint(ali));
  }
}

// ---- tape accessors: a bound tape (row pointer) or the in-kernel Philox stream -----------------------------------
// The Philox paths are separate NON-INLINED functions: inlined, their ~60 instructions and live registers were paid by
// every kernel even with bound tapes (tiny_step_kernel 96 -> 112 registers, 87 -> 99 us at cfg5).
static __device__ __noinline__ int philox_retarget(unsigned k0, unsigned k1, unsigned step, int env) {
  return philox4x32_10(make_uint4((unsigned)env, step, 0u, 1u), k0, k1).x < 429496730u;
}
static __device__ __noinline__ int philox_slab(unsigned k0, unsigned k1, unsigned step, int env, int d) {
  const uint4 v = philox4x32_10(make_uint4((unsigned)env, step, 1u + (unsigned)(d >> 4), 1u), k0, k1);
  return (int)((word_of(v, (d >> 2) & 3) >> (8 * (d & 3))) & 7u);
}
static __device__ __noinline__ int philox_place(unsigned k0, unsigned k1, int S, int env, unsigned rc, int i) {
  const uint4 v = philox4x32_10(make_uint4((unsigned)env, rc, (unsigned)(i >> 2), 2u), k0, k1);
  return (int)__umulhi(word_of(v, i & 3), (unsigned)S);
}
static __device__ __noinline__ unsigned philox_pred(unsigned k0, unsigned k1, int S, int env, unsigned rc, int a) {
  const uint4 v = philox4x32_10(make_uint4((unsigned)env, rc, (unsigned)(a >> 1), 3u), k0, k1);
  return pack2((int)__umulhi((a & 1) ? v.z : v.x, (unsigned)S), (int)__umulhi((a & 1) ? v.w : v.y, (unsigned)S));
}

__device__ __forceinline__ int tape_retarget(const StepParams& p, int env) {
  if (p.retarget) return p.retarget[env];
  return philox_retarget(p.rng_seed_lo, p.rng_seed_hi, p.rng_step, env + (int)p.env_offset);
}
__device__ __forceinline__ int tape_slab(const StepParams& p, const uint8_t* __restrict__ row, int env, int d) {
  if (p.slab) return row[d] & 7;
  return philox_slab(p.rng_seed_lo, p.rng_seed_hi, p.rng_step, env + (int)p.env_offset, d);
}
__device__ __forceinline__ const int16_t* place_row(const StepParams& p, int env, unsigned rc) {
  return p.place ? p.place + ((size_t)(rc % (unsigned)p.R) * (size_t)p.E + (size_t)env) * (size_t)p.PL : nullptr;
}
__device__ __forceinline__ const int16_t* ptape_row(const StepParams& p, int env, unsigned rc) {
  return p.ptape ? p.ptape + ((size_t)(rc % (unsigned)p.R) * (size_t)p.E + (size_t)env) * (size_t)p.KP * 2 : nullptr;
}
__device__ __forceinline__ int tape_place(const StepParams& p, const int16_t* __restrict__ row, int env, unsigned rc, int i) {
  if (row) return row[i];
  return philox_place(p.rst_seed_lo, p.rst_seed_hi, p.S, env + (int)p.env_offset, rc, i);
}
__device__ __forceinline__ void tape_pred(const StepParams& p, const int16_t* __restrict__ row, int env, unsigned rc, int a,
                                          int& px, int& py) {
  if (row) { px = row[2 * a]; py = row[2 * a + 1]; return; }
  const unsigned v = philox_pred(p.rst_seed_lo, p.rst_seed_hi, p.S, env + (int)p.env_offset, rc, a);
  px = (int)(v & 0xFFFFu); py = (int)(v >> 16);
}

}  // namespace swarm
