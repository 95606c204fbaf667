// fused_step.cuh -- the fused warp-group transition kernel for N <= 256 (sm_100a).
//
// One environment is owned by a group of LPE lanes (LPE in {4,8,16,32}; 32/LPE envs per warp), each
// lane holding APL consecutive agents in registers (blocked layout -> vectorised int16 / u8 HBM
// access).  A single pass over HBM performs, per env (reference: swarm_discrete_env.py "SDE"):
//   move + wrap (SDE:237-244, 60-80) -> collision detect + re-draw rounds (SDE:246-255, 269-292)
//   -> predator retarget/pursuit on the OLD positions (SDE:258-259, 111-165)
//   -> termination test on the NEW positions (SDE:307-309, 321-327)
//   -> pairwise attraction / repulsion counts + alignment (SDE:330-379), N-body style: the env's
//      positions are staged in shared memory as packed 16x2 "pair words" and every lane sweeps them
//      with DPX VIADDMNMX.S16x2 ops, two pairs per instruction
//   -> reward epilogue, episode statistics, in-kernel auto-reset (SDE:197-221, 90-109).
// All cross-lane traffic is warp shuffles restricted to the env's lane group; no __syncthreads.
#pragma once
#include "swarm_common.cuh"

#ifndef FUSED_WARPS
#define FUSED_WARPS 4    // warps per CTA of the fused kernels
#endif
#ifndef FUSED_MINB
#define FUSED_MINB 8     // minimum resident CTAs per SM the step kernel is compiled for (caps registers at 64; measured
                         // on cfg3, ms per step: round 1 8 warps x 2 CTAs 0.339, 4 x 5 0.328, 4 x 6 0.316; round 2 (per-env
                         // scalars stored before the pair sweep) 4 x 6 0.283, 4 x 7 0.280, 4 x 8 0.272)
#endif
#ifndef FUSED_P1_MINB
#define FUSED_P1_MINB 8  // the same for the transition launch of the two-launch step
#endif
#ifndef FUSED_VF_MINB
#define FUSED_VF_MINB 5  // the same for the sparse visual-field variant (hit bookkeeping needs a few more registers)
#endif

namespace swarm {

template <int LPE>
struct Seg {
  unsigned mask;
  int li, seg;
  __device__ __forceinline__ Seg() {
    const int lane = threadIdx.x & 31;
    seg = lane / LPE;
    li = lane % LPE;
    mask = (LPE == 32) ? 0xFFFFFFFFu : (((1u << LPE) - 1u) << (seg * LPE));
  }
  // group reductions: one REDUX instruction (works for any lane mask) instead of a log2(LPE)-step shuffle butterfly
  __device__ __forceinline__ int sum(int v) const { return __reduce_add_sync(mask, v); }
  __device__ __forceinline__ unsigned umin(unsigned v) const { return __reduce_min_sync(mask, v); }
  // exclusive prefix sum over the group's lanes; total = group sum
  __device__ __forceinline__ int excl_scan(int v, int& total) const {
    int inc = v;
#pragma unroll
    for (int d = 1; d < LPE; d <<= 1) {
      const int t = __shfl_up_sync(mask, inc, d, LPE);
      if (li >= d) inc += t;
    }
    total = __shfl_sync(mask, inc, LPE - 1, LPE);
    return inc - v;
  }
  __device__ __forceinline__ bool any(bool pred) const { return (__ballot_sync(mask, pred) & mask) != 0u; }
  __device__ __forceinline__ unsigned bcast(unsigned v, int src) const { return __shfl_sync(mask, v, src, LPE); }
#ifdef SWARM_RACE_JITTER
  // Race canary build (tools/race_canary.py): after every lane-group barrier the lanes are pushed out of lockstep by
  // lane- and call-dependent sleeps, so that an exchange through shared memory that lacks its barrier -- and only works
  // because a converged warp happens to run in lockstep -- produces results that differ from the production build.
  mutable unsigned jit = 0x9E3779B9u;
  __device__ __forceinline__ void sync() const {
    __syncwarp(mask);
    jit = jit * 1664525u + 1013904223u + (unsigned)li * 0x85EBCA6Bu;
    if ((jit >> 29) < 3u) __nanosleep(64u + ((jit >> 20) & 255u));
  }
#else
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
#endif
};

// ---- vectorised per-lane loads / stores of APL consecutive elements -------------------------------
template <int APL>
__device__ __forceinline__ void load_i16(const int16_t* __restrict__ p, int i0, int N, bool vec, int (&out)[APL]) {
  if (vec && i0 < N) {     // N % APL == 0 and i0 % APL == 0: the whole vector is inside the env
    if constexpr (APL == 1) {
      out[0] = p[i0];
    } else if constexpr (APL == 2) {
      const unsigned v = *reinterpret_cast<const unsigned*>(p + i0);
      out[0] = (int16_t)(v & 0xFFFF); out[1] = (int16_t)(v >> 16);
    } else if constexpr (APL == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(p + i0);
      out[0] = (int16_t)(v.x & 0xFFFF); out[1] = (int16_t)(v.x >> 16);
      out[2] = (int16_t)(v.y & 0xFFFF); out[3] = (int16_t)(v.y >> 16);
    } else {
      static_assert(APL == 8, "APL");
      const uint4 v = *reinterpret_cast<const uint4*>(p + i0);
      out[0] = (int16_t)(v.x & 0xFFFF); out[1] = (int16_t)(v.x >> 16);
      out[2] = (int16_t)(v.y & 0xFFFF); out[3] = (int16_t)(v.y >> 16);
      out[4] = (int16_t)(v.z & 0xFFFF); out[5] = (int16_t)(v.z >> 16);
      out[6] = (int16_t)(v.w & 0xFFFF); out[7] = (int16_t)(v.w >> 16);
    }
  } else {
#pragma unroll
    for (int k = 0; k < APL; ++k) out[k] = (i0 + k < N) ? (int)p[i0 + k] : 0;
  }
}

template <int APL>
__device__ __forceinline__ void store_i16(int16_t* __restrict__ p, int i0, int N, bool vec, const int (&v)[APL]) {
  if (vec && i0 < N) {
    if constexpr (APL == 1) {
      p[i0] = (int16_t)v[0];
    } else if constexpr (APL == 2) {
      *reinterpret_cast<unsigned*>(p + i0) = pack2(v[0], v[1]);
    } else if constexpr (APL == 4) {
      *reinterpret_cast<uint2*>(p + i0) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
    } else {
      *reinterpret_cast<uint4*>(p + i0) =
          make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    }
  } else {
#pragma unroll
    for (int k = 0; k < APL; ++k)
      if (i0 + k < N) p[i0 + k] = (int16_t)v[k];
  }
}

template <int APL>
__device__ __forceinline__ void load_u8(const uint8_t* __restrict__ p, int i0, int N, bool vec, int (&out)[APL]) {
  if (vec && i0 < N) {
    if constexpr (APL == 1) {
      out[0] = p[i0];
    } else if constexpr (APL == 2) {
      const unsigned v = *reinterpret_cast<const uint16_t*>(p + i0);
      out[0] = v & 0xFF; out[1] = v >> 8;
    } else if constexpr (APL == 4) {
      const unsigned v = *reinterpret_cast<const unsigned*>(p + i0);
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = (v >> (8 * k)) & 0xFF;
    } else {
      const uint2 v = *reinterpret_cast<const uint2*>(p + i0);
#pragma unroll
      for (int k = 0; k < 4; ++k) { out[k] = (v.x >> (8 * k)) & 0xFF; out[4 + k] = (v.y >> (8 * k)) & 0xFF; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < APL; ++k) out[k] = (i0 + k < N) ? (int)p[i0 + k] : 8;
  }
}

template <int APL>
__device__ __forceinline__ void store_u8(uint8_t* __restrict__ p, int i0, int N, bool vec, const int (&v)[APL]) {
  if (vec && i0 < N) {
    if constexpr (APL == 1) {
      p[i0] = (uint8_t)v[0];
    } else if constexpr (APL == 2) {
      *reinterpret_cast<uint16_t*>(p + i0) = (uint16_t)(v[0] | (v[1] << 8));
    } else if constexpr (APL == 4) {
      *reinterpret_cast<unsigned*>(p + i0) = (unsigned)(v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24));
    } else {
      *reinterpret_cast<uint2*>(p + i0) = make_uint2((unsigned)(v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24)),
                                                     (unsigned)(v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24)));
    }
  } else {
#pragma unroll
    for (int k = 0; k < APL; ++k)
      if (i0 + k < N) p[i0 + k] = (uint8_t)v[k];
  }
}

// ---- exact occupancy counts c_i (SDE:269-292) by an all-pairs sweep over smem keys ---------------
// skey: this env's NMAX-word key region.  c[k] = number of agents on agent k's cell (1 = alone).
template <int LPE, int APL>
__device__ __forceinline__ void occupancy_counts(const Seg<LPE>& g, unsigned* skey, int N, int i0,
                                                 const int (&cx)[APL], const int (&cy)[APL], int (&c)[APL]) {
  unsigned key[APL];
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    key[k] = (i0 + k < N) ? pack2(cx[k], cy[k]) : (0xFFFF0000u | (unsigned)(i0 + k));
    skey[i0 + k] = key[k];
    c[k] = 0;
  }
  g.sync();
  for (int j = 0; j < N; ++j) {
    const unsigned kj = skey[j];
#pragma unroll
    for (int k = 0; k < APL; ++k) c[k] += (kj == key[k]);
  }
#pragma unroll
  for (int k = 0; k < APL; ++k)
    if (i0 + k >= N) c[k] = 1;
  g.sync();
}

// ---- exact occupancy counts from a per-env occupancy bitmap in smem (S*S bits) -----------------------
// Every agent sets its cell's bit with atomicOr; an agent that finds the bit already set is a "hit".  A cell
// holding c agents produces exactly c-1 hits (all but the first arrival), so
//     c_i = 1 + #{hit agents standing on agent i's cell}.
// The (few) hit agents broadcast their cell one after the other and everybody compares: O(hits * APL) per lane
// instead of the O(N * APL) all-pairs sweep.  Returns false (c untouched) when no cell is shared.
template <int LPE, int APL>
__device__ __forceinline__ bool bitmap_counts(const Seg<LPE>& g, unsigned* sbit, int N, int S, int i0,
                                              const int (&cx)[APL], const int (&cy)[APL], int (&c)[APL]) {
  unsigned pending = 0u;
  int word[APL];
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    word[k] = -1;
    if (i0 + k < N) {
      const int cell = cy[k] * S + cx[k];
      word[k] = cell >> 5;
      const unsigned bit = 1u << (cell & 31);
      const unsigned old = atomicOr(&sbit[word[k]], bit);
      if (old & bit) pending |= 1u << k;
    }
  }
  const bool any = g.any(pending != 0u);   // ballot also orders the atomics before the clears below
  g.sync();
#pragma unroll
  for (int k = 0; k < APL; ++k)
    if (word[k] >= 0) sbit[word[k]] = 0u;
  g.sync();
  if (!any) return false;
  unsigned key[APL];
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    key[k] = (i0 + k < N) ? pack2(cx[k], cy[k]) : (0xFFFF0000u | (unsigned)(i0 + k));
    c[k] = 1;
  }
  const int lane = threadIdx.x & 31;
  for (;;) {
    const unsigned ball = __ballot_sync(g.mask, pending != 0u) & g.mask;
    if (ball == 0u) break;
    const int src = __ffs(ball) - 1;
    const int kk = __ffs(pending) - 1;           // meaningful in the src lane only
    unsigned mine = 0u;
#pragma unroll
    for (int k = 0; k < APL; ++k)
      if (k == kk) mine = key[k];
    const unsigned hk = __shfl_sync(g.mask, mine, src);
#pragma unroll
    for (int k = 0; k < APL; ++k) c[k] += (key[k] == hk);
    if (lane == src) pending &= pending - 1u;
  }
  return true;
}

// ---- nearest agent to (qx,qy) with the canonical tie rule (SDE:111-131) --------------------------
template <int LPE, int APL>
__device__ __forceinline__ int closest_target(const Seg<LPE>& g, int N, int S, int i0, int qx, int qy,
                                              const int (&ax)[APL], const int (&ay)[APL]) {
  unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    if (i0 + k < N) {
      const int d = max(iabs(qx - ax[k]), iabs(qy - ay[k]));
      klo = min(klo, nn_key_lo(d, i0 + k));
      khi = min(khi, nn_key_hi(d, i0 + k));
    }
  }
  klo = g.umin(klo);
  khi = g.umin(khi);
  return nn_pick(klo, khi, S);
}

// ---- reference reset (SDE:197-221 + Predator.__init__ SDE:90-109) for one env group ---------------
template <int LPE, int APL>
__device__ __noinline__ void reset_group(const StepParams& p, const Seg<LPE>& g, int env, unsigned* skey, unsigned* sbit,
                                         int i0, int (&x)[APL], int (&y)[APL], int& px, int& py, int& target,
                                         unsigned& err) {
  const int N = p.N, S = p.S;
  unsigned rc = p.reset_count[env];
  g.sync();   // every lane has read reset_count before lane 0 bumps it
  if (g.li == 0) p.reset_count[env] = rc + 1u;
  const int16_t* __restrict__ tape = place_row(p, env, rc);
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const bool v = i0 + k < N;
    x[k] = v ? tape_place(p, tape, env, rc, i0 + k) : 0;          // SDE:202-203: all x's then all y's
    y[k] = v ? tape_place(p, tape, env, rc, N + i0 + k) : 0;
  }
  int cur = 2 * N;
  for (;;) {
    int c[APL];
    if (sbit) {                                   // (clean) per-env bitmap of the step kernel: O(hits) instead of O(N)
      if (!bitmap_counts<LPE, APL>(g, sbit, N, S, i0, x, y, c)) break;
    } else {
      occupancy_counts<LPE, APL>(g, skey, N, i0, x, y, c);
    }
    int extra = 0;
#pragma unroll
    for (int k = 0; k < APL; ++k) extra += c[k] - 1;
    int L;
    int off = g.excl_scan(extra, L);
    if (L == 0) break;
    if (cur + 2 * L > p.PL) { err |= SWARM_ERR_PLACE_OVERFLOW; break; }
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      const int e = c[k] - 1;
      if (e > 0) {                                // SDE:208-209: (2,L) block, last duplicate wins
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
  for (;;) {                                      // SDE:92-105
    if (a >= p.KP) { err |= SWARM_ERR_PRED_OVERFLOW; break; }
    tape_pred(p, pt, env, rc, a, px, py);
    ++a;
    bool hit = false;
#pragma unroll
    for (int k = 0; k < APL; ++k) hit |= (i0 + k < N) && x[k] == px && y[k] == py;
    if (!g.any(hit)) break;
  }
  target = closest_target<LPE, APL>(g, N, S, i0, px, py, x, y);   // SDE:109
}

// ---- symmetric ring sweep (see the comment at its call site) ---------------------------------------------------
// FAR = false: the two far thresholds were already counted by the border pass (far_border_counts), the sweep carries
// only the {rep, att} counter pair.
//
// VIS = true (visual-field steps, sparse mode): every round also records WHICH of its pairs are within the visual
// field: bit W * APL - 1 - (w * APL + k) of the low / high half of a hit word stands for the visiting agents 2w / 2w+1
// against the owned agent k.  Non-zero hit words (few) are appended, tagged with round and lane, to the env's list
// shits[0, nhit) by a ballot compaction; the list is turned into the gated alignment sums after the sweep
// (vf_sparse_accumulate).  The dense alternative evaluates every ordered pair (kernel, VFM == 1).
template <int LPE, int APL, bool FAR, bool VIS>
__device__ __forceinline__ void ring_sweep(const Seg<LPE>& g, const uint4* __restrict__ spw, const unsigned (&XI)[APL],
                                           const unsigned (&YI)[APL], const unsigned (&NXI)[APL], const unsigned (&NYI)[APL],
                                           unsigned TR, unsigned TA, unsigned TRf, unsigned TAf, unsigned one, unsigned c256,
                                           unsigned (&aR)[APL], unsigned (&aA)[APL], unsigned (&aRf)[APL], unsigned (&aAf)[APL],
                                           unsigned TV, unsigned* __restrict__ shits, int& nhit) {
  constexpr int W = APL / 2;
  constexpr int HALF = LPE / 2;
  unsigned hits = 0u;
  const unsigned two = one + one;
  const unsigned lt_mask = (1u << (threadIdx.x & 31)) - 1u;
  nhit = 0;
  unsigned tag = (unsigned)g.li << 24;                   // {hit bits | round << 8 | lane << 24}
  auto push_hits = [&](int) {                            // (called once per round, in round order)
    const bool nz = hits != 0u;
    const unsigned b = __ballot_sync(g.mask, nz) & g.mask;
    if (nz) shits[nhit + __popc(b & lt_mask)] = hits | tag;
    nhit += __popc(b);
    tag += 256u;
    hits = 0u;
  };
  // Counters are BYTE packed: a 16x2 indicator word holds 0/1 in bytes 0 and 2, so ind(t0) + 256 * ind(t1)
  // carries two thresholds for two agents in one register ({t0,t1} = {rep,att} and {rep_far,att_far}); it is
  // built once (IMAD) and added to the row and to the column counter.  No byte can overflow: a row byte
  // meets at most NMAX / 2 <= 64 pair words, a column byte at most (LPE / 2 - 1) * APL <= 60 rows.
  unsigned row[APL][2], col[W][2];
#pragma unroll
  for (int k = 0; k < APL; ++k) row[k][0] = row[k][1] = 0u;
#pragma unroll
  for (int w = 0; w < W; ++w) col[w][0] = col[w][1] = 0u;
  auto rows_only = [&](int src) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const uint4 vw = spw[w * LPE + src];
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], vw);
        row[k][0] = lt2(TR, m) * one + row[k][0];
        row[k][0] = lt2(TA, m) * c256 + row[k][0];
        if constexpr (FAR) {
          row[k][1] = lt2(TRf, m) * one + row[k][1];
          row[k][1] = lt2(TAf, m) * c256 + row[k][1];
        }
        if constexpr (VIS) hits = hits * two + lt2(TV, m);   // (FMA pipe) test n = w * APL + k ends up in bit W * APL - 1 - n
      }
    }
  };
  rows_only(g.li);                                       // round 0: own block
  if constexpr (VIS) {
    unsigned self = 0u;                                  // an agent is not its own neighbour
#pragma unroll
    for (int k = 0; k < APL; ++k) self |= 1u << (W * APL - 1 - ((k >> 1) * APL + k) + 16 * (k & 1));
    hits &= ~self;
    push_hits(0);
  }
  const int nxt = (g.li + 1) & (LPE - 1);
#pragma unroll 1
  for (int r = 1; r < HALF; ++r) {
    const int src = (g.li + r) & (LPE - 1);
    if (r > 1) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        col[w][0] = __shfl_sync(g.mask, col[w][0], nxt, LPE);
        if constexpr (FAR) col[w][1] = __shfl_sync(g.mask, col[w][1], nxt, LPE);
      }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const uint4 vw = spw[w * LPE + src];
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], vw);
        const unsigned n0 = lt2(TA, m) * c256 + lt2(TR, m);
        row[k][0] = n0 * one + row[k][0];
        col[w][0] = n0 * one + col[w][0];
        if constexpr (FAR) {
          const unsigned n1 = lt2(TAf, m) * c256 + lt2(TRf, m);
          row[k][1] = n1 * one + row[k][1];
          col[w][1] = n1 * one + col[w][1];
        }
        if constexpr (VIS) hits = hits * two + lt2(TV, m);   // (FMA pipe) test n = w * APL + k ends up in bit W * APL - 1 - n
      }
    }
    if constexpr (VIS) push_hits(r);
  }
  // send the column counters home: lane l holds those of block l + HALF - 1
  if constexpr (HALF > 1) {
    const int from = (g.li - (HALF - 1)) & (LPE - 1);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      col[w][0] = __shfl_sync(g.mask, col[w][0], from, LPE);
      if constexpr (FAR) col[w][1] = __shfl_sync(g.mask, col[w][1], from, LPE);
    }
  }
  rows_only((g.li + HALF) & (LPE - 1));                  // round LPE/2: met from both sides
  if constexpr (VIS) push_hits(HALF);
  // fold: total = both agents' bytes of the row counter + this agent's bytes of the column counter
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const int sh = 16 * (k & 1);
    aR[k] = __dp4a(row[k][0], 0x00010001u, 0u) + ((col[k >> 1][0] >> sh) & 0xFFu);
    aA[k] = __dp4a(row[k][0], 0x01000100u, 0u) + ((col[k >> 1][0] >> (sh + 8)) & 0xFFu);
    if constexpr (FAR) {
      aRf[k] = __dp4a(row[k][1], 0x00010001u, 0u) + ((col[k >> 1][1] >> sh) & 0xFFu);
      aAf[k] = __dp4a(row[k][1], 0x01000100u, 0u) + ((col[k >> 1][1] >> (sh + 8)) & 0xFFu);
    }
  }
}

// ---- visual-field alignment from the recorded hits (sparse mode) ------------------------------------------------
// The env's lanes share the hit list: entry -> (lane and round that saw it) -> owned agent i, visiting agent j per set
// bit.  i receives j's move vector and one visible neighbour; in the symmetric rounds 1 .. LPE/2 - 1 the pair was
// evaluated once only, so j receives i's as well.  sacc = {count[NMAX], axis sums[NMAX], diagonal sums[NMAX]}
// (shared-memory atomics: a hit is rare -- (2 vf + 1)^2 / S^2 of the pairs).  svf[j] = j's move as two biased 16x2
// words (axis, diagonal).
template <int LPE, int APL>
__device__ __forceinline__ void vf_sparse_accumulate(const Seg<LPE>& g, const unsigned* __restrict__ shits, int nhit,
                                                     const uint2* __restrict__ svf, unsigned* __restrict__ sacc, int N) {
  constexpr int NMAX = LPE * APL;
  constexpr int HALF = LPE / 2;
  for (int e = g.li; e < nhit; e += LPE) {
    const unsigned ent = shits[e];
    const int r = (int)((ent >> 8) & 31u), ln = (int)(ent >> 24);
    unsigned h = ent & 0x00FF00FFu;
    const int ib = ln * APL, j0 = ((ln + r) & (LPE - 1)) * APL;
    const bool both = r > 0 && r < HALF;
    while (h) {
      const int b = __ffs(h) - 1;
      h &= h - 1u;
      const int t = (APL / 2) * APL - 1 - (b & 15);
      const int i = ib + (t & (APL - 1));
      const int j = j0 + 2 * (t / APL) + (b >> 4);
      if (i < N && j < N) {                                // (two padding agents share a cell)
        const uint2 uj = svf[j];
        atomicAdd(&sacc[i], 1u);
        atomicAdd(&sacc[NMAX + i], uj.x);
        atomicAdd(&sacc[2 * NMAX + i], uj.y);
        if (both) {
          const uint2 ui = svf[i];
          atomicAdd(&sacc[j], 1u);
          atomicAdd(&sacc[NMAX + j], ui.x);
          atomicAdd(&sacc[2 * NMAX + j], ui.y);
        }
      }
    }
  }
}

// ---- near thresholds from the env's occupancy bitmap (reward launch of the two-launch step) ------------------------
// After the collision rounds the N cells are distinct, so #{j : D_ij < t} (self included) is the popcount of the S x S
// occupancy bitmap over the (2t-1)^2 box around agent i, clipped to the grid (plain Chebyshev, no wrap).  The reward
// launch rebuilds the bitmap in a PADDED layout: rB = max radius zero rows above and below the grid and one zero word
// after every row (stride S/32 + 1 words, S % 32 == 0), so that a row's box columns always sit in the 32-bit window
// obtained by a funnel shift of two consecutive words and no row or word needs a range check.  Per (row, agent):
// 2 LDS + SHF + LOP + POPC + 2 IADD, against ~9 instructions per PAIR WORD of the ring sweep: 9 rows x 4 agents per lane
// instead of 136 word evaluations at N = S = 128, att = 5.  (In the single-launch kernel a box variant bought nothing
// while the instruction cache was the limit -- round 1 measured 0.270 vs 0.275 ms; in the reward launch, which is issue
// bound, instructions are time.)  The odd row stride also spreads the gathers over the banks.
template <int APL>
__device__ __forceinline__ void box_near_counts(const unsigned* __restrict__ bits, int S, int rR, int rA,
                                                const int (&x)[APL], const int (&y)[APL], unsigned (&cR)[APL],
                                                unsigned (&cA)[APL]) {
  const int rB = max(max(rR, rA), 0), rS = min(rR, rA);   // uniform over the warp (kernel parameters); rB = layout radius
  const int stride = (S >> 5) + 1;
  const unsigned* ptr[APL];
  unsigned sh[APL], mB[APL], mS[APL], cB[APL], cS[APL];
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const int lo_c = max(x[k] - rB, 0);
    sh[k] = (unsigned)(lo_c & 31);
    ptr[k] = bits + y[k] * stride + (lo_c >> 5);         // padded row y = grid row y - rB: the first row of the box
    // the boxes' columns inside the window that starts at column lo_c (a radius of -1 is the empty box of t = 0)
    const int hB = min(x[k] + rB, S - 1), lS = max(x[k] - rS, 0), hS = min(x[k] + rS, S - 1);
    mB[k] = max(rR, rA) >= 0 ? (0xFFFFFFFFu >> (31 - (hB - lo_c))) : 0u;
    mS[k] = rS >= 0 ? ((0xFFFFFFFFu >> (31 - (hS - lS))) << (lS - lo_c)) : 0u;
    cB[k] = 0u;
    cS[k] = 0u;
  }
#pragma unroll 1
  for (int i = 0; i <= 2 * rB; ++i) {
    const bool small = i >= rB - rS && i <= rB + rS;     // uniform: rows that belong to the smaller box as well
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      const unsigned win = __funnelshift_r(ptr[k][0], ptr[k][1], sh[k]);
      cB[k] += __popc(win & mB[k]);
      if (small) cS[k] += __popc(win & mS[k]);
      ptr[k] += stride;
    }
  }
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    cA[k] = rA >= rR ? cB[k] : cS[k];
    cR[k] = rA >= rR ? cS[k] : cB[k];
  }
}

// ---- far thresholds from the border agents only ------------------------------------------------------------------
// For a threshold t > S/2 a pair with D >= t needs |dx| >= t or |dy| >= t, i.e. one agent within b = S - t cells of a
// border and the other within b cells of the OPPOSITE border.  So #{j : D_ij < t} = N-1 for every agent outside the
// border band, and only LOW-band x HIGH-band candidate pairs have to be looked at, per axis:
//   X pass: i in LX (x < b), j in HX (x >= S-b):  far  <=>  xj - xi >= t
//   Y pass: i in LY (y < b), j in HY (y >= S-b):  far  <=>  yj - yi >= t  and not |xj - xi| >= t   (a pair far apart in
//           both axes sits in opposite X bands as well and was counted by the X pass)
// ~ (N b / S)^2 candidates per axis instead of the (4 N b / S)^2 of an all-band x all-band loop (16 + 16 instead of 240
// at N = S = 128, att = 5).  Runs after the ring sweep: the positions are read from the staged pair words, the lists
// hold agent indices.  sfar (per env): [0,4) list lengths nLX nHX nLY nHY; [4, 4+NMAX) per-agent counters
// (#far(t_repf) | #far(t_attf) << 16); then two u16 arrays of NMAX entries: the X lists (LX grows from the front, HX
// from the back -- an agent is in at most one of them) and the Y lists.
// Outputs aRf / aAf in the sweep's convention (self included): N - #{j != i : D_ij >= t}.
template <int LPE, int APL, bool RING_LAYOUT = true>
__device__ __forceinline__ void far_border_counts(const StepParams& p, const Seg<LPE>& g, const uint4* __restrict__ spw,
                                                  unsigned* __restrict__ sfar, int N, int S, int i0,
                                                  const unsigned (&XI)[APL], const unsigned (&YI)[APL],
                                                  unsigned (&aRf)[APL], unsigned (&aAf)[APL]) {
  constexpr int NMAX = LPE * APL;
  unsigned* __restrict__ ctr = sfar + 4;
  unsigned short* __restrict__ xl = reinterpret_cast<unsigned short*>(sfar + 4 + NMAX);
  unsigned short* __restrict__ yl = xl + NMAX;
  const int b = p.far_band;
  if (g.li < 4) sfar[g.li] = 0u;
#pragma unroll
  for (int k = 0; k < APL; ++k) ctr[i0 + k] = 0u;
  g.sync();
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const int x = (int)(XI[k] & 0xFFFFu), y = (int)(YI[k] & 0xFFFFu);
    if (i0 + k < N && (x < b || x >= S - b || y < b || y >= S - b)) {      // ~ 4 b / S of the agents
      const unsigned short id = (unsigned short)(i0 + k);
      if (x < b) xl[atomicAdd(&sfar[0], 1u)] = id;
      else if (x >= S - b) xl[NMAX - 1 - atomicAdd(&sfar[1], 1u)] = id;
      if (y < b) yl[atomicAdd(&sfar[2], 1u)] = id;
      else if (y >= S - b) yl[NMAX - 1 - atomicAdd(&sfar[3], 1u)] = id;
    }
  }
  g.sync();
  const int nLX = (int)sfar[0], nHX = (int)sfar[1], nLY = (int)sfar[2], nHY = (int)sfar[3];
  // candidate (ii, jj) of a pass = (q >> sh, q & (2^sh - 1)) with 2^sh >= its high-band list (no division; at most half
  // of the slots of a pass are padding)
  const int shX = 32 - __clz(max(nHX, 1) - 1), shY = 32 - __clz(max(nHY, 1) - 1);
  const int totX = nHX ? nLX << shX : 0, tot = totX + (nHY ? nLY << shY : 0);
  for (int q = g.li; q < tot; q += LPE) {
    const bool yp = q >= totX;
    const int qq = yp ? q - totX : q, sh = yp ? shY : shX, nh = yp ? nHY : nHX;
    const int ii = qq >> sh, jj = qq & ((1 << sh) - 1);
    if (jj >= nh) continue;
    const unsigned short* __restrict__ lst = yp ? yl : xl;
    const int a = lst[ii], c = lst[NMAX - 1 - jj];
    // agent n sits in half n % 2 of pair word (n % APL / 2) * LPE + n / APL (the ring sweep's [word][lane] layout) or of
    // pair word n / 2 (the row sweep's linear layout)
    const uint4 wa = spw[RING_LAYOUT ? ((a % APL) >> 1) * LPE + a / APL : a >> 1];
    const uint4 wc = spw[RING_LAYOUT ? ((c % APL) >> 1) * LPE + c / APL : c >> 1];
    const int sa = 16 * (a & 1), sc = 16 * (c & 1);
    const int xa = (int)((wa.x >> sa) & 0xFFFFu), ya = (int)((wa.y >> sa) & 0xFFFFu);
    const int xc = (int)((wc.x >> sc) & 0xFFFFu), yc = (int)((wc.y >> sc) & 0xFFFFu);
    const int dm = yp ? yc - ya : xc - xa;                 // along the pass axis (c is the high-band agent: dm > 0)
    const int dox = yp ? iabs(xc - xa) : -1;               // Y pass: the pair must not be far in x already
    const unsigned v = (unsigned)(dm >= p.t_repf && dox < p.t_repf) | ((unsigned)(dm >= p.t_attf && dox < p.t_attf) << 16);
    if (v) { atomicAdd(&ctr[a], v); atomicAdd(&ctr[c], v); }
  }
  g.sync();
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const unsigned v = ctr[i0 + k];
    aRf[k] = (unsigned)N - (v & 0xFFFFu);
    aAf[k] = (unsigned)N - (v >> 16);
  }
}

// ---- the fused step kernel ------------------------------------------------------------------------
template <int LPE, int APL>
struct FusedCfg {
  static constexpr int EPW = 32 / LPE;       // envs per warp
  static constexpr int NMAX = LPE * APL;     // agents per env capacity
  static constexpr int PW = NMAX / 2;        // pair words per env
  static constexpr int WARPS = FUSED_WARPS;  // warps per CTA
  // bytes of smem per warp: pair words (aliased by the collision keys) + VF words + bitmap
  __host__ __device__ static constexpr int pair_bytes() { return EPW * PW * 16; }
  // visual-field steps: VFM = 1 dense (move words of every agent), VFM = 2 sparse (+ three accumulators per agent and
  // one hit word per (round, lane) of the ring sweep)
  __host__ __device__ static constexpr bool ring_ok() { return (APL == 2 || APL == 4) && LPE >= 4; }
  __host__ __device__ static constexpr int hit_words() { return (LPE / 2 + 1) * LPE; }
  __host__ __device__ static constexpr int vf_env_bytes(int vfm) {
    return vfm == 0 ? 0 : vfm == 1 ? NMAX * 8 : NMAX * 20 + hit_words() * 4;
  }
  __host__ __device__ static constexpr int vf_bytes(int vfm) { return EPW * vf_env_bytes(vfm); }
  // scratch of the far-threshold border pass (ring kernels): 4 list lengths + NMAX counters + 2 x NMAX u16 list entries
  __host__ __device__ static constexpr int far_env_words() { return (ring_ok() || APL == 8) ? 4 + 2 * NMAX : 0; }
  __host__ __device__ static constexpr int far_bytes() { return EPW * far_env_words() * 4; }
};

// Terminal step of one env group (SDE:321-327) + auto-reset (SDE:197-221): rare (an env finishes every ~100 steps), so
// it lives out of line and its registers do not count against the hot path.
template <int LPE, int APL>
__device__ __noinline__ void terminal_step(const StepParams& p, const Seg<LPE>& g, int env, unsigned* skey, unsigned* sbit,
                                           int i0, int (&nx)[APL], int (&ny)[APL], int px, int py, int target, int orient,
                                           int eaten, uint4 ep0) {
  const int N = p.N;
  const size_t base = (size_t)env * (size_t)N;
  const bool vec = (N % APL) == 0;
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    if (i0 + k < N) {
      if (p.agent_reward)
        reinterpret_cast<float4*>(p.agent_reward)[base + i0 + k] = make_float4(0.f, 0.f, 0.f, (i0 + k == eaten) ? p.eat : 0.f);
      if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[base + i0 + k] = make_int2(0, 0);
      if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[base + i0 + k] = make_uint2(0u, 0u);
    }
  }
  if (g.li == 0) {
    float* rw = p.reward + 5 * (size_t)env;
    rw[0] = p.eat; rw[1] = 0.f; rw[2] = 0.f; rw[3] = 0.f; rw[4] = p.eat;
    p.counts[2 * env] = 0;
    p.counts[2 * env + 1] = 0;
    if (p.track_stats) episode_update(p, env, ep0, true, p.eat, 0.f, 0.f, 0.f);
  }
  if (p.term_x) {
    store_i16<APL>(p.term_x + base, i0, N, vec, nx);
    store_i16<APL>(p.term_y + base, i0, N, vec, ny);
  }
  unsigned err = 0;
  if (p.auto_reset) {
    reset_group<LPE, APL>(p, g, env, skey, sbit, i0, nx, ny, px, py, target, err);
    orient = 0;                                             // SDE:107
  } else if (g.li == 0) {
    p.frozen[env] = 1;
  }
  store_i16<APL>(p.x + base, i0, N, vec, nx);
  store_i16<APL>(p.y + base, i0, N, vec, ny);
  if (g.li == 0) {
    p.px[env] = (int16_t)px;
    p.py[env] = (int16_t)py;
    p.target[env] = target;
    p.orient[env] = (uint8_t)orient;
    if (err) p.err[env] |= err;
  }
}

// PHASE 0: the whole step in one launch.  PHASE 1 / 2: the same step as two launches -- 1 = transition (move, collision
// re-draws, predator, termination, commit, terminal step + auto-reset), 2 = reward of the running episodes, reading the
// committed positions and the effective actions phase 1 left in p.eff_buf.  One launch is the shorter latency (small
// batches, CUDA-graph rollouts); for batches of many waves the split wins because the step's straight-line code
// (~48 KB executed per env) no longer fits the SM's 32 KB instruction cache when 32 resident warps sit in unrelated
// phases of it (ncu on the fused kernel at cfg3: "no instruction" was the top stall reason, 3.1 of 7.2 warps per
// scheduler), while each half does.
template <int LPE, int APL, int VFM, int PHASE>
__global__ void __launch_bounds__(FusedCfg<LPE, APL>::WARPS * 32,
                                  PHASE == 1 ? (APL == 8 ? 4 : FUSED_P1_MINB)
                                             : (APL == 8 || VFM == 1) ? 3 : (VFM == 2 ? FUSED_VF_MINB : FUSED_MINB))
fused_step_kernel(const __grid_constant__ StepParams p) {
  using C = FusedCfg<LPE, APL>;
  constexpr bool HAS_VF = VFM != 0;
  constexpr bool VF_SPARSE = VFM == 2;
  static_assert(!VF_SPARSE || C::ring_ok(), "sparse visual-field mode rides on the ring sweep");
  constexpr bool RING = VFM != 1 && C::ring_ok();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const Seg<LPE> g;
  const int env = (blockIdx.x * C::WARPS + warp) * C::EPW + g.seg;
  const int N = p.N, S = p.S;

  // smem carve-up (per warp, then per env group): pair words (aliased by the collision keys) | visual-field words |
  // border-pass scratch | occupancy bitmap
  const int warp_bytes = C::pair_bytes() + C::vf_bytes(VFM) + C::far_bytes() + C::EPW * p.bitmap_words * 4;
  unsigned char* wbase = smem_raw + (size_t)warp * warp_bytes;
  uint4* spw = reinterpret_cast<uint4*>(wbase) + g.seg * C::PW;
  unsigned* skey = reinterpret_cast<unsigned*>(spw);   // aliases the pair words (NMAX words <= PW*4 words)
  unsigned char* vbase = wbase + C::pair_bytes() + g.seg * C::vf_env_bytes(VFM);
  uint2* svf = reinterpret_cast<uint2*>(vbase);
  unsigned* sacc = reinterpret_cast<unsigned*>(vbase + C::NMAX * 8);   // sparse mode: count / axis / diagonal sums
  unsigned* shits = sacc + 3 * C::NMAX;                                 // sparse mode: hit words of the sweep
  unsigned* sfar = reinterpret_cast<unsigned*>(wbase + C::pair_bytes() + C::vf_bytes(VFM)) + g.seg * C::far_env_words();
  unsigned* sbit = reinterpret_cast<unsigned*>(wbase + C::pair_bytes() + C::vf_bytes(VFM) + C::far_bytes()) +
                   g.seg * p.bitmap_words;
  if ((PHASE != 2 || (VFM == 0 && p.box_near)) && p.bitmap_words) {   // (a multiple of 4 words, 16-byte aligned)
    for (int w = g.li; w < (p.bitmap_words >> 2); w += LPE) reinterpret_cast<uint4*>(sbit)[w] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (env >= p.E) return;
  g.sync();

  const int i0 = g.li * APL;
  const size_t base = (size_t)env * (size_t)N;
  const bool vec = (N % APL) == 0;
  int nx[APL], ny[APL];
  unsigned effp = 0u;                                  // the effective actions of the owned agents, 4 bits each

  if constexpr (PHASE != 2) {
  // every load the first phases need is issued here, before the first dependent branch, so that one DRAM round trip
  // covers all of it
  const int is_frozen = p.frozen[env];
  int x[APL], y[APL], act[APL];
  load_i16<APL>(p.x + base, i0, N, vec, x);
  load_i16<APL>(p.y + base, i0, N, vec, y);
  load_u8<APL>(p.actions + base, i0, N, vec, act);
  int px = p.px[env], py = p.py[env];
  int target = p.target[env];
  // the first 16 re-draw actions of the env (one broadcast LDG.128) ride on the same round trip: a colliding env
  // (~45 % of them at cfg3, ~1 draw each) would otherwise pay a second, dependent trip to DRAM in the middle of the step
  const uint8_t* __restrict__ slab = p.slab ? p.slab + (size_t)env * (size_t)p.K : nullptr;
  const bool slab_pre = slab != nullptr && p.K >= 16 && (((uintptr_t)slab) & 15) == 0;
  uint4 slab16 = make_uint4(0u, 0u, 0u, 0u);
  if (slab_pre) slab16 = *reinterpret_cast<const uint4*>(slab);

  if (is_frozen) {                          // SDE:233-234: finished episode, auto-reset off: nothing moves, every output
    if (g.li == 0) {                        // of the step is defined (zero reward, the predator where it stands)
      p.err[env] |= SWARM_ERR_STEP_WHEN_DONE;
      p.done[env] = 1;
      p.eaten[env] = -1;
      p.counts[2 * env] = 0; p.counts[2 * env + 1] = 0;
      p.consumed[env] = 0;
#pragma unroll
      for (int q = 0; q < 5; ++q) p.reward[5 * env + q] = 0.f;
      p.pred_prev[2 * env] = (int16_t)px; p.pred_prev[2 * env + 1] = (int16_t)py;
      p.pred_next[2 * env] = (int16_t)px; p.pred_next[2 * env + 1] = (int16_t)py;
      p.step_orient[env] = p.orient[env];
      p.step_target[env] = target;
    }
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      if (i0 + k < N) {
        if (p.agent_reward) reinterpret_cast<float4*>(p.agent_reward)[base + i0 + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[base + i0 + k] = make_int2(0, 0);
        if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[base + i0 + k] = make_uint2(0u, 0u);
        if (p.eff_action) p.eff_action[base + i0 + k] = 8;
      }
    }
    return;
  }

  const int retarget = tape_retarget(p, env);
  unsigned err = 0;

  // ---- move + collision re-draw rounds (SDE:237-255) --------------------------------------------
  int eff[APL];
  {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < APL; ++k) { bad |= act[k] > 8; eff[k] = min(act[k], 8); }
    if (g.any(bad)) err |= SWARM_ERR_BAD_ACTION;
  }
  int cur = 0;
  for (;;) {
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      nx[k] = wrap1(x[k] + move_dx(eff[k]), S);          // SDE:60-80
      ny[k] = wrap1(y[k] + move_dy(eff[k]), S);
    }
    int c[APL];
    if (p.bitmap_words) {
      if (!bitmap_counts<LPE, APL>(g, sbit, N, S, i0, nx, ny, c)) break;
    } else {
      occupancy_counts<LPE, APL>(g, skey, N, i0, nx, ny, c);
    }
    int extra = 0;
#pragma unroll
    for (int k = 0; k < APL; ++k) extra += c[k] - 1;
    int L;
    int off = g.excl_scan(extra, L);
    if (L == 0) break;                                   // change_position_id returned None
    if (cur + L > p.K) { err |= SWARM_ERR_SLAB_OVERFLOW; break; }
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      const int e = c[k] - 1;
      if (e > 0) {                                       // SDE:250-252: last of its c-1 draws wins
        const int d = cur + off + e - 1;
        eff[k] = (slab_pre && d < 16) ? (int)((word_of(slab16, d >> 2) >> (8 * (d & 3))) & 7u) : tape_slab(p, slab, env, d);
      }
      off += e;
    }
    cur += L;
  }

  // ---- predator (SDE:258-259, 133-165) on the OLD agent positions ------------------------------
  const int prev_px = px, prev_py = py;
  if (retarget) target = closest_target<LPE, APL>(g, N, S, i0, px, py, x, y);
  target = min(max(target, 0), N - 1);
  int orient;
  {
    const int lt = target / APL, kt = target % APL;
    unsigned mine = 0;
#pragma unroll
    for (int k = 0; k < APL; ++k)
      if (k == kt) mine = pack2(x[k], y[k]);
    const unsigned tp = g.bcast(mine, lt);
    pursue(px, py, (int)(int16_t)(tp & 0xFFFFu), (int)(int16_t)(tp >> 16), S, px, py, orient);
  }

  // ---- termination (SDE:307-309, 321-327) on the NEW positions ---------------------------------
  unsigned hit = 0xFFFFu;
#pragma unroll
  for (int k = 0; k < APL; ++k)
    if (i0 + k < N && nx[k] == px && ny[k] == py) hit = min(hit, (unsigned)(i0 + k));
  hit = g.umin(hit);
  const bool done = hit != 0xFFFFu;

  // ---- everything that does not depend on the reward is written NOW: the per-env scalars die here instead of
  // staying in registers across the pair sweep (which is what bounds the number of envs in flight per SM) ----------
  if (p.eff_action) store_u8<APL>(p.eff_action + base, i0, N, vec, eff);
  if (g.li == 0) {
    p.done[env] = done ? 1 : 0;
    p.eaten[env] = done ? (int)hit : -1;
    p.pred_prev[2 * env] = (int16_t)prev_px; p.pred_prev[2 * env + 1] = (int16_t)prev_py;
    p.pred_next[2 * env] = (int16_t)px; p.pred_next[2 * env + 1] = (int16_t)py;
    p.step_orient[env] = (uint8_t)orient;
    p.step_target[env] = target;
    p.consumed[env] = cur;
    if (err) p.err[env] |= err;
  }
  if (done) {                                               // rare, out of line
    uint4 ep0 = make_uint4(0u, 0u, 0u, 0u);
    if (p.track_stats && g.li == 0) ep0 = reinterpret_cast<const uint4*>(p.ep_rec)[env];
    terminal_step<LPE, APL>(p, g, env, skey, p.bitmap_words ? sbit : nullptr, i0, nx, ny, px, py, target, orient, (int)hit, ep0);
    return;
  }
  // commit of the running episode (SDE:260)
  store_i16<APL>(p.x + base, i0, N, vec, nx);
  store_i16<APL>(p.y + base, i0, N, vec, ny);
  if (g.li == 0) {
    p.px[env] = (int16_t)px;
    p.py[env] = (int16_t)py;
    p.target[env] = target;
    p.orient[env] = (uint8_t)orient;
  }
  if constexpr (PHASE == 1) {                          // hand the effective actions to the reward launch
    if (p.eff_buf != p.eff_action) store_u8<APL>(p.eff_buf + base, i0, N, vec, eff);
    return;
  } else {
#pragma unroll
    for (int k = 0; k < APL; ++k) effp |= (unsigned)eff[k] << (4 * k);
  }
  } else {                                             // PHASE 2: the reward launch picks the running episodes up
    if (p.done[env]) return;                           // finished (or frozen): phase 1 wrote the terminal outputs
    int e2[APL];
    load_i16<APL>(p.x + base, i0, N, vec, nx);         // committed NEW positions
    load_i16<APL>(p.y + base, i0, N, vec, ny);
    load_u8<APL>(p.eff_buf + base, i0, N, vec, e2);
#pragma unroll
    for (int k = 0; k < APL; ++k) effp |= (unsigned)e2[k] << (4 * k);
  }

  // ---- reward (SDE:330-379) ---------------------------------------------------------------------
  unsigned aR[APL], aA[APL], aRf[APL], aAf[APL];       // threshold counts of each owned agent, self included
  unsigned aV[APL], accA[APL], accD[APL];
#pragma unroll
  for (int k = 0; k < APL; ++k) aR[k] = aA[k] = aRf[k] = aAf[k] = aV[k] = accA[k] = accD[k] = 0u;
  {
    // stage the env's positions as packed pair words {x2, y2, -x2, -y2}; padding agents sit at x=-S
    unsigned XI[APL], YI[APL], NXI[APL], NYI[APL];
    {
      int sx[APL], sy[APL];
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const bool v = i0 + k < N;
        sx[k] = v ? nx[k] : -S;
        sy[k] = v ? ny[k] : 0;
      }
      if constexpr (APL == 1) {
        const unsigned mine = pack2(sx[0], sy[0]);
        const unsigned oth = __shfl_xor_sync(g.mask, mine, 1, LPE);
        if ((g.li & 1) == 0) {
          const int ox = (int)(int16_t)(oth & 0xFFFFu), oy = (int)(int16_t)(oth >> 16);
          spw[g.li >> 1] = make_uint4(pack2(sx[0], ox), pack2(sy[0], oy), pack2(-sx[0], -ox), pack2(-sy[0], -oy));
        }
      } else {
        // ring sweep: word w of every lane contiguous ([w][lane]) so that the rotated LDS.128 is conflict free
#pragma unroll
        for (int k = 0; k < APL; k += 2)
          spw[RING ? (k >> 1) * LPE + g.li : (i0 + k) >> 1] =
              make_uint4(pack2(sx[k], sx[k + 1]), pack2(sy[k], sy[k + 1]), pack2(-sx[k], -sx[k + 1]),
                         pack2(-sy[k], -sy[k + 1]));
      }
      if constexpr (HAS_VF) {
#pragma unroll
        for (int k = 0; k < APL; ++k) {
          const int a = (int)((effp >> (4 * k)) & 15u);
          const bool mover = (i0 + k < N) && a < 8;
          const int mx = mover ? move_dx(a) : 0, my = mover ? move_dy(a) : 0;
          const bool diag = (a & 1) != 0;
          svf[i0 + k] = make_uint2(pack2((diag ? 0 : mx) + 1, (diag ? 0 : my) + 1),
                                   pack2((diag ? mx : 0) + 1, (diag ? my : 0) + 1));
          if constexpr (VF_SPARSE) sacc[i0 + k] = sacc[C::NMAX + i0 + k] = sacc[2 * C::NMAX + i0 + k] = 0u;
        }
      }
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        XI[k] = rep2(sx[k]); YI[k] = rep2(sy[k]); NXI[k] = rep2(-sx[k]); NYI[k] = rep2(-sy[k]);
      }
    }
    g.sync();

    const unsigned TR = rep2(p.t_rep), TA = rep2(p.t_att), TRf = rep2(p.t_repf), TAf = rep2(p.t_attf);
    const unsigned TV = rep2(HAS_VF ? p.t_vf : 0);
    if constexpr (RING) {
      // ---- symmetric ring sweep: every unordered block pair is evaluated once --------------------
      // Lane l owns block B_l (APL agents = W pair words).  In round r it meets the block of lane l+r (pair words
      // read from smem), adds each indicator to its own ROW counters and to the visiting block's COLUMN counters;
      // the column counters travel round the ring by shuffles together with the block they belong to and are
      // sent home after the last symmetric round.  Rounds 0 (own block) and LPE/2 (met from both sides) only
      // count rows.  The accumulations are IMADs by a runtime 1 so that they issue on the FMA pipe and leave the
      // ALU pipe to the DPX ops.
      const unsigned one = (unsigned)p.one, c256 = (unsigned)p.one << 8;
      int nhit = 0;
      if (PHASE != 1 && VFM == 0 && p.box_near) {
        // reward launch: {rep, att} from box popcounts on the rebuilt (padded) occupancy bitmap; far thresholds below
        const int rB = max(max(p.t_rep, p.t_att) - 1, 0), bstride = (S >> 5) + 1;
        int bx[APL], by[APL];
#pragma unroll
        for (int k = 0; k < APL; ++k) {
          const bool v = i0 + k < N;
          bx[k] = v ? nx[k] : 0;             // padding agents read (in-range) rows too: the row loop stays convergent
          by[k] = v ? ny[k] : 0;
          if (v) atomicOr(&sbit[(ny[k] + rB) * bstride + (nx[k] >> 5)], 1u << (nx[k] & 31));
        }
        g.sync();
        box_near_counts<APL>(sbit, S, p.t_rep - 1, p.t_att - 1, bx, by, aR, aA);
      } else if (p.far_border)      // both far thresholds > S/2: they come from the border pass below, the sweep carries {rep, att}
        ring_sweep<LPE, APL, false, VF_SPARSE>(g, spw, XI, YI, NXI, NYI, TR, TA, TRf, TAf, one, c256, aR, aA, aRf, aAf, TV, shits, nhit);
      else
        ring_sweep<LPE, APL, true, VF_SPARSE>(g, spw, XI, YI, NXI, NYI, TR, TA, TRf, TAf, one, c256, aR, aA, aRf, aAf, TV, shits, nhit);
      if constexpr (VF_SPARSE) {
        g.sync();
        vf_sparse_accumulate<LPE, APL>(g, shits, nhit, svf, sacc, N);
        g.sync();
#pragma unroll
        for (int k = 0; k < APL; ++k) { aV[k] = sacc[i0 + k]; accA[k] = sacc[C::NMAX + i0 + k]; accD[k] = sacc[2 * C::NMAX + i0 + k]; }
      }
      if (p.far_border) far_border_counts<LPE, APL>(p, g, spw, sfar, N, S, i0, XI, YI, aRf, aAf);
    } else if (PHASE != 1 && VFM == 0 && APL == 8 && p.box_near && p.far_border) {
      // N = 129 .. 256 (8 agents per lane, no ring sweep; reward launch AND single launch -- the collision rounds left the
      // bitmap clean): all four counts come without a pair sweep --
      // {rep, att} from box popcounts on the rebuilt (padded) occupancy bitmap, the far thresholds from the border pass
      // (72 + ~30 instead of 8 x 128 word evaluations per lane at N = 256)
      const int rB = max(max(p.t_rep, p.t_att) - 1, 0), bstride = (S >> 5) + 1;
      int bx[APL], by[APL];
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const bool v = i0 + k < N;
        bx[k] = v ? nx[k] : 0;
        by[k] = v ? ny[k] : 0;
        if (v) atomicOr(&sbit[(ny[k] + rB) * bstride + (nx[k] >> 5)], 1u << (nx[k] & 31));
      }
      g.sync();
      box_near_counts<APL>(sbit, S, p.t_rep - 1, p.t_att - 1, bx, by, aR, aA);
      far_border_counts<LPE, APL, false>(p, g, spw, sfar, N, S, i0, XI, YI, aRf, aAf);
    } else {
      const int NW = (N + 1) >> 1;
#pragma unroll 2
      for (int q = 0; q < NW; ++q) {
        const uint4 w = spw[q];
        uint4 u = make_uint4(0, 0, 0, 0);
        if constexpr (VFM == 1) u = *reinterpret_cast<const uint4*>(&svf[2 * q]);
#pragma unroll
        for (int k = 0; k < APL; ++k) {
          const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], w);
          aR[k] += lt2(TR, m);
          aA[k] += lt2(TA, m);
          aRf[k] += lt2(TRf, m);
          aAf[k] += lt2(TAf, m);
          if constexpr (VFM == 1) {
            const unsigned gate = lt2(TV, m);
            const unsigned g0 = gate & 0xFFFFu, g1 = gate >> 16;
            aV[k] += gate;
            accA[k] += g0 * u.x + g1 * u.z;
            accD[k] += g0 * u.y + g1 * u.w;
          }
        }
      }
    }
  }
  // the running-episode record is fetched now (lane 0): its latency hides behind the epilogue arithmetic
  uint4 ep0 = make_uint4(0u, 0u, 0u, 0u);
  if (p.track_stats && g.li == 0) ep0 = reinterpret_cast<const uint4*>(p.ep_rec)[env];

  // alignment sums (SDE:346-359)
  int Ax = 0, Ay = 0, Dx = 0, Dy = 0;
  if constexpr (!HAS_VF) {
    int ax = 0, ay = 0, dx = 0, dy = 0;
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      const int a = (int)((effp >> (4 * k)) & 15u);
      if (i0 + k < N && a < 8) {
        const int mx = move_dx(a), my = move_dy(a);
        if (a & 1) { dx += mx; dy += my; } else { ax += mx; ay += my; }
      }
    }
    // two biased 16-bit fields per word keep the group reduction at 2 words
    const unsigned s1 = (unsigned)g.sum((int)pack2(ax + APL, ay + APL));
    const unsigned s2 = (unsigned)g.sum((int)pack2(dx + APL, dy + APL));
    Ax = (int)(s1 & 0xFFFFu) - C::NMAX; Ay = (int)(s1 >> 16) - C::NMAX;
    Dx = (int)(s2 & 0xFFFFu) - C::NMAX; Dy = (int)(s2 >> 16) - C::NMAX;
  }

  const int selfR = p.t_rep >= 1, selfA = p.t_att >= 1, selfRf = p.t_repf >= 1, selfAf = p.t_attf >= 1;
  const int selfV = VF_SPARSE ? 0 : HAS_VF ? (p.t_vf >= 1) : 1;   // (the sparse sums never include the agent itself)
  int s_rep = 0, s_att = 0, s_P = 0, s_Q = 0, s_vis = 0;
#pragma unroll
  for (int k = 0; k < APL; ++k) {
    const bool v = i0 + k < N;
    const int a = (int)((effp >> (4 * k)) & 15u);
    int vis, P, Q;
    if constexpr (HAS_VF) {
      const int vin = VF_SPARSE ? (int)aV[k] : hsum2(aV[k]);   // gated count (dense: incl. self)
      const int gAx = (int)(accA[k] & 0xFFFFu) - vin, gAy = (int)(accA[k] >> 16) - vin;
      const int gDx = (int)(accD[k] & 0xFFFFu) - vin, gDy = (int)(accD[k] >> 16) - vin;
      vis = vin - selfV;
      align_PQ(a, gAx, gAy, gDx, gDy, selfV, P, Q);
    } else {
      vis = N - 1;
      align_PQ(a, Ax, Ay, Dx, Dy, 1, P, Q);
    }
    const AgentTerms t = agent_terms(p, hsum2(aR[k]) - selfR, hsum2(aA[k]) - selfA, hsum2(aRf[k]) - selfRf,
                                     hsum2(aAf[k]) - selfAf, vis, P, Q);
    if (v) {
      s_rep += t.rep_i; s_att += t.att_i; s_P += P; s_Q += Q; s_vis += vis;
      if (p.agent_reward) {                                 // SDE:371-379
        const double a_rep = (p.w_rep * (double)t.rep_i) * p.inv_nn;
        const double a_att = (p.w_att * (double)t.att_i) * p.inv_nn;
        const double a_ali = (p.w_ali * (-t.una)) * p.inv_nn;
        reinterpret_cast<float4*>(p.agent_reward)[base + i0 + k] =
            make_float4((float)a_att, (float)a_rep, (float)a_ali, (float)(a_rep + a_att + a_ali));
      }
      if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[base + i0 + k] = make_int2(t.rep_i, t.att_i);
      if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[base + i0 + k] = pack_terms(t.rep_i, t.att_i, vis, P, Q);
    }
  }
  const int g_rep = g.sum(s_rep);
  const int g_att = g.sum(s_att);
  const int g_P = g.sum(s_P), g_Q = g.sum(s_Q);
  const int g_vis = HAS_VF ? g.sum(s_vis) : N * (N - 1);
  if (g.li == 0) {
    const double g_una = 0.5 * (double)g_vis - 0.25 * (double)g_P - 0.35355339059327376220 * (double)g_Q;
    const double r_rep = p.w_rep * (double)g_rep * p.inv_nn;             // SDE:363-366
    const double r_att = p.w_att * (double)g_att * p.inv_nn;
    const double r_ali = p.w_ali * (-g_una) * p.inv_nn;
    float* rw = p.reward + 5 * (size_t)env;
    rw[0] = 0.f; rw[1] = (float)r_att; rw[2] = (float)r_rep; rw[3] = (float)r_ali; rw[4] = (float)(r_rep + r_att + r_ali);
    p.counts[2 * env] = g_rep;
    p.counts[2 * env + 1] = g_att;
    if (p.track_stats) episode_update(p, env, ep0, false, 0.f, (float)r_att, (float)r_rep, (float)r_ali);
  }
}

// ---- reset kernel: every env (mask == nullptr) or the masked envs ----------------------------------
template <int LPE, int APL>
__global__ void __launch_bounds__(FusedCfg<LPE, APL>::WARPS * 32)
fused_reset_kernel(const __grid_constant__ StepParams p, const uint8_t* __restrict__ mask, int zero_counts) {
  using C = FusedCfg<LPE, APL>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const Seg<LPE> g;
  const int env = (blockIdx.x * C::WARPS + warp) * C::EPW + g.seg;
  if (env >= p.E) return;
  if (mask && !mask[env]) return;
  unsigned* skey = reinterpret_cast<unsigned*>(smem_raw) + (warp * C::EPW + g.seg) * C::NMAX;
  const int N = p.N;
  const int i0 = g.li * APL;
  const size_t base = (size_t)env * (size_t)N;
  const bool vec = (N % APL) == 0;
  if (zero_counts) {
    if (g.li == 0) p.reset_count[env] = 0u;
    g.sync();
  }
  int x[APL], y[APL], px, py, target;
  unsigned err = 0;
  reset_group<LPE, APL>(p, g, env, skey, nullptr, i0, x, y, px, py, target, err);
  store_i16<APL>(p.x + base, i0, N, vec, x);
  store_i16<APL>(p.y + base, i0, N, vec, y);
  if (g.li == 0) {
    p.px[env] = (int16_t)px;
    p.py[env] = (int16_t)py;
    p.target[env] = target;
    p.orient[env] = 0;
    p.frozen[env] = 0;
    if (zero_counts) p.err[env] = err;
    else if (err) p.err[env] |= err;              // masked re-reset: error bits stay sticky
    if (p.track_stats) {
      reinterpret_cast<uint4*>(p.ep_rec)[env] = make_uint4(0u, 0u, 0u, 0u);
      if (zero_counts) {
        p.fin_count[env] = 0u;
        p.fin_len[env] = 0u;
        reinterpret_cast<float4*>(p.fin_ret)[env] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

}  // namespace swarm
