// staged.cuh -- the staged transition for large N (N > 256, up to 16384 agents per env) and the
// standalone roofline kernels named by the C ABI (swarm_move / swarm_predator / swarm_pairwise).
//
//   K1  move_kernel             flat, 128-bit vectorised; HBM bound                    SDE:60-80, 467-475
//   K2  collide_kernel          CTA per env, exact occupancy counts via a global (L2-resident) hash table + block
//                               scan for the re-draw slots (grids too large for shared memory)   SDE:246-255, 269-292
//   K2b collide_bitmap_kernel   the same rounds on the env's occupancy bitmap in SHARED memory, positions in registers
//   K2c boxcount_kernel         reward threshold counts as box popcounts on that bitmap (also inside K2b)   SDE:330-343
//   K3  predator_{group,block}_kernel  lane group / CTA per env: shuffle arg-min, pursuit, termination test
//                                                                                       SDE:111-165, 307-309
//   K4  pairprep_kernel         packs positions into 16x2 "pair words", per-env alignment sums
//   K5s pairwise_sym_kernel     symmetric tile sweep (ring of 128-agent tiles per warp), DPX + IMAD counters  SDE:330-343
//   K5  pairwise_kernel<true>   vf_size variant: row sweep, j-chunks staged by TMA bulk copies (cp.async.bulk +
//                               mbarrier, double buffered)
//   K6a/b finalize_{agents,env}_kernel  flat per-agent epilogue + per-env reward / statistics   SDE:363-379, 260
//   K7  reset_kernel / K7b reset_bitmap_kernel  CTA per env (masked): reference reset       SDE:197-221, 90-109
//   observe_vf_kernel           visual-field windows (the policy-side output)
//
// (SDE = /root/reference/gym_swarm/envs/swarm_discrete_env.py)
#pragma once
#include "swarm_common.cuh"

namespace swarm {

constexpr unsigned HASH_EMPTY = 0xFFFFFFFFu;

// Scratch owned by the library for the staged path.
struct StagedScratch {
  int16_t *nx, *ny;        // [E,N] candidate / new positions
  uint8_t* eff;            // [E,N] effective action
  unsigned* hash;          // [grid][2*HS] {keys[HS], counts[HS]}
  unsigned* aslot;         // [grid][N] hash slot of each agent
  int log2hs;
  int env_grid;            // CTAs of the per-env kernels
  uint4* pairw;            // [E,NW] packed pair words
  uint2* vfw;              // [E,2*NW] (VF) packed unit-vector words
  unsigned* pcounts;       // [E,N,8] per-agent accumulators {cR,cA,cRf,cAf,cV,accA,accD,-}
  unsigned* counter;       // work counter of pairwise_sym_kernel (behind pcounts)
  int* envacc;             // [E,12] per-env integer sums {Ax,Ay,Dx,Dy, s_rep,s_att,s_P,s_Q,s_vis,-,-,-}
  int NW;
};

// ------------------------------------------------------------------------------------------------
// block-level helpers (blockDim.x multiple of 32, <= 1024)
// ------------------------------------------------------------------------------------------------
struct BlockScratch {
  int warp_vals[32];
  int bcast[4];
  unsigned ubcast[4];
};

__device__ __forceinline__ int block_excl_scan(int v, int& total, BlockScratch& bs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();
  if (lane == 31) bs.warp_vals[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < nw ? bs.warp_vals[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
      if (lane >= d) winc += t;
    }
    bs.warp_vals[lane] = winc - w;
    if (lane == 31) bs.bcast[0] = winc;
  }
  __syncthreads();
  total = bs.bcast[0];
  return bs.warp_vals[warp] + inc - v;
}

__device__ __forceinline__ int block_sum(int v, BlockScratch& bs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  __syncthreads();
  if (lane == 0) bs.warp_vals[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? bs.warp_vals[lane] : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, d);
    if (lane == 0) bs.bcast[0] = w;
  }
  __syncthreads();
  return bs.bcast[0];
}

__device__ __forceinline__ unsigned block_umin(unsigned v, BlockScratch& bs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
  __syncthreads();
  if (lane == 0) bs.warp_vals[warp] = (int)v;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < nw ? (unsigned)bs.warp_vals[lane] : 0xFFFFFFFFu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w = min(w, __shfl_xor_sync(0xFFFFFFFFu, w, d));
    if (lane == 0) bs.ubcast[0] = w;
  }
  __syncthreads();
  return bs.ubcast[0];
}

// ------------------------------------------------------------------------------------------------
// K1: flat move kernel.  8 agents per thread: 2 x LDG.128 (x, y) + 1 x LDG.64 (actions) in,
// 2 x STG.128 (+ optional STG.64 effective action) out.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void move8(const uint4& xv, const uint4& yv, const uint2& av, int S, uint4& ox, uint4& oy,
                                      uint2& oa, bool& bad) {
  const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
  unsigned rx[4], ry[4];
  unsigned ra[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    int nxv[2], nyv[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int e = 2 * h + s;
      int a = (int)(((e < 4 ? av.x : av.y) >> (8 * (e & 3))) & 0xFFu);
      bad |= a > 8;
      a = min(a, 8);
      ra[e >> 2] |= (unsigned)a << (8 * (e & 3));
      const int xo = (int)(int16_t)((xs[h] >> (16 * s)) & 0xFFFFu);
      const int yo = (int)(int16_t)((ys[h] >> (16 * s)) & 0xFFFFu);
      nxv[s] = wrap1(xo + move_dx(a), S);
      nyv[s] = wrap1(yo + move_dy(a), S);
    }
    rx[h] = pack2(nxv[0], nxv[1]);
    ry[h] = pack2(nyv[0], nyv[1]);
  }
  ox = make_uint4(rx[0], rx[1], rx[2], rx[3]);
  oy = make_uint4(ry[0], ry[1], ry[2], ry[3]);
  oa = make_uint2(ra[0], ra[1]);
}

__global__ void __launch_bounds__(256) move_kernel(long long total, int N, int S, const int16_t* __restrict__ x,
                                                   const int16_t* __restrict__ y, const uint8_t* __restrict__ act,
                                                   int16_t* __restrict__ ox, int16_t* __restrict__ oy,
                                                   uint8_t* __restrict__ oeff, uint32_t* __restrict__ err) {
  const long long nvec = total >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x) + v);
    const uint4 yv = __ldg(reinterpret_cast<const uint4*>(y) + v);
    const uint2 av = __ldg(reinterpret_cast<const uint2*>(act) + v);
    uint4 rx, ry;
    uint2 ra;
    bool bad = false;
    move8(xv, yv, av, S, rx, ry, ra, bad);
    reinterpret_cast<uint4*>(ox)[v] = rx;
    reinterpret_cast<uint4*>(oy)[v] = ry;
    if (oeff) reinterpret_cast<uint2*>(oeff)[v] = ra;
    if (bad && err) {
      for (int e = 0; e < 8; ++e)
        if (act[v * 8 + e] > 8) atomicOr(&err[(v * 8 + e) / N], SWARM_ERR_BAD_ACTION);
    }
  }
  // scalar tail (total not a multiple of 8)
  for (long long i = (nvec << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int a = act[i];
    if (a > 8) { if (err) atomicOr(&err[i / N], SWARM_ERR_BAD_ACTION); a = 8; }
    ox[i] = (int16_t)wrap1((int)x[i] + move_dx(a), S);
    oy[i] = (int16_t)wrap1((int)y[i] + move_dy(a), S);
    if (oeff) oeff[i] = (uint8_t)a;
  }
}

// ------------------------------------------------------------------------------------------------
// Exact occupancy counts (SDE:269-292) for one env inside a CTA: open-addressing hash table in
// global memory (L2 resident; HS = power of two >= 2N slots of {key, count}).  Thread t owns the
// contiguous agents [lo,hi) (keeps agent order for the re-draw prefix sum).  Each agent's slot is
// remembered in aslot[], so nothing probes the table after the insert phase.
// Returns the thread's sum of (c_i - 1); c_i is re-read from the table by the caller.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash_slot(unsigned key, int log2hs) { return (key * 2654435761u) >> (32 - log2hs); }

__device__ __forceinline__ unsigned hash_insert(unsigned* hkeys, unsigned* hcnt, unsigned hmask, int log2hs,
                                                unsigned key) {
  unsigned s = hash_slot(key, log2hs);
  for (;;) {
    const unsigned old = atomicCAS(&hkeys[s], HASH_EMPTY, key);
    if (old == HASH_EMPTY || old == key) { atomicAdd(&hcnt[s], 1u); return s; }
    s = (s + 1u) & hmask;
  }
}

struct EnvHash {
  unsigned *keys, *cnt, *aslot;
  unsigned mask;
  int log2hs;
};

// phase 1+2: insert all, sync, return this thread's extra = sum(c_i - 1)
template <typename PosFn>
__device__ __forceinline__ int hash_count_extra(const EnvHash& h, int lo, int hi, PosFn pos) {
  for (int i = lo; i < hi; ++i) h.aslot[i] = hash_insert(h.keys, h.cnt, h.mask, h.log2hs, pos(i));
  __syncthreads();
  int extra = 0;
  for (int i = lo; i < hi; ++i) extra += (int)__ldcg(&h.cnt[h.aslot[i]]) - 1;
  return extra;
}
__device__ __forceinline__ int hash_c(const EnvHash& h, int i) { return (int)__ldcg(&h.cnt[h.aslot[i]]); }
// phase 3: wipe the used slots (call after every thread finished reading counts)
__device__ __forceinline__ void hash_clear(const EnvHash& h, int lo, int hi) {
  __syncthreads();
  for (int i = lo; i < hi; ++i) {
    const unsigned s = h.aslot[i];
    h.keys[s] = HASH_EMPTY;
    h.cnt[s] = 0u;
  }
  __syncthreads();
}

__device__ __forceinline__ EnvHash env_hash(const StagedScratch& sc, int N) {
  EnvHash h;
  const unsigned hs = 1u << sc.log2hs;
  h.keys = sc.hash + (size_t)blockIdx.x * 2u * hs;
  h.cnt = h.keys + hs;
  h.aslot = sc.aslot + (size_t)blockIdx.x * (size_t)N;
  h.mask = hs - 1u;
  h.log2hs = sc.log2hs;
  return h;
}

// ------------------------------------------------------------------------------------------------
// K2: collision re-draw rounds for large N.  grid-stride over envs, one CTA per env at a time.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) collide_kernel(const StepParams p, const StagedScratch sc) {
  __shared__ BlockScratch bs;
  const int N = p.N, S = p.S;
  const EnvHash h = env_hash(sc, N);
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
  for (int env = blockIdx.x; env < p.E; env += gridDim.x) {
    if (threadIdx.x < 12) sc.envacc[(size_t)env * 12 + threadIdx.x] = 0;   // per-env sums of this step
    if (p.frozen[env]) continue;              // uniform per CTA
    const size_t base = (size_t)env * (size_t)N;
    int16_t* __restrict__ nx = sc.nx + base;
    int16_t* __restrict__ ny = sc.ny + base;
    uint8_t* __restrict__ eff = sc.eff + base;
    const uint8_t* __restrict__ slab = p.slab ? p.slab + (size_t)env * (size_t)p.K : nullptr;
    int cur = 0;
    unsigned err = 0;
    for (;;) {
      const int extra = hash_count_extra(h, lo, hi, [&](int i) { return pack2(nx[i], ny[i]); });
      int L;
      int off = block_excl_scan(extra, L, bs);
      const bool overflow = cur + L > p.K;
      if (L > 0 && !overflow) {
        for (int i = lo; i < hi; ++i) {
          const int e = hash_c(h, i) - 1;
          if (e > 0) {                                       // SDE:250-255
            const int a = tape_slab(p, slab, env, cur + off + e - 1);
            eff[i] = (uint8_t)a;
            nx[i] = (int16_t)wrap1((int)p.x[base + i] + move_dx(a), S);
            ny[i] = (int16_t)wrap1((int)p.y[base + i] + move_dy(a), S);
          }
          off += e;
        }
      }
      hash_clear(h, lo, hi);
      if (L == 0) break;
      if (overflow) { err |= SWARM_ERR_SLAB_OVERFLOW; break; }
      cur += L;
    }
    if (threadIdx.x == 0) {
      p.consumed[env] = cur;
      if (err) p.err[env] |= err;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2b: collision re-draw rounds with the env's occupancy bitmap in SHARED memory (S*S bits: 128 KB at
// S = 1024) -- the large-N analogue of bitmap_counts in the fused kernel.  Every agent sets its cell's bit
// (atomicOr); an agent that finds the bit set is a "hit" and a cell holding c agents yields exactly c-1 hits,
// so c_i = 1 + #{hits on agent i's cell}.  Hits are counted per cell in a small shared hash table
// (COLL_HC slots), every agent then looks its cell up once.  If a round has more hits than the table can
// take (very dense grids) that round falls back to the global hash of collide_kernel.
// Requires ceil(N / blockDim) <= 16 (c_i - 1 is kept in 4 bits per agent; c_i <= 9 because the old cells
// are distinct and a move is at most one cell).
// ------------------------------------------------------------------------------------------------
constexpr int COLL_HC = 4096;   // shared hash slots (power of two)
constexpr int COLL_HL = 8;      // a round with at most this many hits keeps them in a list instead of the hash

__global__ void __launch_bounds__(1024) collide_bitmap_kernel(const StepParams p, const StagedScratch sc, int bm_words,
                                                              const BoxPlan bp) {
  extern __shared__ __align__(16) unsigned coll_smem[];
  __shared__ BlockScratch bs;
  __shared__ unsigned hit_list[COLL_HL];
  __shared__ int hit_n;
  unsigned* sbit = coll_smem;                    // [bm_words]
  unsigned* hkeys = coll_smem + bm_words;        // [COLL_HC]
  unsigned* hcnt = hkeys + COLL_HC;              // [COLL_HC]
  int* hx = reinterpret_cast<int*>(hcnt + COLL_HC);   // [S] (box plan with histograms only)
  int* hy = hx + p.S;                            // [S]
  const int N = p.N, S = p.S;
  const EnvHash gh = env_hash(sc, N);
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
  for (int w = threadIdx.x; w < bm_words; w += blockDim.x) sbit[w] = 0u;
  for (int w = threadIdx.x; w < COLL_HC; w += blockDim.x) { hkeys[w] = HASH_EMPTY; hcnt[w] = 0u; }
  if (threadIdx.x == 0) hit_n = 0;
  __syncthreads();
  for (int env = blockIdx.x; env < p.E; env += gridDim.x) {
    if (threadIdx.x < 12) sc.envacc[(size_t)env * 12 + threadIdx.x] = 0;   // per-env sums of this step
    if (p.frozen[env]) continue;              // uniform per CTA
    const size_t base = (size_t)env * (size_t)N;
    int16_t* __restrict__ nx = sc.nx + base;
    int16_t* __restrict__ ny = sc.ny + base;
    uint8_t* __restrict__ eff = sc.eff + base;
    const uint8_t* __restrict__ slab = p.slab ? p.slab + (size_t)env * (size_t)p.K : nullptr;
    int cur = 0;
    unsigned err = 0;
    // the thread's (<= 16) candidate positions live in registers for all rounds: re-reading them from global in every
    // phase (scalar int16 loads at a 32 B stride) made the L1 sector rate, not the shared-memory work, the limit
    unsigned pos[16];
    const int nown = hi - lo;
    if (nown == 16 && ((((uintptr_t)(nx + lo)) | ((uintptr_t)(ny + lo))) & 15) == 0) {      // 2 x LDG.128 per axis
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const uint4 xv = reinterpret_cast<const uint4*>(nx + lo)[v], yv = reinterpret_cast<const uint4*>(ny + lo)[v];
        const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          pos[8 * v + 2 * h] = __byte_perm(xs[h], ys[h], 0x5410);
          pos[8 * v + 2 * h + 1] = __byte_perm(xs[h], ys[h], 0x7632);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) pos[k] = k < nown ? pack2(nx[lo + k], ny[lo + k]) : 0u;
    }
    // INCREMENTAL rounds: the bitmap is kept across the rounds of an env.  Round 1 inserts every agent; a re-drawn agent
    // takes its old bit out (every occupant of a shared cell re-draws, so the cell empties) and only the agents that
    // moved are inserted in the next round.  After round 1 a round has a handful of hits: they go to an 8-entry list
    // that every thread compares its cells with (c_i - 1 = number of hits on agent i's cell, exactly as with the hash),
    // instead of 16 hash probes per thread.  ~4,500 -> ~1,700 warp instructions per env at cfg4.
    unsigned moved = nown >= 16 ? 0xFFFFu : ((1u << nown) - 1u);
    for (;;) {
      // phase A: bitmap insert of the agents that moved, hit flags
      unsigned hitmask = 0u;
      if (moved) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (moved & (1u << k)) {
            const int cell = (int)(pos[k] >> 16) * S + (int)(pos[k] & 0xFFFFu);
            const unsigned bit = 1u << (cell & 31);
            if (atomicOr(&sbit[cell >> 5], bit) & bit) {
              hitmask |= 1u << k;
              const int at = atomicAdd(&hit_n, 1);
              if (at < COLL_HL) hit_list[at] = pos[k];
            }
          }
        }
      }
      const int H = block_sum(__popc(hitmask), bs);     // (syncs: all bits are set, all hits are listed)
      if (H == 0 && bp.enabled == 1) {
        // no shared cell: the bitmap is the exact occupancy of the final positions -> reward counts (SDE:330-343)
        if (bp.need_hist) {
          for (int a = threadIdx.x; a < 2 * S; a += blockDim.x) hx[a] = 0;
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < nown) { atomicAdd(&hx[pos[k] & 0xFFFFu], 1); atomicAdd(&hy[pos[k] >> 16], 1); }
          __syncthreads();
          // inclusive prefix sums over the S columns / S rows (thread t owns a contiguous chunk of each)
          const int hper = (S + blockDim.x - 1) / blockDim.x;
          const int h0 = min(S, (int)threadIdx.x * hper), h1 = min(S, h0 + hper);
          for (int q = 0; q < 2; ++q) {
            int* h = q ? hy : hx;
            int run = 0;
            for (int a = h0; a < h1; ++a) run += h[a];
            int tot;
            int carry = block_excl_scan(run, tot, bs);
            for (int a = h0; a < h1; ++a) { carry += h[a]; h[a] = carry; }
          }
          __syncthreads();
        }
        int ax = 0, ay = 0, dx = 0, dy = 0;
#pragma unroll 1
        for (int k = 0; k < nown; ++k) {
          unsigned pk = 0u;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q == k) pk = pos[q];
          const int i = lo + k;
          const int xi = (int)(pk & 0xFFFFu), yi = (int)(pk >> 16);
          int cnt[4] = {0, 0, 0, 0};
          if (bp.fast_rmax >= 0) box_direct_sweep(bp, sbit, xi, yi, S, cnt);     // all direct boxes in one row sweep
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (bp.fast_rmax < 0 || bp.mode[q] != BOX_DIRECT) cnt[q] = box_count(bp, q, sbit, hx, hy, xi, yi, S, N);
          const uint4 c4 = make_uint4((unsigned)cnt[0], (unsigned)cnt[1], (unsigned)cnt[2], (unsigned)cnt[3]);
          reinterpret_cast<uint4*>(sc.pcounts + (base + i) * 8)[0] = c4;
          const int a = eff[i];
          if (a < 8) {                                   // summed axis / diagonal move vectors (SDE:346-347)
            const int mx = move_dx(a), my = move_dy(a);
            if (a & 1) { dx += mx; dy += my; } else { ax += mx; ay += my; }
          }
        }
        ax = block_sum(ax, bs); ay = block_sum(ay, bs); dx = block_sum(dx, bs); dy = block_sum(dy, bs);
        if (threadIdx.x == 0) {
          int* ea = sc.envacc + (size_t)env * 12;
          ea[0] = ax; ea[1] = ay; ea[2] = dx; ea[3] = dy;
        }
      }
      if (H == 0) break;                                                            // change_position_id -> None
      unsigned long long cpack = 0ull;                  // (c_i - 1) in 4 bits per owned agent
      int extra = 0;
      const bool listed = H <= COLL_HL;
      const bool small = !listed && H <= COLL_HC / 2;
      if (listed) {
        unsigned hk[COLL_HL];
#pragma unroll
        for (int q = 0; q < COLL_HL; ++q) hk[q] = q < H ? hit_list[q] : 0xFFFFFFFFu;   // (no cell has y = 0xFFFF)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < nown) {
            unsigned cnt = 0u;
#pragma unroll
            for (int q = 0; q < COLL_HL; ++q) cnt += (hk[q] == pos[k]);
            cpack |= (unsigned long long)cnt << (4 * k);
            extra += (int)cnt;
          }
        }
      } else if (small) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (hitmask & (1u << k)) hash_insert(hkeys, hcnt, COLL_HC - 1, 12, pos[k]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < nown) {
            const unsigned key = pos[k];
            unsigned sl = hash_slot(key, 12);
            unsigned cnt = 0u;
            for (;;) {
              const unsigned kk = hkeys[sl];
              if (kk == key) { cnt = hcnt[sl]; break; }
              if (kk == HASH_EMPTY) break;
              sl = (sl + 1u) & (COLL_HC - 1);
            }
            cnt = min(cnt, 15u);
            cpack |= (unsigned long long)cnt << (4 * k);
            extra += (int)cnt;
          }
        }
      } else {
        extra = hash_count_extra(gh, lo, hi, [&](int i) { return pack2(nx[i], ny[i]); });
        for (int i = lo; i < hi; ++i) cpack |= (unsigned long long)min(hash_c(gh, i) - 1, 15) << (4 * (i - lo));
      }
      int L;
      int off = block_excl_scan(extra, L, bs);           // (syncs: all lookups are done)
      const bool overflow = cur + L > p.K;
      moved = 0u;
      if (!overflow && cpack != 0ull) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int e = (int)((cpack >> (4 * k)) & 15ull);
          if (e > 0) {                                       // SDE:250-255
            const int i = lo + k;
            const int oc = (int)(pos[k] >> 16) * S + (int)(pos[k] & 0xFFFFu);
            atomicAnd(&sbit[oc >> 5], ~(1u << (oc & 31)));   // every occupant of this shared cell leaves it
            const int a = tape_slab(p, slab, env, cur + off + e - 1);
            const int vx = wrap1((int)p.x[base + i] + move_dx(a), S), vy = wrap1((int)p.y[base + i] + move_dy(a), S);
            eff[i] = (uint8_t)a;
            nx[i] = (int16_t)vx;
            ny[i] = (int16_t)vy;
            pos[k] = pack2(vx, vy);
            moved |= 1u << k;
          }
          off += e;
        }
      }
      if (threadIdx.x == 0) hit_n = 0;
      if (small) {
        for (int w = threadIdx.x; w < COLL_HC; w += blockDim.x) { hkeys[w] = HASH_EMPTY; hcnt[w] = 0u; }
        __syncthreads();
      } else if (!listed) {
        hash_clear(gh, lo, hi);
      } else {
        __syncthreads();                                   // old bits are out and the list is empty before the next inserts
      }
      if (overflow) { err |= SWARM_ERR_SLAB_OVERFLOW; break; }
      cur += L;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)                          // leave the bitmap clean for the next env
      if (k < nown) sbit[((int)(pos[k] >> 16) * S + (int)(pos[k] & 0xFFFFu)) >> 5] = 0u;
    __syncthreads();
    if (err && bp.enabled && threadIdx.x == 0) sc.envacc[(size_t)env * 12 + 9] = 1;   // boxcount_kernel skips this env
    if (err && bp.enabled) {
      // Slab exhausted: cells are still shared, which a bitmap cannot represent.  The reference semantics go on
      // with those positions, so this (exceptional, flagged) env gets its counts from a plain all-pairs loop.
      __syncthreads();
      int ax = 0, ay = 0, dx = 0, dy = 0;
      for (int i = lo; i < hi; ++i) {
        const int xi = nx[i], yi = ny[i];
        uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
        for (int j = 0; j < N; ++j) {
          const int d = max(iabs(xi - (int)nx[j]), iabs(yi - (int)ny[j]));
          c4.x += d < p.t_rep; c4.y += d < p.t_att; c4.z += d < p.t_repf; c4.w += d < p.t_attf;
        }
        reinterpret_cast<uint4*>(sc.pcounts + (base + i) * 8)[0] = c4;
        const int a = eff[i];
        if (a < 8) {
          const int mx = move_dx(a), my = move_dy(a);
          if (a & 1) { dx += mx; dy += my; } else { ax += mx; ay += my; }
        }
      }
      ax = block_sum(ax, bs); ay = block_sum(ay, bs); dx = block_sum(dx, bs); dy = block_sum(dy, bs);
      if (threadIdx.x == 0) {
        int* ea = sc.envacc + (size_t)env * 12;
        ea[0] = ax; ea[1] = ay; ea[2] = dx; ea[3] = dy;
      }
    }
    if (threadIdx.x == 0) {
      p.consumed[env] = cur;
      if (err) p.err[env] |= err;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2c: box counts as their own kernel when there are fewer envs than SMs (cfg4: 16 envs): `parts` CTAs per env
// each rebuild the env's occupancy bitmap (+ histograms) in their shared memory -- cheap next to the counting --
// and count 1/parts of the agents, so the reward phase runs on parts x E SMs instead of E.
// Envs whose re-draw slab overflowed (flag at envacc[9]) were counted by collide_bitmap_kernel's all-pairs loop.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) boxcount_kernel(const StepParams p, const StagedScratch sc, int bm_words,
                                                        const BoxPlan bp, int parts) {
  extern __shared__ __align__(16) unsigned coll_smem[];
  __shared__ BlockScratch bs;
  unsigned* sbit = coll_smem;                                  // [bm_words]
  int* hx = reinterpret_cast<int*>(coll_smem + bm_words);      // [S]
  int* hy = hx + p.S;                                          // [S]
  const int N = p.N, S = p.S;
  const int env = blockIdx.x / parts, part = blockIdx.x % parts;
  if (p.frozen[env] || sc.envacc[(size_t)env * 12 + 9]) return;
  const size_t base = (size_t)env * (size_t)N;
  const int16_t* __restrict__ nx = sc.nx + base;
  const int16_t* __restrict__ ny = sc.ny + base;
  for (int w = threadIdx.x; w < (bm_words >> 2); w += blockDim.x) reinterpret_cast<uint4*>(sbit)[w] = make_uint4(0u, 0u, 0u, 0u);
  for (int w = (bm_words & ~3) + threadIdx.x; w < bm_words; w += blockDim.x) sbit[w] = 0u;
  if (bp.need_hist)
    for (int a = threadIdx.x; a < 2 * S; a += blockDim.x) hx[a] = 0;
  __syncthreads();
  if ((N & 7) == 0 && ((((uintptr_t)nx) | ((uintptr_t)ny)) & 15) == 0) {           // 8 agents per 128-bit load
    for (int v = threadIdx.x; v < (N >> 3); v += blockDim.x) {
      const uint4 xv = reinterpret_cast<const uint4*>(nx)[v], yv = reinterpret_cast<const uint4*>(ny)[v];
      const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int xi = (int)((xs[e >> 1] >> (16 * (e & 1))) & 0xFFFFu), yi = (int)((ys[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
        const int cell = yi * S + xi;
        atomicOr(&sbit[cell >> 5], 1u << (cell & 31));
        if (bp.need_hist) { atomicAdd(&hx[xi], 1); atomicAdd(&hy[yi], 1); }
      }
    }
  } else {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const int xi = nx[i], yi = ny[i];
      const int cell = yi * S + xi;
      atomicOr(&sbit[cell >> 5], 1u << (cell & 31));
      if (bp.need_hist) { atomicAdd(&hx[xi], 1); atomicAdd(&hy[yi], 1); }
    }
  }
  __syncthreads();
  if (bp.need_hist) {
    const int hper = (S + blockDim.x - 1) / blockDim.x;
    const int h0 = min(S, (int)threadIdx.x * hper), h1 = min(S, h0 + hper);
    for (int q = 0; q < 2; ++q) {
      int* h = q ? hy : hx;
      int run = 0;
      for (int a = h0; a < h1; ++a) run += h[a];
      int tot;
      int carry = block_excl_scan(run, tot, bs);
      for (int a = h0; a < h1; ++a) { carry += h[a]; h[a] = carry; }
    }
    __syncthreads();
  }
  const int begin = (int)(((long long)N * part) / parts), end = (int)(((long long)N * (part + 1)) / parts);
  int ax = 0, ay = 0, dx = 0, dy = 0;
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int xi = nx[i], yi = ny[i];
    int cnt[4] = {0, 0, 0, 0};
    if (bp.fast_rmax >= 0) box_direct_sweep(bp, sbit, xi, yi, S, cnt);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (bp.fast_rmax < 0 || bp.mode[q] != BOX_DIRECT) cnt[q] = box_count(bp, q, sbit, hx, hy, xi, yi, S, N);
    reinterpret_cast<uint4*>(sc.pcounts + (base + i) * 8)[0] =
        make_uint4((unsigned)cnt[0], (unsigned)cnt[1], (unsigned)cnt[2], (unsigned)cnt[3]);
    const int a = sc.eff[base + i];
    if (a < 8) {                                     // summed axis / diagonal move vectors (SDE:346-347)
      const int mx = move_dx(a), my = move_dy(a);
      if (a & 1) { dx += mx; dy += my; } else { ax += mx; ay += my; }
    }
  }
  ax = block_sum(ax, bs); ay = block_sum(ay, bs); dx = block_sum(dx, bs); dy = block_sum(dy, bs);
  if (threadIdx.x == 0) {
    int* ea = sc.envacc + (size_t)env * 12;
    if (ax) atomicAdd(ea + 0, ax);
    if (ay) atomicAdd(ea + 1, ay);
    if (dx) atomicAdd(ea + 2, dx);
    if (dy) atomicAdd(ea + 3, dy);
  }
}

// ------------------------------------------------------------------------------------------------
// K3: predator, one warp per env (any N).  Reads the OLD positions only when the env retargets.
// Also the standalone swarm_predator kernel (then frozen == nullptr, outputs subset).
// ------------------------------------------------------------------------------------------------
struct PredIO {
  int E, N, S;
  const int16_t *xo, *yo, *xn, *yn;
  const uint8_t *retarget, *frozen;   // retarget == NULL: Philox stream 1 (rng_* below)
  unsigned rng_seed_lo, rng_seed_hi, rng_step, env_offset;
  int16_t *px, *py;
  int32_t* target;
  int16_t* tcell;                     // [E,2] in/out cell of the target (swarm_state.tcell); NULL: read xo/yo[target]
  uint8_t* orient;
  int16_t *pred_prev, *pred_next;
  uint8_t* step_orient;
  int32_t* step_target;
  uint8_t* done;
  int32_t* eaten;
  uint32_t* err;
};

__device__ __forceinline__ unsigned warp_umin(unsigned v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
  return v;
}

__device__ __forceinline__ int pred_retarget(const PredIO& io, int env) {
  if (io.retarget) return io.retarget[env];
  return philox_retarget(io.rng_seed_lo, io.rng_seed_hi, io.rng_step, env + (int)io.env_offset);
}

// 8 packed agents (one uint4 of x, one of y) against the predator cell: nearest-neighbour keys
__device__ __forceinline__ void nn_keys8(const uint4& xv, const uint4& yv, int px, int py, int i0, unsigned& klo,
                                         unsigned& khi) {
  const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ax = (int)(int16_t)((xs[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
    const int ay = (int)(int16_t)((ys[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
    const int d = max(iabs(px - ax), iabs(py - ay));
    klo = min(klo, nn_key_lo(d, i0 + e));
    khi = min(khi, nn_key_hi(d, i0 + e));
  }
}
// finished episode, auto-reset off (SDE:233-234): nothing moves, the step's predator outputs are the standing state
__device__ __forceinline__ void predator_frozen(const PredIO& io, int env) {
  io.err[env] |= SWARM_ERR_STEP_WHEN_DONE;
  io.done[env] = 1;
  io.eaten[env] = -1;
  const unsigned pp = pack2(io.px[env], io.py[env]);
  if (io.pred_prev) reinterpret_cast<unsigned*>(io.pred_prev)[env] = pp;
  if (io.pred_next) reinterpret_cast<unsigned*>(io.pred_next)[env] = pp;
  if (io.step_orient) io.step_orient[env] = io.orient[env];
  if (io.step_target) io.step_target[env] = io.target[env];
}

// element e (0..7) of a packed vector pair as a cell word (x | y << 16)
__device__ __forceinline__ unsigned cell8(const uint4& xv, const uint4& yv, int e) {
  const int h = e >> 1;
  const unsigned wx = h == 0 ? xv.x : (h == 1 ? xv.y : (h == 2 ? xv.z : xv.w));
  const unsigned wy = h == 0 ? yv.x : (h == 1 ? yv.y : (h == 2 ? yv.z : yv.w));
  const int sh = 16 * (e & 1);
  return ((wx >> sh) & 0xFFFFu) | (((wy >> sh) & 0xFFFFu) << 16);
}
__device__ __forceinline__ void predator_write(const PredIO& io, int env, int ppx, int ppy, int px, int py, int target,
                                               int orient, unsigned hit) {
  const bool done = hit != 0xFFFFFFFFu;
  io.px[env] = (int16_t)px;
  io.py[env] = (int16_t)py;
  io.target[env] = target;
  io.orient[env] = (uint8_t)orient;
  io.done[env] = done ? 1 : 0;
  io.eaten[env] = done ? (int)hit : -1;
  if (io.pred_prev) reinterpret_cast<unsigned*>(io.pred_prev)[env] = pack2(ppx, ppy);
  if (io.pred_next) reinterpret_cast<unsigned*>(io.pred_next)[env] = pack2(px, py);
  if (io.step_orient) io.step_orient[env] = (uint8_t)orient;
  if (io.step_target) io.step_target[env] = target;
}

// Lane-group variant: LPE lanes per env (32/LPE envs per warp), 8 agents per vector, up to VPL vectors per
// lane prefetched into registers (VPL = 1 / 2 for envs of 8 / 16 agents: a thread per env with no padding vectors in
// registers, so that more envs are in flight per SM).  N % 8 == 0.  Round trip 1 is the predator's state (tiny, coalesced over envs);
// round trip 2 issues the bulk together: NEW positions, the stored target's OLD cell, and -- for the envs that
// retarget -- the OLD positions.
#ifndef PRED_VPL_DEFAULT
#define PRED_VPL_DEFAULT 4
#endif
constexpr int PRED_VPL = PRED_VPL_DEFAULT;

#ifndef PRED_THREADS
#define PRED_THREADS 64    // measured at 1 M x 128: 256 threads 0.159 ms, 128 0.153, 64 0.147 (finer CTAs fill the SMs more evenly)
#endif
template <int LPE, int VPL = PRED_VPL>
__global__ void __launch_bounds__(PRED_THREADS) predator_group_kernel(const PredIO io) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int env = gtid / LPE;
  const int li = threadIdx.x % LPE;
  const unsigned lane = threadIdx.x & 31;
  const unsigned gmask = (LPE == 32) ? 0xFFFFFFFFu : (((1u << LPE) - 1u) << ((lane / LPE) * LPE));
  if (env >= io.E) return;
  const int N = io.N, S = io.S;
  const int NV = N >> 3;
  if (io.frozen && io.frozen[env]) {
    if (li == 0) predator_frozen(io, env);
    return;
  }
  const size_t vbase = ((size_t)env * (size_t)N) >> 3;
  const uint4* __restrict__ xn4 = reinterpret_cast<const uint4*>(io.xn) + vbase;
  const uint4* __restrict__ yn4 = reinterpret_cast<const uint4*>(io.yn) + vbase;
  const uint4* __restrict__ xo4 = reinterpret_cast<const uint4*>(io.xo) + vbase;
  const uint4* __restrict__ yo4 = reinterpret_cast<const uint4*>(io.yo) + vbase;
  // ONE round trip for everything that does not depend on the retarget draw: the predator's state, the cached target
  // cell (4 B, read even by the ~10 % of envs that will retarget: cheaper than a dependent second trip for all the
  // others) and the NEW positions
  int px = io.px[env], py = io.py[env];
  int target = io.target[env];
  const int retarget = pred_retarget(io, env);
  unsigned tc = 0u;
  if (io.tcell) tc = reinterpret_cast<const unsigned*>(io.tcell)[env];
  uint4 xn[VPL], yn[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int v = li + q * LPE;
    if (v < NV) { xn[q] = __ldg(xn4 + v); yn[q] = __ldg(yn4 + v); }
    else { xn[q] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); yn[q] = xn[q]; }
  }
  target = min(max(target, 0), N - 1);
  int tx = (int)(int16_t)(tc & 0xFFFFu), ty = (int)(int16_t)(tc >> 16);
  if (!retarget && !io.tcell) {
    tx = io.xo[(size_t)env * N + target];
    ty = io.yo[(size_t)env * N + target];
  }
  const int ppx = px, ppy = py;
  if (retarget) {                                              // SDE:138-140: second trip, the OLD positions
    unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
    if constexpr (LPE == 1 && VPL <= 2) {                      // the whole env sits in this thread's registers
      uint4 xo[VPL], yo[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) { xo[q] = __ldg(xo4 + q); yo[q] = __ldg(yo4 + q); }
#pragma unroll
      for (int q = 0; q < VPL; ++q) nn_keys8(xo[q], yo[q], px, py, q * 8, klo, khi);
      target = nn_pick(klo, khi, S);
      unsigned cell = cell8(xo[0], yo[0], target & 7);         // the new target's cell, selected from the registers
      if constexpr (VPL == 2) { if (target >= 8) cell = cell8(xo[1], yo[1], target & 7); }
      tx = (int)(int16_t)(cell & 0xFFFFu);
      ty = (int)(int16_t)(cell >> 16);
    } else {
#pragma unroll 4
      for (int v = li; v < NV; v += LPE) nn_keys8(__ldg(xo4 + v), __ldg(yo4 + v), px, py, v * 8, klo, khi);
#pragma unroll
      for (int d = LPE / 2; d > 0; d >>= 1) {
        klo = min(klo, __shfl_xor_sync(gmask, klo, d));
        khi = min(khi, __shfl_xor_sync(gmask, khi, d));
      }
      target = nn_pick(klo, khi, S);
      tx = io.xo[(size_t)env * N + target];                    // L1/L2 hit: the line was just streamed
      ty = io.yo[(size_t)env * N + target];
    }
  }
  int orient;
  pursue(px, py, tx, ty, S, px, py, orient);
  unsigned hit = 0xFFFFFFFFu;                                  // SDE:307-309 on the NEW positions
  const unsigned kx = rep2(px), ky = rep2(py);
#pragma unroll
  for (int q = 0; q < VPL; ++q) hit8(xn[q], yn[q], kx, ky, (li + q * LPE) * 8, hit);   // padding = 0xFFFF: never a cell
#pragma unroll 4
  for (int v = li + VPL * LPE; v < NV; v += LPE) hit8(__ldg(xn4 + v), __ldg(yn4 + v), kx, ky, v * 8, hit);
#pragma unroll
  for (int d = LPE / 2; d > 0; d >>= 1) hit = min(hit, __shfl_xor_sync(gmask, hit, d));
  if (io.tcell && li == 0) {   // the target's NEW cell = what the next step pursues; the line was just streamed (L1 hit)
    const size_t t = (size_t)env * (size_t)N + (size_t)target;   // (selecting it from the prefetched registers instead
    reinterpret_cast<unsigned*>(io.tcell)[env] = pack2(io.xn[t], io.yn[t]);   // pushed the vectors into local memory)
  }
  if (li == 0) predator_write(io, env, ppx, ppy, px, py, target, orient, hit);
}

// CTA-per-env variant for large N (few, big environments): 256 threads stream one env.
__global__ void __launch_bounds__(256) predator_block_kernel(const PredIO io) {
  __shared__ BlockScratch bs;
  const int N = io.N, S = io.S;
  const int NV = N >> 3;
  for (int env = blockIdx.x; env < io.E; env += gridDim.x) {
    if (io.frozen && io.frozen[env]) {                        // uniform per CTA
      if (threadIdx.x == 0) predator_frozen(io, env);
      continue;
    }
    const size_t vbase = ((size_t)env * (size_t)N) >> 3;
    const uint4* __restrict__ xn4 = reinterpret_cast<const uint4*>(io.xn) + vbase;
    const uint4* __restrict__ yn4 = reinterpret_cast<const uint4*>(io.yn) + vbase;
    const uint4* __restrict__ xo4 = reinterpret_cast<const uint4*>(io.xo) + vbase;
    const uint4* __restrict__ yo4 = reinterpret_cast<const uint4*>(io.yo) + vbase;
    int px = io.px[env], py = io.py[env];
    int target = io.target[env];
    unsigned tc = 0u;
    if (io.tcell) tc = reinterpret_cast<const unsigned*>(io.tcell)[env];
    const int ppx = px, ppy = py;
    const int retarget = pred_retarget(io, env);
    if (retarget) {
      unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
#pragma unroll 4
      for (int v = threadIdx.x; v < NV; v += blockDim.x) nn_keys8(__ldg(xo4 + v), __ldg(yo4 + v), px, py, v * 8, klo, khi);
      klo = block_umin(klo, bs);
      khi = block_umin(khi, bs);
      target = nn_pick(klo, khi, S);
    }
    target = min(max(target, 0), N - 1);
    int tx = (int)(int16_t)(tc & 0xFFFFu), ty = (int)(int16_t)(tc >> 16);
    if (retarget || !io.tcell) {                               // (the cached cell spares the dependent read otherwise)
      tx = (int)io.xo[(size_t)env * N + target];
      ty = (int)io.yo[(size_t)env * N + target];
    }
    int orient;
    pursue(px, py, tx, ty, S, px, py, orient);
    unsigned hit = 0xFFFFFFFFu;
    const unsigned kx = rep2(px), ky = rep2(py);
    const int tv = target >> 3;
    if (io.tcell) __syncthreads();          // every thread has read the cached cell before one of them replaces it
#pragma unroll 4
    for (int v = threadIdx.x; v < NV; v += blockDim.x) {
      const uint4 xv = __ldg(xn4 + v), yv = __ldg(yn4 + v);
      hit8(xv, yv, kx, ky, v * 8, hit);
      if (v == tv && io.tcell) reinterpret_cast<unsigned*>(io.tcell)[env] = cell8(xv, yv, target & 7);   // next step's cell
    }
    hit = block_umin(hit, bs);
    if (threadIdx.x == 0) predator_write(io, env, ppx, ppy, px, py, target, orient, hit);
    __syncthreads();   // outputs of this env are written before the state reads of the next iteration
  }
}

// Scalar fallback (N not a multiple of 8): one warp per env.
__global__ void __launch_bounds__(256) predator_kernel(const PredIO io) {
  const int lane = threadIdx.x & 31;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (env >= io.E) return;
  const int N = io.N, S = io.S;
  if (io.frozen && io.frozen[env]) {
    if (lane == 0) predator_frozen(io, env);
    return;
  }
  const size_t base = (size_t)env * (size_t)N;
  int px = io.px[env], py = io.py[env];
  int target = io.target[env];
  const int ppx = px, ppy = py;
  if (pred_retarget(io, env)) {                                      // SDE:138-140
    unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
    for (int i = lane; i < N; i += 32) {
      const int d = max(iabs(px - (int)io.xo[base + i]), iabs(py - (int)io.yo[base + i]));
      klo = min(klo, nn_key_lo(d, i));
      khi = min(khi, nn_key_hi(d, i));
    }
    target = nn_pick(warp_umin(klo), warp_umin(khi), S);
  }
  target = min(max(target, 0), N - 1);
  int orient;
  pursue(px, py, (int)io.xo[base + target], (int)io.yo[base + target], S, px, py, orient);
  unsigned hit = 0xFFFFFFFFu;                                  // SDE:307-309 on the NEW positions
  for (int i = lane; i < N; i += 32)
    if ((int)io.xn[base + i] == px && (int)io.yn[base + i] == py) hit = min(hit, (unsigned)i);
  hit = warp_umin(hit);
  if (lane == 0) {
    if (io.tcell) reinterpret_cast<unsigned*>(io.tcell)[env] = pack2(io.xn[base + target], io.yn[base + target]);
    predator_write(io, env, ppx, ppy, px, py, target, orient, hit);
  }
}

}  // namespace swarm
#include "predator_async.cuh"
namespace swarm {

// host-side dispatch (used by swarm_step's staged path and by swarm_predator)
static inline void launch_predator(const PredIO& io, int num_sms, cudaStream_t s) {
  const int N = io.N;
  const long long E = io.E;
  {                                                            // N <= 256 with the cached target cell: cp.async stream
    cudaError_t ae = cudaSuccess;
    if (predator_async_try(io, num_sms, s, &ae) && ae == cudaSuccess) return;
    if (predator_block_async_try(io, num_sms, s, &ae) && ae == cudaSuccess) return;   // N >= 4096: CTA-per-env stream
  }
  if ((N & 7) != 0) {
    predator_kernel<<<(int)((E * 32 + 255) / 256), 256, 0, s>>>(io);
    return;
  }
  const int nv = N >> 3;
  if (nv >= 512 && E < (long long)num_sms * 64) {              // few big envs: a CTA streams one env
    predator_block_kernel<<<(int)std::min<long long>(E, (long long)num_sms * 8), 256, 0, s>>>(io);
    return;
  }
  int lpe = 1;
  while (lpe < 32 && lpe * PRED_VPL < nv) lpe <<= 1;
  const int grid = (int)((E * lpe + PRED_THREADS - 1) / PRED_THREADS);
  if (nv == 1) { predator_group_kernel<1, 1><<<grid, PRED_THREADS, 0, s>>>(io); return; }
  if (nv == 2) { predator_group_kernel<1, 2><<<grid, PRED_THREADS, 0, s>>>(io); return; }
  switch (lpe) {
    case 1: predator_group_kernel<1><<<grid, PRED_THREADS, 0, s>>>(io); break;
    case 2: predator_group_kernel<2><<<grid, PRED_THREADS, 0, s>>>(io); break;
    case 4: predator_group_kernel<4><<<grid, PRED_THREADS, 0, s>>>(io); break;
    case 8: predator_group_kernel<8><<<grid, PRED_THREADS, 0, s>>>(io); break;
    case 16: predator_group_kernel<16><<<grid, PRED_THREADS, 0, s>>>(io); break;
    default: predator_group_kernel<32><<<grid, PRED_THREADS, 0, s>>>(io); break;
  }
}

// Add v of every lane to acc[env] with one atomic per distinct env in the warp (integer: order independent).
__device__ __forceinline__ void warp_env_add(int env, bool active, int v, int* __restrict__ acc, int stride, int slot) {
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, active ? env : -1);
  const int sum = __reduce_add_sync(peers, v);
  const int lane = threadIdx.x & 31;
  if (active && lane == __ffs(peers) - 1 && sum != 0) atomicAdd(acc + (size_t)env * stride + slot, sum);
}

// ------------------------------------------------------------------------------------------------
// K4: pack positions into pair words {x2, y2, -x2, -y2} (+ VF unit-vector words) and zero the
// per-agent accumulators.  One thread per pair word.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pairprep_kernel(int E, int N, int S, int NW, const int16_t* __restrict__ x,
                                                       const int16_t* __restrict__ y, const uint8_t* __restrict__ eff,
                                                       uint4* __restrict__ pairw, uint2* __restrict__ vfw,
                                                       unsigned* __restrict__ pcounts, unsigned* __restrict__ counter,
                                                       int* __restrict__ envacc) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && counter) *counter = 0u;              // work counter of pairwise_sym_kernel
  const bool live = t < (long long)E * NW;
  const int env = live ? (int)(t / NW) : 0, q = live ? (int)(t % NW) : 0;
  if (envacc) {                                      // summed axis / diagonal move vectors of the env (SDE:346-347)
    int ax = 0, ay = 0, dx = 0, dy = 0;
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * q + h;
        const int a = j < N ? (int)eff[(size_t)env * N + j] : 8;
        if (a < 8) {
          const int mx = move_dx(a), my = move_dy(a);
          if (a & 1) { dx += mx; dy += my; } else { ax += mx; ay += my; }
        }
      }
    }
    warp_env_add(env, live, ax, envacc, 12, 0);
    warp_env_add(env, live, ay, envacc, 12, 1);
    warp_env_add(env, live, dx, envacc, 12, 2);
    warp_env_add(env, live, dy, envacc, 12, 3);
  }
  if (!live) return;
  const size_t base = (size_t)env * (size_t)N;
  const int j0 = 2 * q, j1 = 2 * q + 1;
  const int x0 = x[base + j0], y0 = y[base + j0];
  const int x1 = j1 < N ? (int)x[base + j1] : -S, y1 = j1 < N ? (int)y[base + j1] : 0;
  pairw[t] = make_uint4(pack2(x0, x1), pack2(y0, y1), pack2(-x0, -x1), pack2(-y0, -y1));
  if (vfw) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = j0 + s;
      const int a = j < N ? (int)eff[base + j] : 8;
      const bool mover = a < 8;
      const int mx = mover ? move_dx(a) : 0, my = mover ? move_dy(a) : 0;
      const bool diag = (a & 1) != 0;
      vfw[2 * t + s] = make_uint2(pack2((diag ? 0 : mx) + 1, (diag ? 0 : my) + 1),
                                  pack2((diag ? mx : 0) + 1, (diag ? my : 0) + 1));
    }
  }
  uint4* pc = reinterpret_cast<uint4*>(pcounts + (base + j0) * 8);
  pc[0] = make_uint4(0, 0, 0, 0);
  pc[1] = make_uint4(0, 0, 0, 0);
  if (j1 < N) { pc[2] = make_uint4(0, 0, 0, 0); pc[3] = make_uint4(0, 0, 0, 0); }
}

// ------------------------------------------------------------------------------------------------
// K5: tiled pairwise kernel (the INT32/DPX-bound contract kernel).
//   unit = (env, i-tile of PW_THREADS*PW_APL agents, j-chunk of PW_JC pair words)
//   CTA b sweeps units [b*U/G, (b+1)*U/G): consecutive units share the i-tile, so the accumulators
//   live in registers across j-chunks and are flushed with one atomicAdd per counter when the
//   i-tile changes.  j-chunks are staged by TMA bulk copies (cp.async.bulk, mbarrier completion),
//   double buffered; all threads read the same smem word (broadcast, no bank conflicts).
// ------------------------------------------------------------------------------------------------
constexpr int PW_THREADS = 128;
constexpr int PW_APL = 4;
constexpr int PW_TILE = PW_THREADS * PW_APL;   // 512 agents per i-tile
constexpr int PW_JC = 256;                     // pair words per j-chunk (512 agents, 4 KB)

__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                             unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct PairIO {
  int E, N, S, NW;
  int t_rep, t_att, t_repf, t_attf, t_vf;
  const uint4* pairw;
  const uint2* vfw;
  const int16_t *x, *y;
  const uint8_t* skip;      // [E] (nullable)
  unsigned* pcounts;        // [E,N,8]
  int tiles_per_env, chunks_per_env;
  long long units;
  int one;                  // runtime 1 (see StepParams::one)
};

template <bool HAS_VF>
__global__ void __launch_bounds__(PW_THREADS) pairwise_kernel(const PairIO io) {
  __shared__ __align__(128) uint4 sj[2][PW_JC];
  __shared__ __align__(128) uint4 sv[HAS_VF ? 2 : 1][HAS_VF ? PW_JC : 1];
  __shared__ __align__(8) unsigned long long full_bar[2];
  const int tid = threadIdx.x;
  const long long u_begin = (io.units * blockIdx.x) / gridDim.x;
  const long long u_end = (io.units * (blockIdx.x + 1)) / gridDim.x;
  if (u_begin >= u_end) return;
  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int upe = io.tiles_per_env * io.chunks_per_env;
  auto decode = [&](long long u, int& env, int& tile, int& chunk) {
    env = (int)(u / upe);
    const int r = (int)(u % upe);
    tile = r / io.chunks_per_env;
    chunk = r % io.chunks_per_env;
  };
  auto chunk_words = [&](int chunk) { return min(PW_JC, io.NW - chunk * PW_JC); };
  auto issue = [&](long long u, int stage) {       // thread 0 only
    int env, tile, chunk;
    decode(u, env, tile, chunk);
    const int words = chunk_words(chunk);
    const unsigned bytes = (unsigned)words * 16u * (HAS_VF ? 2u : 1u);
    mbar_expect_tx(&full_bar[stage], bytes);
    tma_bulk_g2s(&sj[stage][0], io.pairw + (size_t)env * io.NW + (size_t)chunk * PW_JC, (unsigned)words * 16u,
                 &full_bar[stage]);
    if (HAS_VF)
      tma_bulk_g2s(&sv[HAS_VF ? stage : 0][0],
                   reinterpret_cast<const uint4*>(io.vfw) + (size_t)env * io.NW + (size_t)chunk * PW_JC,
                   (unsigned)words * 16u, &full_bar[stage]);
  };

  const unsigned TR = rep2(io.t_rep), TA = rep2(io.t_att), TRf = rep2(io.t_repf), TAf = rep2(io.t_attf);
  const unsigned TV = rep2(HAS_VF ? io.t_vf : 0);
  const unsigned one = (unsigned)io.one;

  unsigned XI[PW_APL], YI[PW_APL], NXI[PW_APL], NYI[PW_APL];
  unsigned aR[PW_APL], aA[PW_APL], aRf[PW_APL], aAf[PW_APL], aV[PW_APL], accA[PW_APL], accD[PW_APL];
  int cur_env = -1, cur_tile = -1;
  bool cur_skip = true;

  auto flush = [&]() {
    if (cur_env < 0 || cur_skip) return;
#pragma unroll
    for (int k = 0; k < PW_APL; ++k) {
      const int i = cur_tile * PW_TILE + tid * PW_APL + k;
      if (i < io.N) {
        unsigned* pc = io.pcounts + ((size_t)cur_env * io.N + i) * 8;
        atomicAdd(pc + 0, (unsigned)hsum2(aR[k]));
        atomicAdd(pc + 1, (unsigned)hsum2(aA[k]));
        atomicAdd(pc + 2, (unsigned)hsum2(aRf[k]));
        atomicAdd(pc + 3, (unsigned)hsum2(aAf[k]));
        if (HAS_VF) {
          atomicAdd(pc + 4, (unsigned)hsum2(aV[k]));
          atomicAdd(pc + 5, accA[k]);
          atomicAdd(pc + 6, accD[k]);
        }
      }
    }
  };

  if (tid == 0) issue(u_begin, 0);
  unsigned phase[2] = {0u, 0u};
  for (long long u = u_begin; u < u_end; ++u) {
    const int stage = (int)((u - u_begin) & 1);
    if (tid == 0 && u + 1 < u_end) issue(u + 1, stage ^ 1);   // stage^1 was released by the barrier below
    int env, tile, chunk;
    decode(u, env, tile, chunk);
    if (env != cur_env || tile != cur_tile) {
      flush();
      cur_env = env; cur_tile = tile;
      cur_skip = io.skip && io.skip[env];
      const size_t base = (size_t)env * io.N;
#pragma unroll
      for (int k = 0; k < PW_APL; ++k) {
        const int i = tile * PW_TILE + tid * PW_APL + k;
        const int xi = i < io.N ? (int)io.x[base + i] : -io.S, yi = i < io.N ? (int)io.y[base + i] : 0;
        XI[k] = rep2(xi); YI[k] = rep2(yi); NXI[k] = rep2(-xi); NYI[k] = rep2(-yi);
        aR[k] = aA[k] = aRf[k] = aAf[k] = aV[k] = accA[k] = accD[k] = 0u;
      }
    }
    mbar_wait(&full_bar[stage], phase[stage]);
    phase[stage] ^= 1u;
    if (!cur_skip) {
      const int words = chunk_words(chunk);
#pragma unroll 2
      for (int q = 0; q < words; ++q) {
        const uint4 w = sj[stage][q];
        uint4 uv = make_uint4(0, 0, 0, 0);
        if (HAS_VF) uv = sv[HAS_VF ? stage : 0][q];
#pragma unroll
        for (int k = 0; k < PW_APL; ++k) {
          const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], w);
          aR[k] = lt2(TR, m) * one + aR[k];          // IMAD by a runtime 1: FMA pipe, the ALU pipe keeps the DPX ops
          aA[k] = lt2(TA, m) * one + aA[k];
          aRf[k] = lt2(TRf, m) * one + aRf[k];
          aAf[k] = lt2(TAf, m) * one + aAf[k];
          if (HAS_VF) {
            const unsigned gate = lt2(TV, m);
            aV[k] += gate;
            accA[k] += (gate & 0xFFFFu) * uv.x + (gate >> 16) * uv.z;
            accD[k] += (gate & 0xFFFFu) * uv.y + (gate >> 16) * uv.w;
          }
        }
      }
    }
    __syncthreads();   // everyone is done with sj[stage] before it is refilled two units later
  }
  flush();
}

// ------------------------------------------------------------------------------------------------
// K5s: symmetric tiled pairwise kernel (no visual-field gate).  D_ij = D_ji, so every unordered pair of
// 128-agent tiles (A <= B) of an env is evaluated ONCE by one warp and credited to both sides:
//   lane l owns 4 agents of tile A (rows, in registers); the 32 blocks of tile B are staged in the warp's
//   shared memory ([word][lane] layout: the rotated LDS.128 is conflict free) and met one per round; each
//   indicator goes to the lane's ROW counters and to the visiting block's COLUMN counters, which travel round
//   the ring by shuffles with the block they belong to (after 32 rotations they are home again).  The diagonal
//   unit A == B uses the half ring of the fused kernel.  Rows are flushed once per item, columns once per unit
//   (integer atomicAdd: order independent, results are exact).
// Work items (env, A, chunk of PS_CHUNK B-tiles) are handed out by an atomic counter (zeroed by pairprep).
// Executed instruction mix per word evaluation (2 unordered = 4 ordered pairs): 8 DPX on the ALU pipe,
// 8 IMAD on the FMA pipe -- versus 10 ALU instructions per 2 ordered pairs in the row-only sweep.
// ------------------------------------------------------------------------------------------------
constexpr int PS_WARPS = 4;
constexpr int PS_CHUNK = 8;

struct PairSymIO {
  int E, N, S, NW, T, CH;          // T tiles of 128 agents per env, CH = ceil(T / PS_CHUNK) chunk columns
  int t_rep, t_att, t_repf, t_attf;
  const uint4* pairw;              // [E, NW]
  const uint8_t* skip;             // [E] (nullable)
  unsigned* pcounts;               // [E, N, 8]
  unsigned* counter;               // work counter (zero on entry)
  long long items;                 // E * T * CH (about half of them are empty and skipped)
  int one;
};

__device__ __forceinline__ uint4 pair_word_or_pad(const uint4* __restrict__ pw, int q, int NW, int S) {
  if (q < NW) return __ldg(pw + q);
  return make_uint4(pack2(-S, -S), 0u, pack2(S, S), 0u);
}

__global__ void __launch_bounds__(PS_WARPS * 32, 6) pairwise_sym_kernel(const PairSymIO io) {
  __shared__ __align__(16) uint4 sB[PS_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* sb = sB[warp];
  const unsigned one = (unsigned)io.one, c256 = (unsigned)io.one << 8;
  const unsigned TR = rep2(io.t_rep), TA = rep2(io.t_att), TRf = rep2(io.t_repf), TAf = rep2(io.t_attf);
  const int nxt = (lane + 1) & 31;
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(io.counter, 1u);
    item = __shfl_sync(0xFFFFFFFFu, item, 0);
    if (item >= io.items) break;
    const int per_env = io.T * io.CH;
    const int env = (int)(item / per_env);
    const int idx = (int)(item % per_env);
    const int A = idx / io.CH;
    const int B0 = A + (idx % io.CH) * PS_CHUNK;
    if (B0 >= io.T) continue;
    if (io.skip && io.skip[env]) continue;
    const int B1 = min(io.T, B0 + PS_CHUNK);
    const uint4* __restrict__ pw = io.pairw + (size_t)env * io.NW;
    unsigned* __restrict__ pc = io.pcounts + (size_t)env * io.N * 8;

    unsigned XI[4], YI[4], NXI[4], NYI[4], rowsum[4][4];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const uint4 ow = pair_word_or_pad(pw, A * 64 + 2 * lane + w, io.NW, io.S);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 2 * w + h;
        XI[k] = rep2((int)(int16_t)((ow.x >> (16 * h)) & 0xFFFFu));
        YI[k] = rep2((int)(int16_t)((ow.y >> (16 * h)) & 0xFFFFu));
        NXI[k] = rep2((int)(int16_t)((ow.z >> (16 * h)) & 0xFFFFu));
        NYI[k] = rep2((int)(int16_t)((ow.w >> (16 * h)) & 0xFFFFu));
#pragma unroll
        for (int t = 0; t < 4; ++t) rowsum[k][t] = 0u;
      }
    }

    for (int B = B0; B < B1; ++B) {
      __syncwarp();                                   // previous unit's reads of sb are done
      sb[lane] = pair_word_or_pad(pw, B * 64 + 2 * lane, io.NW, io.S);
      sb[32 + lane] = pair_word_or_pad(pw, B * 64 + 2 * lane + 1, io.NW, io.S);
      __syncwarp();
      // BYTE packed counters (see the fused kernel): [.][0] = {rep, att}, [.][1] = {rep_far, att_far}.
      // Per unit a row byte meets <= 64 pair words and a column byte <= 128 rows: no overflow.
      unsigned row[4][2], col[2][2];
#pragma unroll
      for (int k = 0; k < 4; ++k) row[k][0] = row[k][1] = 0u;
#pragma unroll
      for (int w = 0; w < 2; ++w) col[w][0] = col[w][1] = 0u;

      auto rows_only = [&](int src) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint4 vw = sb[w * 32 + src];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], vw);
            row[k][0] = lt2(TR, m) * one + row[k][0];
            row[k][0] = lt2(TA, m) * c256 + row[k][0];
            row[k][1] = lt2(TRf, m) * one + row[k][1];
            row[k][1] = lt2(TAf, m) * c256 + row[k][1];
          }
        }
      };
      auto rows_cols = [&](int src) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint4 vw = sb[w * 32 + src];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned m = neg_cheb2(XI[k], YI[k], NXI[k], NYI[k], vw);
            const unsigned n0 = lt2(TA, m) * c256 + lt2(TR, m);
            const unsigned n1 = lt2(TAf, m) * c256 + lt2(TRf, m);
            row[k][0] = n0 * one + row[k][0];
            row[k][1] = n1 * one + row[k][1];
            col[w][0] = n0 * one + col[w][0];
            col[w][1] = n1 * one + col[w][1];
          }
        }
      };
      auto rotate = [&](int from) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          col[w][0] = __shfl_sync(0xFFFFFFFFu, col[w][0], from);
          col[w][1] = __shfl_sync(0xFFFFFFFFu, col[w][1], from);
        }
      };

      if (B == A) {                                   // diagonal unit: half ring (as in the fused kernel)
        rows_only(lane);
#pragma unroll 1
        for (int r = 1; r < 16; ++r) {
          if (r > 1) rotate(nxt);
          rows_cols((lane + r) & 31);
        }
        rotate((lane - 15) & 31);                     // lane l held the counters of block l + 15
        rows_only((lane + 16) & 31);
      } else {                                        // off-diagonal unit: full ring, 32 rounds
#pragma unroll 1
        for (int r = 0; r < 32; ++r) {
          if (r > 0) rotate(nxt);
          rows_cols((lane + r) & 31);
        }
        rotate(nxt);                                  // 32nd rotation: home
      }
      // widen the row bytes, flush the column counters of tile B
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rowsum[k][0] += __dp4a(row[k][0], 0x00010001u, 0u);
        rowsum[k][1] += __dp4a(row[k][0], 0x01000100u, 0u);
        rowsum[k][2] += __dp4a(row[k][1], 0x00010001u, 0u);
        rowsum[k][3] += __dp4a(row[k][1], 0x01000100u, 0u);
      }
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = B * 128 + 4 * lane + 2 * w + h;
          if (j < io.N) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const unsigned v = (col[w][t >> 1] >> (16 * h + 8 * (t & 1))) & 0xFFu;
              if (v) atomicAdd(pc + (size_t)j * 8 + t, v);
            }
          }
        }
    }
    // flush the row counters of tile A
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = A * 128 + 4 * lane + k;
      if (i < io.N) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (rowsum[k][t]) atomicAdd(pc + (size_t)i * 8 + t, rowsum[k][t]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6a: finalize, per agent (flat, one thread per agent): per-agent terms from the pairwise accumulators,
// per-agent outputs, commit of the new positions, and the env's integer sums (one atomic per warp and env).
// K6b: finalize, per env (one thread per env): global reward from the integer sums, outputs, statistics.
// ------------------------------------------------------------------------------------------------
template <bool HAS_VF>
__global__ void __launch_bounds__(256) finalize_agents_kernel(const StepParams p, const StagedScratch sc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.E * p.N;
  const bool live = i < total;
  const int N = p.N;
  const int env = live ? (int)(i / N) : 0;
  const int k = live ? (int)(i - (long long)env * N) : 0;
  const bool frozen = live && p.frozen[env] != 0;
  const bool done = live && p.done[env] != 0;
  const bool scored = live && !frozen && !done;
  int rep_i = 0, att_i = 0, P = 0, Q = 0, vis = 0;
  if (scored) {
    const int* __restrict__ ea = sc.envacc + (size_t)env * 12;
    const uint4 c0 = reinterpret_cast<const uint4*>(sc.pcounts + i * 8)[0];
    const int a = sc.eff[i];
    const int selfR = p.t_rep >= 1, selfA = p.t_att >= 1, selfRf = p.t_repf >= 1, selfAf = p.t_attf >= 1;
    const int selfV = HAS_VF ? (p.t_vf >= 1) : 1;
    if (HAS_VF) {
      const uint4 c1 = reinterpret_cast<const uint4*>(sc.pcounts + i * 8)[1];
      const int vin = (int)c1.x;
      vis = vin - selfV;
      align_PQ(a, (int)(c1.y & 0xFFFFu) - vin, (int)(c1.y >> 16) - vin, (int)(c1.z & 0xFFFFu) - vin,
               (int)(c1.z >> 16) - vin, selfV, P, Q);
    } else {
      vis = N - 1;
      align_PQ(a, ea[0], ea[1], ea[2], ea[3], 1, P, Q);
    }
    const AgentTerms t = agent_terms(p, (int)c0.x - selfR, (int)c0.y - selfA, (int)c0.z - selfRf, (int)c0.w - selfAf, vis,
                                     P, Q);
    rep_i = t.rep_i; att_i = t.att_i;
    if (p.agent_reward) {                                   // SDE:371-379
      const double a_rep = p.w_rep * (double)t.rep_i * p.inv_nn;
      const double a_att = p.w_att * (double)t.att_i * p.inv_nn;
      const double a_ali = p.w_ali * (-t.una) * p.inv_nn;
      reinterpret_cast<float4*>(p.agent_reward)[i] =
          make_float4((float)a_att, (float)a_rep, (float)a_ali, (float)(a_rep + a_att + a_ali));
    }
    if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[i] = make_int2(rep_i, att_i);
    if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[i] = pack_terms(rep_i, att_i, vis, P, Q);
  } else if (live && !frozen) {                             // terminated this step (SDE:321-327)
    if (p.agent_reward)
      reinterpret_cast<float4*>(p.agent_reward)[i] = make_float4(0.f, 0.f, 0.f, k == p.eaten[env] ? p.eat : 0.f);
    if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[i] = make_int2(0, 0);
    if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[i] = make_uint2(0u, 0u);
  }
  if (frozen) {                                             // SDE:233-234: defined (zero) per-agent outputs
    if (p.agent_reward) reinterpret_cast<float4*>(p.agent_reward)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.agent_counts) reinterpret_cast<int2*>(p.agent_counts)[i] = make_int2(0, 0);
    if (p.agent_terms) reinterpret_cast<uint2*>(p.agent_terms)[i] = make_uint2(0u, 0u);
    if (p.eff_action) p.eff_action[i] = 8;
  }
  if (live && !frozen) {                                    // commit (SDE:260) + terminal copy
    const int16_t vx = sc.nx[i], vy = sc.ny[i];
    p.x[i] = vx;
    p.y[i] = vy;
    if (done && p.term_x) { p.term_x[i] = vx; p.term_y[i] = vy; }
    if (p.eff_action) p.eff_action[i] = sc.eff[i];
  }
  // N <= 16384: |sum P|, |sum Q| <= 2 N^2 = 5.4e8 and sum vis <= N(N-1) = 2.7e8 fit int32
  warp_env_add(env, scored, rep_i, sc.envacc, 12, 4);
  warp_env_add(env, scored, att_i, sc.envacc, 12, 5);
  warp_env_add(env, scored, P, sc.envacc, 12, 6);
  warp_env_add(env, scored, Q, sc.envacc, 12, 7);
  if (HAS_VF) warp_env_add(env, scored, vis, sc.envacc, 12, 8);
}

template <bool HAS_VF>
__global__ void __launch_bounds__(128) finalize_env_kernel(const StepParams p, const StagedScratch sc) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.E) return;
  const int N = p.N;
  if (p.frozen[env]) {                                      // SDE:233-234 (flagged by K3)
    p.counts[2 * env] = 0; p.counts[2 * env + 1] = 0;
    p.consumed[env] = 0;
    for (int q = 0; q < 5; ++q) p.reward[5 * env + q] = 0.f;
    return;
  }
  const bool done = p.done[env] != 0;
  double r_surv = 0.0, r_att = 0.0, r_rep = 0.0, r_ali = 0.0, r_sum = 0.0;
  int g_rep = 0, g_att = 0;
  if (!done) {
    const int* __restrict__ ea = sc.envacc + (size_t)env * 12;
    g_rep = ea[4]; g_att = ea[5];
    const int g_vis = HAS_VF ? ea[8] : N * (N - 1);
    const double g_una = 0.5 * (double)g_vis - 0.25 * (double)ea[6] - 0.35355339059327376220 * (double)ea[7];
    r_rep = p.w_rep * (double)g_rep * p.inv_nn;             // SDE:363-366
    r_att = p.w_att * (double)g_att * p.inv_nn;
    r_ali = p.w_ali * (-g_una) * p.inv_nn;
    r_sum = r_rep + r_att + r_ali;
  } else {
    r_surv = (double)p.eat;
    r_sum = (double)p.eat;
  }
  float* rw = p.reward + 5 * (size_t)env;
  rw[0] = (float)r_surv; rw[1] = (float)r_att; rw[2] = (float)r_rep; rw[3] = (float)r_ali; rw[4] = (float)r_sum;
  p.counts[2 * env] = g_rep;
  p.counts[2 * env + 1] = g_att;
  if (done && !p.auto_reset) p.frozen[env] = 1;
  if (p.track_stats)
    episode_update(p, env, reinterpret_cast<const uint4*>(p.ep_rec)[env], done, (float)r_surv, (float)r_att, (float)r_rep,
                   (float)r_ali);
}

// ------------------------------------------------------------------------------------------------
// K7: reference reset for large N, CTA per env (grid-stride), optionally masked.
//   mode 0: reset every env, zero reset_count / statistics     (swarm_reset)
//   mode 1: reset envs with mask[env] != 0                     (auto-reset after a step, swarm_autoreset)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) reset_kernel(const StepParams p, const StagedScratch sc,
                                                     const uint8_t* __restrict__ mask, int mode) {
  __shared__ BlockScratch bs;
  const int N = p.N, S = p.S;
  const EnvHash h = env_hash(sc, N);
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
  for (int env = blockIdx.x; env < p.E; env += gridDim.x) {
    if (mode == 1 && mask && !mask[env]) continue;
    if (mode == 1 && p.frozen[env]) continue;      // frozen envs report done=1 but are never reset
    const size_t base = (size_t)env * (size_t)N;
    int16_t* __restrict__ x = p.x + base;
    int16_t* __restrict__ y = p.y + base;
    unsigned rc = mode == 0 ? 0u : p.reset_count[env];
    __syncthreads();
    if (threadIdx.x == 0) p.reset_count[env] = rc + 1u;
    const int16_t* __restrict__ tape = place_row(p, env, rc);
    unsigned err = 0;
    for (int i = lo; i < hi; ++i) {                                         // SDE:202-203
      x[i] = (int16_t)tape_place(p, tape, env, rc, i);
      y[i] = (int16_t)tape_place(p, tape, env, rc, N + i);
    }
    int cur = 2 * N;
    for (;;) {
      const int extra = hash_count_extra(h, lo, hi, [&](int i) { return pack2(x[i], y[i]); });
      int L;
      int off = block_excl_scan(extra, L, bs);
      const bool overflow = cur + 2 * L > p.PL;
      if (L > 0 && !overflow) {
        for (int i = lo; i < hi; ++i) {
          const int e = hash_c(h, i) - 1;
          if (e > 0) {                                // SDE:208-209
            const int s = off + e - 1;
            x[i] = (int16_t)tape_place(p, tape, env, rc, cur + s);
            y[i] = (int16_t)tape_place(p, tape, env, rc, cur + L + s);
          }
          off += e;
        }
      }
      hash_clear(h, lo, hi);
      if (L == 0) break;
      if (overflow) { err |= SWARM_ERR_PLACE_OVERFLOW; break; }
      cur += 2 * L;
    }
    const int16_t* __restrict__ pt = ptape_row(p, env, rc);
    int a = 0, px = 0, py = 0;
    for (;;) {                                        // SDE:92-105
      if (a >= p.KP) { err |= SWARM_ERR_PRED_OVERFLOW; break; }
      tape_pred(p, pt, env, rc, a, px, py);
      ++a;
      int hit = 0;
      for (int i = lo; i < hi; ++i) hit |= ((int)x[i] == px && (int)y[i] == py);
      if (!__syncthreads_or(hit)) break;
    }
    unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;    // SDE:109
    for (int i = lo; i < hi; ++i) {
      const int d = max(iabs(px - (int)x[i]), iabs(py - (int)y[i]));
      klo = min(klo, nn_key_lo(d, i));
      khi = min(khi, nn_key_hi(d, i));
    }
    klo = block_umin(klo, bs);
    khi = block_umin(khi, bs);
    if (threadIdx.x == 0) {
      p.px[env] = (int16_t)px;
      p.py[env] = (int16_t)py;
      const int tgt = nn_pick(klo, khi, S);
      p.target[env] = tgt;
      if (p.tcell)     // (another thread's stores, ordered by the block barriers of the reduction above)
        reinterpret_cast<unsigned*>(p.tcell)[env] = pack2(*(volatile int16_t*)&x[tgt], *(volatile int16_t*)&y[tgt]);
      p.orient[env] = 0;
      if (mode == 0) {
        p.frozen[env] = 0;
        p.err[env] = err;
      } else if (err) {
        p.err[env] |= err;
      }
      if (p.track_stats && mode == 0) {
        reinterpret_cast<uint4*>(p.ep_rec)[env] = make_uint4(0u, 0u, 0u, 0u);
        p.fin_count[env] = 0u;
        p.fin_len[env] = 0u;
        reinterpret_cast<float4*>(p.fin_ret)[env] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K7b: reference reset with the env's occupancy bitmap in shared memory (same machinery as collide_bitmap_kernel:
// bitmap hits -> per-cell hit counts in a shared hash -> block scan for the re-draw slots; the thread's <= 16
// positions stay in registers).  The global-hash reset_kernel above took ~0.3 ms per finished 16,384-agent env --
// longer than the whole step.  A round with more hits than the shared hash can take falls back to the global hash.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) reset_bitmap_kernel(const StepParams p, const StagedScratch sc, int bm_words,
                                                            const uint8_t* __restrict__ mask, int mode) {
  extern __shared__ __align__(16) unsigned coll_smem[];
  __shared__ BlockScratch bs;
  unsigned* sbit = coll_smem;                    // [bm_words]
  unsigned* hkeys = coll_smem + bm_words;        // [COLL_HC]
  unsigned* hcnt = hkeys + COLL_HC;              // [COLL_HC]
  const int N = p.N, S = p.S;
  const EnvHash gh = env_hash(sc, N);
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
  const int nown = hi - lo;
  bool smem_ready = false;
  for (int env = blockIdx.x; env < p.E; env += gridDim.x) {
    if (mode == 1 && mask && !mask[env]) continue;
    if (mode == 1 && p.frozen[env]) continue;      // frozen envs report done=1 but are never reset
    if (!smem_ready) {                             // (uniform per CTA) zero the bitmap / hash only when there is work
      for (int w = threadIdx.x; w < bm_words; w += blockDim.x) sbit[w] = 0u;
      for (int w = threadIdx.x; w < COLL_HC; w += blockDim.x) { hkeys[w] = HASH_EMPTY; hcnt[w] = 0u; }
      smem_ready = true;
    }
    const size_t base = (size_t)env * (size_t)N;
    int16_t* __restrict__ x = p.x + base;
    int16_t* __restrict__ y = p.y + base;
    unsigned rc = mode == 0 ? 0u : p.reset_count[env];
    __syncthreads();
    if (threadIdx.x == 0) p.reset_count[env] = rc + 1u;
    const int16_t* __restrict__ tape = place_row(p, env, rc);
    unsigned err = 0;
    unsigned pos[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)                   // SDE:202-203
      pos[k] = k < nown ? pack2(tape_place(p, tape, env, rc, lo + k), tape_place(p, tape, env, rc, N + lo + k)) : 0u;
    int cur = 2 * N;
    for (;;) {
      unsigned hitmask = 0u;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < nown) {
          const int cell = (int)(pos[k] >> 16) * S + (int)(pos[k] & 0xFFFFu);
          const unsigned bit = 1u << (cell & 31);
          if (atomicOr(&sbit[cell >> 5], bit) & bit) hitmask |= 1u << k;
        }
      }
      const int H = block_sum(__popc(hitmask), bs);
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < nown) sbit[((int)(pos[k] >> 16) * S + (int)(pos[k] & 0xFFFFu)) >> 5] = 0u;
      if (H == 0) { __syncthreads(); break; }
      unsigned long long cpack = 0ull;
      int extra = 0;
      int big = 0;                                   // a cell with more than 16 agents: the 4-bit slots cannot hold it
      const bool small = H <= COLL_HC / 2;
      if (small) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (hitmask & (1u << k)) hash_insert(hkeys, hcnt, COLL_HC - 1, 12, pos[k]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < nown) {
            const unsigned key = pos[k];
            unsigned sl = hash_slot(key, 12);
            unsigned cnt = 0u;
            for (;;) {
              const unsigned kk = hkeys[sl];
              if (kk == key) { cnt = hcnt[sl]; break; }
              if (kk == HASH_EMPTY) break;
              sl = (sl + 1u) & (COLL_HC - 1);
            }
            big |= cnt > 15u;
            cnt = min(cnt, 15u);
            cpack |= (unsigned long long)cnt << (4 * k);
            extra += (int)cnt;
          }
        }
      } else {                                       // very dense placement: exact counts from the global hash
        for (int k = 0; k < nown; ++k) {
          unsigned pk = 0u;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q == k) pk = pos[q];
          x[lo + k] = (int16_t)(pk & 0xFFFFu);
          y[lo + k] = (int16_t)(pk >> 16);
        }
        __syncthreads();
        hash_count_extra(gh, lo, hi, [&](int i) { return pack2(x[i], y[i]); });
        extra = 0;
        for (int i = lo; i < hi; ++i) {
          const int e = hash_c(gh, i) - 1;
          big |= e > 15;
          cpack |= (unsigned long long)min(e, 15) << (4 * (i - lo));
          extra += min(e, 15);
        }
      }
      int L;
      int off = block_excl_scan(extra, L, bs);
      // (more than 16 agents drawn onto one cell is treated like an exhausted tape: flagged, never silently wrong)
      const bool overflow = cur + 2 * L > p.PL || __syncthreads_or(big);
      if (!overflow) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int e = (int)((cpack >> (4 * k)) & 15ull);
          if (e > 0) {                                // SDE:208-209: (2,L) block, last duplicate wins
            const int sidx = off + e - 1;
            pos[k] = pack2(tape_place(p, tape, env, rc, cur + sidx), tape_place(p, tape, env, rc, cur + L + sidx));
          }
          off += e;
        }
      }
      if (small) {
        for (int w = threadIdx.x; w < COLL_HC; w += blockDim.x) { hkeys[w] = HASH_EMPTY; hcnt[w] = 0u; }
        __syncthreads();
      } else {
        hash_clear(gh, lo, hi);
      }
      if (overflow) { err |= SWARM_ERR_PLACE_OVERFLOW; break; }
      cur += 2 * L;
    }
    const int16_t* __restrict__ pt = ptape_row(p, env, rc);
    int a = 0, px = 0, py = 0;
    for (;;) {                                        // SDE:92-105
      if (a >= p.KP) { err |= SWARM_ERR_PRED_OVERFLOW; break; }
      tape_pred(p, pt, env, rc, a, px, py);
      ++a;
      const unsigned pk = pack2(px, py);
      int hit = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) hit |= (k < nown) && pos[k] == pk;
      if (!__syncthreads_or(hit)) break;
    }
    unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;    // SDE:109
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < nown) {
        const int xi = (int)(pos[k] & 0xFFFFu), yi = (int)(pos[k] >> 16);
        x[lo + k] = (int16_t)xi;
        y[lo + k] = (int16_t)yi;
        const int d = max(iabs(px - xi), iabs(py - yi));
        klo = min(klo, nn_key_lo(d, lo + k));
        khi = min(khi, nn_key_hi(d, lo + k));
      }
    }
    klo = block_umin(klo, bs);
    khi = block_umin(khi, bs);
    if (threadIdx.x == 0) {
      p.px[env] = (int16_t)px;
      p.py[env] = (int16_t)py;
      const int tgt = nn_pick(klo, khi, S);
      p.target[env] = tgt;
      if (p.tcell)     // (another thread's stores, ordered by the block barriers of the reduction above)
        reinterpret_cast<unsigned*>(p.tcell)[env] = pack2(*(volatile int16_t*)&x[tgt], *(volatile int16_t*)&y[tgt]);
      p.orient[env] = 0;
      if (mode == 0) {
        p.frozen[env] = 0;
        p.err[env] = err;
      } else if (err) {
        p.err[env] |= err;
      }
      if (p.track_stats && mode == 0) {
        reinterpret_cast<uint4*>(p.ep_rec)[env] = make_uint4(0u, 0u, 0u, 0u);
        p.fin_count[env] = 0u;
        p.fin_len[env] = 0u;
        reinterpret_cast<float4*>(p.fin_ret)[env] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// standalone swarm_pairwise epilogue: accumulators -> (rew_rep_i, rew_attr_i)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pair_counts_kernel(const StepParams p, const unsigned* __restrict__ pcounts,
                                                          const uint8_t* __restrict__ skip, int32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.E * p.N) return;
  if (skip && skip[i / p.N]) { reinterpret_cast<int2*>(out)[i] = make_int2(0, 0); return; }
  const uint4 c0 = reinterpret_cast<const uint4*>(pcounts + i * 8)[0];
  const AgentTerms t = agent_terms(p, (int)c0.x - (p.t_rep >= 1), (int)c0.y - (p.t_att >= 1),
                                   (int)c0.z - (p.t_repf >= 1), (int)c0.w - (p.t_attf >= 1), 0, 0, 0);
  reinterpret_cast<int2*>(out)[i] = make_int2(t.rep_i, t.att_i);
}

// ------------------------------------------------------------------------------------------------
// Visual-field observations (README figure "SwarmEnv - 9 Agents - Visual Field 5x5",
// gym_swarm/images/env_illustration.png -- the generating notebook is not in the reference tree, so the cell
// codes below are this repo's definition): for every agent the egocentric (2 vf + 1)^2 window of the grid,
//   0 = empty (or outside the grid: the distance is the plain, non-periodic Chebyshev one, as in SDE:349-355),
//   1 = the agent itself (centre), 2 = another agent, 3 = the predator;  out[e, i, dy + vf, dx + vf].
// One CTA per (env, 128-agent chunk): the env's positions and the chunk's windows live in shared memory, the
// windows leave as one contiguous, coalesced store.  O(N) cell tests per agent, N <= OBS_MAX_AGENTS.
// ------------------------------------------------------------------------------------------------
constexpr int OBS_MAX_AGENTS = 4096;
constexpr int OBS_CHUNK = 128;

// BITMAP variant (S*S bits fit in shared memory): the CTA builds the env's occupancy bitmap once, every agent then
// reads only the 2 vf + 1 rows of its window -- O(vf) instead of O(N) per agent, which makes the kernel HBM bound.
template <bool BITMAP>
__global__ void __launch_bounds__(OBS_CHUNK) observe_vf_kernel(int E, int N, int S, int vf, int bm_words,
                                                               const int16_t* __restrict__ x, const int16_t* __restrict__ y,
                                                               const int16_t* __restrict__ px, const int16_t* __restrict__ py,
                                                               uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char obs_smem[];
  const int W = 2 * vf + 1, WW = W * W;
  int* spos = reinterpret_cast<int*>(obs_smem);                   // [N] packed (x | y << 16)
  unsigned* sbit = reinterpret_cast<unsigned*>(spos + N);         // [bm_words] (BITMAP)
  unsigned char* swin = obs_smem + (((size_t)(N + bm_words) * 4 + 15) & ~(size_t)15);   // [OBS_CHUNK][WW], 16 B aligned
  const int chunks = (N + OBS_CHUNK - 1) / OBS_CHUNK;
  const int env = blockIdx.x / chunks, i0 = (blockIdx.x % chunks) * OBS_CHUNK;
  const size_t base = (size_t)env * (size_t)N;
  if (BITMAP)
    for (int w = threadIdx.x; w < bm_words; w += OBS_CHUNK) sbit[w] = 0u;
  const int nwin = min(OBS_CHUNK, N - i0) * WW;
  {                                                    // (swin is 4-byte aligned: it follows two word arrays)
    unsigned* w32 = reinterpret_cast<unsigned*>(swin);
    for (int q = threadIdx.x; q < (nwin + 3) / 4; q += OBS_CHUNK) w32[q] = 0u;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += OBS_CHUNK) {
    const int xj = x[base + j], yj = y[base + j];
    spos[j] = (int)pack2(xj, yj);
    if (BITMAP) { const int cell = yj * S + xj; atomicOr(&sbit[cell >> 5], 1u << (cell & 31)); }
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i < N) {
    unsigned char* w = swin + (size_t)threadIdx.x * WW;
    const int xi = (int)(int16_t)(spos[i] & 0xFFFF), yi = (int)(int16_t)((unsigned)spos[i] >> 16);
    if (BITMAP) {
      const int x0 = max(xi - vf, 0), x1 = min(xi + vf, S - 1);
      for (int dy = -vf; dy <= vf; ++dy) {
        const int yy = yi + dy;
        if (yy < 0 || yy >= S) continue;
        const int lo = yy * S + x0, hi = yy * S + x1;           // bit range of this window row (<= 15 bits)
        const unsigned a = sbit[lo >> 5], b = sbit[hi >> 5];
        unsigned bits = __funnelshift_r(a, b, lo & 31) & (0xFFFFFFFFu >> (31 - (x1 - x0)));
        unsigned char* row = w + (dy + vf) * W + (x0 - xi + vf);
        while (bits) {
          const int c = __ffs(bits) - 1;
          bits &= bits - 1u;
          row[c] = 2;
        }
      }
      w[vf * W + vf] = 1;
    } else {
      for (int j = 0; j < N; ++j) {
        const int dx = (int)(int16_t)(spos[j] & 0xFFFF) - xi, dy = (int)(int16_t)((unsigned)spos[j] >> 16) - yi;
        if (iabs(dx) <= vf && iabs(dy) <= vf) w[(dy + vf) * W + dx + vf] = 2;
      }
      w[vf * W + vf] = 1;                                       // the centre is the agent itself
    }
    const int dx = (int)px[env] - xi, dy = (int)py[env] - yi;
    if (iabs(dx) <= vf && iabs(dy) <= vf) w[(dy + vf) * W + dx + vf] = 3;
  }
  __syncthreads();
  uint8_t* __restrict__ o = out + (base + i0) * (size_t)WW;
  if ((((uintptr_t)o) & 15) == 0) {                     // the chunk's windows leave as 128-bit stores
    const uint4* s16 = reinterpret_cast<const uint4*>(swin);
    uint4* o16 = reinterpret_cast<uint4*>(o);
    const int n16 = nwin >> 4;
    for (int q = threadIdx.x; q < n16; q += OBS_CHUNK) o16[q] = s16[q];
    for (int q = (n16 << 4) + threadIdx.x; q < nwin; q += OBS_CHUNK) o[q] = swin[q];
  } else {
    for (int q = threadIdx.x; q < nwin; q += OBS_CHUNK) o[q] = swin[q];
  }
}


// Warp-per-env variant for the lane-group shapes (N <= 256, N % APL == 0, a bitmap of a few KB; e.g. cfg3): no
// __syncthreads, no clearing passes.  A warp keeps ONE clean, row-padded occupancy bitmap (VF empty rows above and below the
// grid: no range checks; stored as 50 % overlapping words so that a window row is always inside ONE word) and one staging area in shared memory and walks envs env, env + stride, ... with the next env's
// positions already in flight.  Lane l owns the APL consecutive agents l APL .. l APL + APL - 1, i.e. APL (2 VF + 1)^2
// CONTIGUOUS, 4-byte aligned output bytes: every window row is one bitmap word and a shift, a multiply spreads
// four cells into four code bytes, and the rows are streamed into whole 32-bit words with PRMT (all byte positions are
// compile-time constants once the loops are unrolled) -- no byte stores, no window clear.  The env's N (2 VF + 1)^2 bytes
// leave as coalesced 128-bit stores.  observe_vf_kernel above pays three __syncthreads, a bitmap clear and a window clear
// per 128 agents and sets its cells with byte stores (0.29 of the HBM peak at 262,144 x 128).
constexpr int OBSW_WARPS = 4;

template <int VF, int APL>
__global__ void __launch_bounds__(OBSW_WARPS * 32) observe_vf_warp_kernel(int E, int N, int S, int bm_words, int warp_bytes,
                                                                          const int16_t* __restrict__ x,
                                                                          const int16_t* __restrict__ y,
                                                                          const int16_t* __restrict__ px,
                                                                          const int16_t* __restrict__ py,
                                                                          uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char obs_smem[];
  constexpr int W = 2 * VF + 1, WW = W * W, NJ = (W + 3) / 4;
  constexpr int LW = APL * WW / 4;                               // output words of one lane (APL % 4 == 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned* sbit = reinterpret_cast<unsigned*>(obs_smem + (size_t)warp * warp_bytes);   // [(S + 2 VF) rows], clean between envs
  unsigned* swin = sbit + ((bm_words + 3) & ~3);                                          // [N * WW / 4] words
  for (int w = lane; w < bm_words; w += 32) sbit[w] = 0u;
  __syncwarp();
  const int nbytes = N * WW;
  const bool own = lane * APL < N;
  const int stride = gridDim.x * OBSW_WARPS;
  int env = blockIdx.x * OBSW_WARPS + warp;
  // raw positions of the lane's agents (APL int16 = APL / 2 words per axis) and the predator, one env ahead
  unsigned xr[APL / 2], yr[APL / 2];
  int pxe = 0, pye = 0;
  auto load_env = [&](int e) {
    if (e < E) {
      if (own) {
        const size_t o = ((size_t)e * (size_t)N + (size_t)lane * APL) >> 1;     // in 32-bit words
#pragma unroll
        for (int q = 0; q < APL / 4; ++q) {
          const uint2 a = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned*>(x) + o)[q];
          const uint2 b = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned*>(y) + o)[q];
          xr[2 * q] = a.x; xr[2 * q + 1] = a.y; yr[2 * q] = b.x; yr[2 * q + 1] = b.y;
        }
      }
      pxe = px[e]; pye = py[e];
    }
  };
  load_env(env);
  for (; env < E; env += stride) {
    int xs[APL], ys[APL];
#pragma unroll
    for (int k = 0; k < APL; ++k) {
      xs[k] = (int)(int16_t)((xr[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
      ys[k] = (int)(int16_t)((yr[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
    }
    const int cpx = pxe, cpy = pye;
    load_env(env + stride);                                      // in flight while this env is built
    if (own) {
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const int cell = (ys[k] + VF) * S + xs[k];            // overlapped words: word v holds the bits [16 v - 16, 16 v + 16)
        atomicOr(&sbit[(cell >> 4) + 1], 1u << (cell & 15));
        atomicOr(&sbit[cell >> 4], 1u << ((cell & 15) + 16));
      }
    }
    __syncwarp();
    if (own) {
      unsigned* ow = swin + lane * LW;
      // the lane's byte stream: the `have` (< 4) bytes hold[hs .. hs + have) are waiting for the rest of their word
      unsigned hold = 0u;
      int hs = 0, have = 0, m = 0;                               // compile-time constants once the loops are unrolled
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const int xi = xs[k], yi = ys[k];
        const int x0 = max(xi - VF, 0), x1 = min(xi + VF, S - 1);
        const unsigned cmask = 0xFFFFFFFFu >> (31 - (x1 - x0));
        const int cshift = x0 - (xi - VF);                       // window columns cut off at the left border
        int lo = yi * S + x0;                                    // padded rows: row yi - VF of the grid is row yi here
#pragma unroll
        for (int r = 0; r < W; ++r, lo += S) {
          unsigned bits = ((sbit[(lo >> 4) + 1] >> (lo & 15)) & cmask) << cshift;   // <= 11 + 15 bits: one word
          if (r == VF) bits &= ~(1u << VF);                      // the centre gets code 1 below
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int nb = (W - 4 * j) < 4 ? (W - 4 * j) : 4;    // valid bytes of this source word (the others are 0)
            unsigned e;                                          // four cells -> four bytes of code 2
            if (nb == 1) e = (bits >> (4 * j - 1)) & 2u;         // (j >= 1)
            else e = (((bits >> (4 * j)) & 0xFu) * 0x00408102u) & 0x02020202u;
            if (r == VF && j == VF / 4) e |= 1u << (8 * (VF % 4));
            unsigned sel = 0u;
            if (have + nb >= 4) {                                // a word completes: pending bytes, then e[0 .. 4 - have)
#pragma unroll
              for (int q = 0; q < 4; ++q) sel |= (unsigned)(q < have ? hs + q : 4 + (q - have)) << (4 * q);
              ow[m] = have == 0 ? e : __byte_perm(hold, e, sel);
              ++m;
              hold = e;
              hs = 4 - have;
              have = have + nb - 4;
            } else {                                             // (last word of a row, W % 4 != 0): keep collecting
#pragma unroll
              for (int q = 0; q < 4; ++q)
                sel |= (unsigned)(q < have ? hs + q : (q < have + nb ? 4 + (q - have) : 7)) << (4 * q);   // e[3] == 0 here
              hold = have == 0 ? e : __byte_perm(hold, e, sel);
              hs = 0;
              have = have + nb;
            }
          }
        }
      }
      // predator: code 3 wins over everything (byte stores into the lane's own words, behind its word stores)
#pragma unroll
      for (int k = 0; k < APL; ++k) {
        const int dx = cpx - xs[k], dy = cpy - ys[k];
        if (iabs(dx) <= VF && iabs(dy) <= VF)
          reinterpret_cast<unsigned char*>(ow)[k * WW + (dy + VF) * W + dx + VF] = 3;
      }
    }
    __syncwarp();
    if (own) {
#pragma unroll
      for (int k = 0; k < APL; ++k) {                            // leave the bitmap clean
        const int v = ((ys[k] + VF) * S + xs[k]) >> 4;
        sbit[v] = 0u;
        sbit[v + 1] = 0u;
      }
    }
    uint8_t* __restrict__ o = out + (size_t)env * (size_t)nbytes;
    if ((((uintptr_t)o) & 15) == 0) {
      const uint4* s16 = reinterpret_cast<const uint4*>(swin);
      uint4* o16 = reinterpret_cast<uint4*>(o);
      const int n16 = nbytes >> 4;
      for (int q = lane; q < n16; q += 32) o16[q] = s16[q];
      for (int q = (n16 << 4) + lane; q < nbytes; q += 32) o[q] = reinterpret_cast<const unsigned char*>(swin)[q];
    } else {
      for (int q = lane; q < nbytes; q += 32) o[q] = reinterpret_cast<const unsigned char*>(swin)[q];
    }
    __syncwarp();                                                // staging and bitmap are free for the next env
  }
}

}  // namespace swarm
