// predator_async.cuh -- the predator step for N <= 256 as a persistent, software-pipelined stream (sm_100a).
//
// SDE:111-165 (closest_target / follow_target) and SDE:307-309 (termination test), same arithmetic as the kernels in
// staged.cuh.  What differs is how the bytes travel: predator_group_kernel lands its loads in REGISTERS, so the bytes in
// flight per SM are bounded by the register file times the share of a warp's life spent waiting (ncu: 23 warps x 4 KB,
// half of them computing -> ~46 KB in flight, 0.65 of the HBM peak), and an env that retargets (SDE:138-140) pays a
// second, dependent DRAM round trip.  Here every warp owns a ring of SHARED-memory buffers and fills them with
// cp.async (LDGSTS) W+1 tiles ahead of the tile it computes on:
//
//   state(k)  px, py, target, cached target cell, retarget flags of the warp tile's 32/LPE envs (4-byte pieces)
//   main(k)   the NEW positions of the tile (16-byte pieces, each lane copies exactly the vectors it will read)
//   rows(k)   the OLD positions of the envs whose retarget flag is set -- issued together with main(k), because
//             state(k) travels W+1 tiles further ahead and has landed by then: no dependent round trip
//
// One commit group per iteration; cp.async groups complete in order, so `wait_group W` at the top of iteration `it`
// guarantees main(it), rows(it) and state(it + W + 1).  Nothing is shared between warps and nothing waits on an
// mbarrier: a lane reads another lane's slot only after wait_group + __syncwarp.  Shared memory, not the register file,
// bounds the bytes in flight (~100 KB per SM), and the compute of a tile overlaps the flight of the next W+1.
#pragma once
#include "swarm_common.cuh"

namespace swarm {

#ifndef PA_THREADS
#define PA_THREADS 64
#endif
#ifndef PA_W
#define PA_W 1
#endif

template <int LPE, int V, int W>
struct PredAsyncCfg {
  static constexpr int EW = 32 / LPE;          // envs of one warp tile
  static constexpr int FW = (EW + 3) / 4;      // 4-byte words of the u8-per-env retarget flags
  // state buffer (bytes): px int16[EW] | py int16[EW] | target i32[EW] | target cell u32[EW] | retarget u8[EW]
  static constexpr int O_PX = 0, O_PY = 2 * EW, O_TG = 4 * EW, O_TC = 8 * EW, O_FL = 12 * EW;
  static constexpr int SB = (13 * EW + 15) / 16 * 16;
  static constexpr int SW = 3 * EW + FW;       // the same buffer in 4-byte words (ragged last tile)
  // row slots per warp tile (binomial(EW, 0.1) exceeds them in < 0.5 % of the tiles: those envs read global memory)
  static constexpr int RMAX = EW >= 32 ? 8 : (EW >= 16 ? 5 : (EW >= 8 ? 3 : 2));
  static constexpr int DM = W + 1, DS = 2 * W + 2;              // tiles ahead: main / rows, state
  static constexpr int NMAIN = DM + 1, NROWS = DM + 1, NSTATE = DS + 1;
  static constexpr int MAIN_U4 = 2 * V * 32;                    // [x|y][V][32 lanes]
  static constexpr int ROWS_U4 = RMAX * 2 * V * LPE;            // [slot][x|y][V][LPE]
  static constexpr int WARP_BYTES = (NMAIN * MAIN_U4 + NROWS * ROWS_U4) * 16 + NSTATE * SB;
  static constexpr int CTA_BYTES = WARP_BYTES * (PA_THREADS / 32);
  // pieces of one state(k): 16-byte copies where a segment is a multiple of 16 bytes, 4-byte copies otherwise
  static constexpr int PS_PX = (2 * EW) % 16 == 0 ? 16 : 4, PS_TG = (4 * EW) % 16 == 0 ? 16 : 4, PS_FL = EW % 16 == 0 ? 16 : 4;
  static constexpr int NP_PX = 2 * EW / PS_PX, NP_TG = 4 * EW / PS_TG, NP_FL = EW / PS_FL;
  static constexpr int NPIECES = 2 * NP_PX + 2 * NP_TG + NP_FL;   // <= 32: one cp.async per lane
};

__device__ __forceinline__ void cp_async16(unsigned dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// EXACT: N == 8 * LPE * V (no lane ever holds a padding vector)
template <int LPE, int V, int W, bool EXACT>
__global__ void __launch_bounds__(PA_THREADS) predator_async_kernel(const PredIO io, int tiles) {
  using C = PredAsyncCfg<LPE, V, W>;
  constexpr int EW = C::EW;
  extern __shared__ uint4 pa_smem[];
  const int wic = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPE, li = lane % LPE;
  const unsigned gmask = (LPE == 32) ? 0xFFFFFFFFu : (((1u << LPE) - 1u) << (g * LPE));
  // the warp's ring, as shared-window byte addresses
  const unsigned mainb = (unsigned)__cvta_generic_to_shared(pa_smem) + (unsigned)wic * C::WARP_BYTES;
  const unsigned rowsb = mainb + C::NMAIN * C::MAIN_U4 * 16;
  const unsigned stateb = rowsb + C::NROWS * C::ROWS_U4 * 16;
  const unsigned char* const gen0 = reinterpret_cast<const unsigned char*>(pa_smem) + (size_t)wic * C::WARP_BYTES;
  const unsigned char* const gen_rows = gen0 + C::NMAIN * C::MAIN_U4 * 16;
  const unsigned char* const gen_state = gen_rows + C::NROWS * C::ROWS_U4 * 16;
  const int nwarps = gridDim.x * (PA_THREADS / 32);
  const int w0 = blockIdx.x * (PA_THREADS / 32) + wic;
  const int my = w0 < tiles ? (tiles - w0 + nwarps - 1) / nwarps : 0;   // warp tiles w0, w0 + nwarps, ...
  const int N = io.N, S = io.S, NV = N >> 3, E = io.E;
  const uint4* __restrict__ xn4 = reinterpret_cast<const uint4*>(io.xn);
  const uint4* __restrict__ yn4 = reinterpret_cast<const uint4*>(io.yn);
  const uint4* __restrict__ xo4 = reinterpret_cast<const uint4*>(io.xo);
  const uint4* __restrict__ yo4 = reinterpret_cast<const uint4*>(io.yo);

  // ---- this lane's piece of every state(k): source array, bytes per env, offsets (fixed for the whole launch)
  const unsigned char* role_src = nullptr;
  unsigned role_bpe = 0, role_dst = 0;
  bool role16 = false;
  {
    int j = lane;
    auto seg = [&](const void* base, int bpe, int off_smem, int npieces, int psize) {
      if (j >= 0 && j < npieces) {
        role_src = reinterpret_cast<const unsigned char*>(base) + j * psize;
        role_bpe = (unsigned)bpe;
        role_dst = (unsigned)(off_smem + j * psize);
        role16 = psize == 16;
      }
      j -= npieces;
    };
    seg(io.px, 2, C::O_PX, C::NP_PX, C::PS_PX);
    seg(io.py, 2, C::O_PY, C::NP_PX, C::PS_PX);
    seg(io.target, 4, C::O_TG, C::NP_TG, C::PS_TG);
    seg(io.tcell, 4, C::O_TC, C::NP_TG, C::PS_TG);
    seg(io.retarget, 1, C::O_FL, C::NP_FL, C::PS_FL);
  }

  // ---- state(tile) into state slot `slot`
  auto issue_state = [&](int tile, int slot) {
    const int e0 = tile * EW;
    const unsigned st = stateb + (unsigned)slot * C::SB;
    if (e0 + EW <= E) {                          // whole tile: one piece per lane
      if (role_src) {
        const unsigned char* src = role_src + (size_t)e0 * role_bpe;
        if (role16) cp_async16(st + role_dst, src);
        else cp_async4(st + role_dst, src);
      }
    } else {                                     // ragged last tile (E % 4 == 0): 4-byte words, each inside the arrays
#pragma unroll
      for (int r = 0; r < (C::SW + 31) / 32; ++r) {
        const int j = lane + 32 * r;
        if (j >= C::SW) break;
        const void* src;
        int env;
        if (j < EW / 2) { env = e0 + 2 * j; src = io.px + env; }
        else if (j < EW) { env = e0 + 2 * (j - EW / 2); src = io.py + env; }
        else if (j < 2 * EW) { env = e0 + (j - EW); src = io.target + env; }
        else if (j < 3 * EW) { env = e0 + (j - 2 * EW); src = io.tcell + 2 * (size_t)env; }
        else { env = e0 + 4 * (j - 3 * EW); src = io.retarget + env; }
        if (env < E) cp_async4(st + 4u * (unsigned)j, src);
      }
    }
  };
  // ---- main(tile) + rows(tile); needs the flags of state(tile) in state slot `sslot`
  auto issue_main_rows = [&](int tile, int sslot, int mslot) {
    const int env = tile * EW + g;
    const bool valid = env < E;
    const bool ret = valid && gen_state[sslot * C::SB + C::O_FL + g] != 0;
    const unsigned rb = __ballot_sync(0xFFFFFFFFu, ret && li == 0);
    if (valid) {
      const unsigned voff = (unsigned)env * (unsigned)NV + (unsigned)li;      // vector index: E * N / 8 < 2^32
      const unsigned mb = mainb + (unsigned)mslot * (C::MAIN_U4 * 16) + (unsigned)lane * 16u;
      const uint4* xs = xn4 + voff;
      const uint4* ys = yn4 + voff;
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (EXACT || li + q * LPE < NV) {
          cp_async16(mb + q * 512, xs + q * LPE);
          cp_async16(mb + (V + q) * 512, ys + q * LPE);
        }
      }
      if (ret) {
        const int rank = __popc(rb & ((1u << (g * LPE)) - 1u));
        if (rank < C::RMAX) {
          const unsigned rw = rowsb + (unsigned)mslot * (C::ROWS_U4 * 16) + (unsigned)(rank * (2 * V * LPE) + li) * 16u;
          const uint4* xos = xo4 + voff;
          const uint4* yos = yo4 + voff;
#pragma unroll
          for (int q = 0; q < V; ++q) {
            if (EXACT || li + q * LPE < NV) {
              cp_async16(rw + q * (LPE * 16), xos + q * LPE);
              cp_async16(rw + (V + q) * (LPE * 16), yos + q * LPE);
            }
          }
        }
      }
    }
  };

  // ---- prologue: the first DS states land before the first main / rows are issued
#pragma unroll
  for (int k = 0; k < C::DS; ++k)
    if (k < my) issue_state(w0 + k * nwarps, k % C::NSTATE);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
#pragma unroll
  for (int k = 0; k < C::DM; ++k) {
    if (k < my) issue_main_rows(w0 + k * nwarps, k % C::NSTATE, k % C::NMAIN);
    cp_async_commit();
  }

  int cs = 0, cm = 0;                            // it % NSTATE, it % NMAIN
  int tile = w0;
  for (int it = 0; it < my; ++it, tile += nwarps) {
    cp_async_wait<W>();          // main(it), rows(it), state(it + DM) have landed
    __syncwarp();                // ... for every lane of the warp; and every lane is done with tile it - 1
    {
      const int ss = cs == 0 ? C::NSTATE - 1 : cs - 1;                        // (it + DS) % NSTATE
      const int ms = cs + C::DM >= C::NSTATE ? cs + C::DM - C::NSTATE : cs + C::DM;   // (it + DM) % NSTATE
      const int im = cm == 0 ? C::NMAIN - 1 : cm - 1;                         // (it + DM) % NMAIN
      if (it + C::DS < my) issue_state(tile + C::DS * nwarps, ss);
      if (it + C::DM < my) issue_main_rows(tile + C::DM * nwarps, ms, im);
      cp_async_commit();
    }
    const int env = tile * EW + g;
    const unsigned char* st = gen_state + cs * C::SB;
    const uint4* mb = reinterpret_cast<const uint4*>(gen0) + cm * C::MAIN_U4;
    const uint4* rwb = reinterpret_cast<const uint4*>(gen_rows) + cm * C::ROWS_U4;
    cs = cs + 1 == C::NSTATE ? 0 : cs + 1;
    cm = cm + 1 == C::NMAIN ? 0 : cm + 1;
    const bool valid = env < E;
    const bool ret = valid && st[C::O_FL + g] != 0;
    const unsigned rb = __ballot_sync(0xFFFFFFFFu, ret && li == 0);
    if (!valid) continue;        // whole lane groups leave; the warp meets again at the __syncwarp of the next iteration
    int px = reinterpret_cast<const int16_t*>(st + C::O_PX)[g];
    int py = reinterpret_cast<const int16_t*>(st + C::O_PY)[g];
    int target = min(max(reinterpret_cast<const int*>(st + C::O_TG)[g], 0), N - 1);
    const unsigned tc = reinterpret_cast<const unsigned*>(st + C::O_TC)[g];
    int tx = (int)(int16_t)(tc & 0xFFFFu), ty = (int)(int16_t)(tc >> 16);
    const int ppx = px, ppy = py;
    if (ret) {                                                   // SDE:138-140 on the OLD positions
      const unsigned NP = rep2(-px), P1 = rep2(px + 1), NQ = rep2(-py), Q1 = rep2(py + 1);
      const int rank = __popc(rb & ((1u << (g * LPE)) - 1u));
      const bool staged = rank < C::RMAX;
      const uint4* rw = rwb + (staged ? rank : 0) * (2 * V * LPE);
      const unsigned voff = (unsigned)env * (unsigned)NV + (unsigned)li;
      auto keys = [&](auto hi_tag) -> unsigned {
        constexpr bool HI = decltype(hi_tag)::value;
        unsigned acc = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < V; ++q) {
          if (EXACT || li + q * LPE < NV) {
            const uint4 xv = staged ? rw[q * LPE + li] : __ldg(xo4 + voff + q * LPE);
            const uint4 yv = staged ? rw[(V + q) * LPE + li] : __ldg(yo4 + voff + q * LPE);
            if (q == 0) nn_keys8_packed<0, HI>(xv, yv, NP, P1, NQ, Q1, acc);
            if (q == 1) nn_keys8_packed<1, HI>(xv, yv, NP, P1, NQ, Q1, acc);
            if (q == 2) nn_keys8_packed<2, HI>(xv, yv, NP, P1, NQ, Q1, acc);
            if (q == 3) nn_keys8_packed<3, HI>(xv, yv, NP, P1, NQ, Q1, acc);
          }
        }
        // lane minimum -> 32-bit key (d << 16) | c' with the GLOBAL agent index (or 0xFFFF - index for HI)
        const unsigned m16 = min(acc & 0xFFFFu, acc >> 16);
        const unsigned l = HI ? 31u - (m16 & 31u) : (m16 & 31u);
        const unsigned idx = ((unsigned)(li + (int)(l >> 3) * LPE) << 3) | (l & 7u);
        unsigned key = ((m16 >> 5) << 16) | (HI ? 0xFFFFu - idx : idx);
        if (!EXACT && li >= NV) key = 0xFFFFFFFFu;               // a lane without a single vector of the env
#pragma unroll
        for (int d = LPE / 2; d > 0; d >>= 1) key = min(key, __shfl_xor_sync(gmask, key, d));
        return key;
      };
      const unsigned klo = keys(std::false_type{});
      target = (int)(klo & 0xFFFFu);
      if (2 * (int)(klo >> 16) >= S) {                           // ties go to the HIGHEST index only at d_min >= S/2 (nn_pick)
        const unsigned khi = keys(std::true_type{});
        target = nn_pick(klo, khi, S);
      }
      if (staged) {                                              // the new target's OLD cell, from the staged rows
        const int v = target >> 3, e = target & 7;
        tx = reinterpret_cast<const int16_t*>(rw + v)[e];        // rows are [q][li]: vector v = li + q LPE sits at index v
        ty = reinterpret_cast<const int16_t*>(rw + V * LPE + v)[e];
      } else {
        const size_t a = (size_t)env * (size_t)N + (size_t)target;
        tx = io.xo[a];
        ty = io.yo[a];
      }
    }
    int orient;
    pursue(px, py, tx, ty, S, px, py, orient);                   // SDE:142-165
    const unsigned kx = rep2(px), ky = rep2(py);
    unsigned acc = 0xFFFFFFFFu;                                  // SDE:307-309 on the NEW positions
#pragma unroll
    for (int q = 0; q < V; ++q)
      if (EXACT || li + q * LPE < NV) hit8_packed(mb[q * 32 + lane], mb[(V + q) * 32 + lane], kx, ky, acc);
#pragma unroll
    for (int d = LPE / 2; d > 0; d >>= 1) acc = __vminu2(acc, __shfl_xor_sync(gmask, acc, d));
    unsigned hit = 0xFFFFFFFFu;
    if (zero_half(acc) != 0u) {                                  // rare: resolve the lowest index standing on the cell
#pragma unroll
      for (int q = 0; q < V; ++q)
        if (EXACT || li + q * LPE < NV) hit8(mb[q * 32 + lane], mb[(V + q) * 32 + lane], kx, ky, (li + q * LPE) * 8, hit);
#pragma unroll
      for (int d = LPE / 2; d > 0; d >>= 1) hit = min(hit, __shfl_xor_sync(gmask, hit, d));
    }
    if (li == 0) {
      const int v = target >> 3, e = target & 7;                 // the target's NEW cell = what the next step pursues
      const int slot = (v / LPE) * 32 + g * LPE + (v % LPE);
      const int nx = reinterpret_cast<const int16_t*>(mb + slot)[e];
      const int ny = reinterpret_cast<const int16_t*>(mb + V * 32 + slot)[e];
      reinterpret_cast<unsigned*>(io.tcell)[env] = pack2(nx, ny);
      predator_write(io, env, ppx, ppy, px, py, target, orient, hit);
    }
  }
}

// shapes <LPE, V> by N (N % 8 == 0, N <= 256)
template <int LPE, int V, bool EXACT>
static cudaError_t predator_async_launch_x(const PredIO& io, int num_sms, cudaStream_t s) {
  using C = PredAsyncCfg<LPE, V, PA_W>;
  static_assert(C::NPIECES <= 32, "one state piece per lane");
  static bool opted[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!opted[dev]) {
    e = cudaFuncSetAttribute(predator_async_kernel<LPE, V, PA_W, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::CTA_BYTES);
    if (e != cudaSuccess) return e;
    opted[dev] = true;
  }
  const int tiles = (io.E + C::EW - 1) / C::EW;
  const int per_sm = std::max(1, std::min(32, (int)((227 * 1024) / (C::CTA_BYTES + 1024))));
  // a warp keeps at least 6 tiles (the pipeline's depth plus its prologue) before another CTA is added
  const long long want = ((long long)tiles + 6 * (PA_THREADS / 32) - 1) / (6 * (PA_THREADS / 32));
  const int grid = (int)std::min<long long>(want, (long long)num_sms * per_sm);
  predator_async_kernel<LPE, V, PA_W, EXACT><<<grid, PA_THREADS, C::CTA_BYTES, s>>>(io, tiles);
  return cudaGetLastError();
}
template <int LPE, int V>
static cudaError_t predator_async_launch_t(const PredIO& io, int num_sms, cudaStream_t s) {
  return io.N == 8 * LPE * V ? predator_async_launch_x<LPE, V, true>(io, num_sms, s)
                             : predator_async_launch_x<LPE, V, false>(io, num_sms, s);
}

// true: launched (or failed with *err set); false: shape not covered, the caller uses the register kernels
static inline bool predator_async_try(const PredIO& io, int num_sms, cudaStream_t s, cudaError_t* err) {
  const int N = io.N;
  // (the standalone entry point's case: tapes bound, no frozen mask -- swarm_step reaches the predator kernels only
  // with N > 256)
  if ((N & 7) || N > 256 || !io.tcell || !io.retarget || io.frozen || (io.E & 3) || io.S > 2048) return false;
  if (((uintptr_t)io.px | (uintptr_t)io.py | (uintptr_t)io.target | (uintptr_t)io.tcell | (uintptr_t)io.retarget |
       (uintptr_t)io.xo | (uintptr_t)io.yo | (uintptr_t)io.xn | (uintptr_t)io.yn) & 15)
    return false;                                              // 16-byte cp.async pieces
  const int nv = N >> 3;
  if (nv == 1) *err = predator_async_launch_t<1, 1>(io, num_sms, s);
  else if (nv == 2) *err = predator_async_launch_t<1, 2>(io, num_sms, s);
  else if (nv <= 4) *err = predator_async_launch_t<1, 4>(io, num_sms, s);
  else if (nv <= 8) *err = predator_async_launch_t<2, 4>(io, num_sms, s);
  else if (nv <= 16) *err = predator_async_launch_t<4, 4>(io, num_sms, s);
  else *err = predator_async_launch_t<8, 4>(io, num_sms, s);
  return true;
}


// ------------------------------------------------------------------------------------------------
// Large environments (N >= 4096): a CTA streams one env after another through the same kind of ring, one thread = one
// landing slot of PB_U x + PB_U y vectors per stage.  The chunk sequence of a CTA is
//     for each of its envs:  [old positions, only if the env retargets]  then  new positions
// and the loads of the next PB_W + 1 chunks are always in flight, across pass and env boundaries, while the CTA computes
// on the current chunk or sits in the block reduction that ends a pass (arg-min of the retarget, termination test).
// predator_block_kernel (staged.cuh) instead pays per env: one trip for the state, one per unrolled batch of loads and
// the reductions, with nothing in flight in between (0.67 of the HBM peak at 8,192 x 16,384).
// Per-env state travels 8 envs ahead through a 16-slot ring of 4-byte cp.async pieces (aligned-down words for the int16 /
// u8 arrays: E % 4 == 0).
// ------------------------------------------------------------------------------------------------
#ifndef PB_U
#define PB_U 2
#endif
#ifndef PB_W
#define PB_W 1
#endif
constexpr int PB_THREADS = 256;
constexpr int PB_NST = PB_W + 2;                       // stages of the landing ring
constexpr int PB_LEAD = 8, PB_RING = 16, PB_SWORDS = 8;  // state ring: {px pair, py pair, target, tcell, retarget quad, frozen quad}
constexpr int PB_SMEM = PB_NST * 2 * PB_U * PB_THREADS * 16 + PB_RING * PB_SWORDS * 4;

__global__ void __launch_bounds__(PB_THREADS) predator_block_async_kernel(const PredIO io) {
  extern __shared__ uint4 pb_smem[];
  __shared__ unsigned red[2][2][PB_THREADS / 32];        // [parity][value][warp]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int N = io.N, S = io.S, NV = N >> 3, E = io.E;
  const int CPE = (NV + PB_THREADS * PB_U - 1) / (PB_THREADS * PB_U);   // chunks per pass
  const int b = blockIdx.x, G = gridDim.x;
  const int my = b < E ? (E - b + G - 1) / G : 0;        // envs b, b + G, ...
  uint4* const ring = pb_smem;                           // [stage][x|y][u][thread]
  unsigned* const sring = reinterpret_cast<unsigned*>(pb_smem + PB_NST * 2 * PB_U * PB_THREADS);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)t * 16u;
  const unsigned sring_s = (unsigned)__cvta_generic_to_shared(sring);
  int rpar = 0;
  // one __syncthreads per reduction: warp minima into red[parity], everyone folds the 8 warp values
  auto block_min2 = [&](unsigned& a, unsigned& c) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a = min(a, __shfl_xor_sync(0xFFFFFFFFu, a, d));
      c = min(c, __shfl_xor_sync(0xFFFFFFFFu, c, d));
    }
    if (lane == 0) { red[rpar][0][warp] = a; red[rpar][1][warp] = c; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < PB_THREADS / 32; ++w) { a = min(a, red[rpar][0][w]); c = min(c, red[rpar][1][w]); }
    rpar ^= 1;
  };
  auto issue_state = [&](int j) {                        // threads 0..5, one piece each
    if (j >= my || t >= 6) return;
    const int env = b + j * G;
    const unsigned dst = sring_s + (unsigned)((j & (PB_RING - 1)) * PB_SWORDS + t) * 4u;
    if (t == 0) cp_async4(dst, io.px + (env & ~1));
    else if (t == 1) cp_async4(dst, io.py + (env & ~1));
    else if (t == 2) cp_async4(dst, io.target + env);
    else if (t == 3) { if (io.tcell) cp_async4(dst, io.tcell + 2 * (size_t)env); }
    else if (t == 4) {
      if (io.retarget) cp_async4(dst, io.retarget + (env & ~3));
      else sring[(j & (PB_RING - 1)) * PB_SWORDS + 4] =
          (unsigned)philox_retarget(io.rng_seed_lo, io.rng_seed_hi, io.rng_step, env + (int)io.env_offset) << (8 * (env & 3));
    } else if (io.frozen) cp_async4(dst, io.frozen + (env & ~3));
  };
  auto sword = [&](int j, int w) -> unsigned { return sring[(j & (PB_RING - 1)) * PB_SWORDS + w]; };
  auto frozen_of = [&](int j) -> bool { return io.frozen && ((sword(j, 5) >> (8 * ((b + j * G) & 3))) & 0xFFu) != 0u; };
  auto first_pass = [&](int j) -> int {                  // 0: old positions first (retarget), 1: new positions only
    return (!frozen_of(j) && ((sword(j, 4) >> (8 * ((b + j * G) & 3))) & 0xFFu) != 0u) ? 0 : 1;
  };
  struct Cur { int j, pass, c; };
  auto advance = [&](Cur& q) {
    if (++q.c < CPE) return;
    q.c = 0;
    if (q.pass == 0) { q.pass = 1; return; }
    ++q.j;
    q.pass = q.j < my ? first_pass(q.j) : 1;
  };
  auto issue_chunk = [&](const Cur& q, int stage) {
    if (q.j >= my) return;
    const size_t vb = (size_t)(b + q.j * G) * (size_t)NV;
    const uint4* xs = reinterpret_cast<const uint4*>(q.pass == 0 ? io.xo : io.xn) + vb;
    const uint4* ys = reinterpret_cast<const uint4*>(q.pass == 0 ? io.yo : io.yn) + vb;
    const unsigned dst = ring_s + (unsigned)stage * (2 * PB_U * PB_THREADS * 16);
#pragma unroll
    for (int u = 0; u < PB_U; ++u) {
      const int v = (q.c * PB_U + u) * PB_THREADS + t;
      if (v < NV) {
        cp_async16(dst + u * (PB_THREADS * 16), xs + v);
        cp_async16(dst + (PB_U + u) * (PB_THREADS * 16), ys + v);
      }
    }
  };

  if (my == 0) return;
#pragma unroll
  for (int j = 0; j < PB_LEAD; ++j) issue_state(j);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  Cur iss = {0, first_pass(0), 0}, con = iss;
#pragma unroll
  for (int d = 0; d <= PB_W; ++d) {
    issue_chunk(iss, d);
    cp_async_commit();
    advance(iss);
  }
  int stage = 0;
  // per-env registers
  int px = 0, py = 0, ppx = 0, ppy = 0, target = 0, orient = 0, tv = 0;
  unsigned klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu, hit = 0xFFFFFFFFu, kx = 0u, ky = 0u;
  unsigned NP = 0u, P1 = 0u, NQ = 0u, Q1 = 0u;
  bool frz = false;
  while (con.j < my) {
    cp_async_wait<PB_W>();                               // this thread's slots of the current chunk have landed
    {
      int is = stage + PB_W + 1;
      if (is >= PB_NST) is -= PB_NST;
      issue_chunk(iss, is);
      if (con.c == 0 && con.pass == 1) issue_state(con.j + PB_LEAD);   // once per env, 8 envs ahead
      cp_async_commit();
      advance(iss);
    }
    const int env = b + con.j * G;
    const uint4* slot = ring + stage * (2 * PB_U * PB_THREADS) + t;
    if (con.c == 0 && (con.pass == 0 || first_pass(con.j) == 1)) {     // first chunk of the env: its state
      px = (int)(int16_t)((sword(con.j, 0) >> (16 * (env & 1))) & 0xFFFFu);
      py = (int)(int16_t)((sword(con.j, 1) >> (16 * (env & 1))) & 0xFFFFu);
      ppx = px; ppy = py;
      target = min(max((int)sword(con.j, 2), 0), N - 1);
      frz = frozen_of(con.j);
      klo = khi = hit = 0xFFFFFFFFu;
      if (con.pass == 0) {
        NP = rep2(-px); P1 = rep2(px + 1); NQ = rep2(-py); Q1 = rep2(py + 1);
      } else {                                           // no retarget: pursue the cached cell (or read it: legacy callers)
        int tx, ty;
        if (io.tcell) { const unsigned tc = sword(con.j, 3); tx = (int)(int16_t)(tc & 0xFFFFu); ty = (int)(int16_t)(tc >> 16); }
        else { const size_t a = (size_t)env * (size_t)N + (size_t)target; tx = io.xo[a]; ty = io.yo[a]; }
        pursue(px, py, tx, ty, S, px, py, orient);
        kx = rep2(px); ky = rep2(py); tv = target >> 3;
      }
    }
    if (con.pass == 0) {                                 // SDE:111-131 on the OLD positions
#pragma unroll
      for (int u = 0; u < PB_U; ++u) {
        const int v = (con.c * PB_U + u) * PB_THREADS + t;
        if (v < NV) {
          unsigned lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;
          const uint4 xv = slot[u * PB_THREADS], yv = slot[(PB_U + u) * PB_THREADS];
          nn_keys8_packed<0, false>(xv, yv, NP, P1, NQ, Q1, lo);
          nn_keys8_packed<0, true>(xv, yv, NP, P1, NQ, Q1, hi);
          const unsigned l16 = min(lo & 0xFFFFu, lo >> 16), h16 = min(hi & 0xFFFFu, hi >> 16);
          klo = min(klo, ((l16 >> 5) << 16) | ((unsigned)v * 8u + (l16 & 7u)));
          khi = min(khi, ((h16 >> 5) << 16) | (0xFFFFu - ((unsigned)v * 8u + (7u - (h16 & 7u)))));
        }
      }
      if (con.c == CPE - 1) {                            // arg-min over the env, then the pursuit (SDE:138-165)
        block_min2(klo, khi);
        target = nn_pick(klo, khi, S);
        const size_t a = (size_t)env * (size_t)N + (size_t)target;
        pursue(px, py, (int)io.xo[a], (int)io.yo[a], S, px, py, orient);
        kx = rep2(px); ky = rep2(py); tv = target >> 3;
      }
    } else {                                             // SDE:307-309 on the NEW positions
#pragma unroll
      for (int u = 0; u < PB_U; ++u) {
        const int v = (con.c * PB_U + u) * PB_THREADS + t;
        if (v < NV) {
          const uint4 xv = slot[u * PB_THREADS], yv = slot[(PB_U + u) * PB_THREADS];
          unsigned acc = 0xFFFFFFFFu;
          hit8_packed(xv, yv, kx, ky, acc);
          if (zero_half(acc) != 0u) hit8(xv, yv, kx, ky, v * 8, hit);
          if (v == tv && io.tcell && !frz) reinterpret_cast<unsigned*>(io.tcell)[env] = cell8(xv, yv, target & 7);
        }
      }
      if (con.c == CPE - 1) {
        unsigned dummy = 0xFFFFFFFFu;
        block_min2(hit, dummy);
        if (t == 0) {
          if (frz) predator_frozen(io, env);
          else predator_write(io, env, ppx, ppy, px, py, target, orient, hit);
        }
      }
    }
    advance(con);
    if (++stage == PB_NST) stage = 0;
  }
}

// true: launched (ok or *err set); false: shape not covered
static inline bool predator_block_async_try(const PredIO& io, int num_sms, cudaStream_t s, cudaError_t* err) {
  const int N = io.N;
  if ((N & 7) || N < 4096 || N > 65528 || (io.E & 3) || io.S > 2048) return false;
  if (((uintptr_t)io.px | (uintptr_t)io.py | (uintptr_t)io.target | (uintptr_t)io.tcell | (uintptr_t)io.retarget |
       (uintptr_t)io.frozen) & 3)
    return false;
  if (((uintptr_t)io.xo | (uintptr_t)io.yo | (uintptr_t)io.xn | (uintptr_t)io.yn) & 15) return false;
  static bool opted[64] = {};
  int dev = 0;
  *err = cudaGetDevice(&dev);
  if (*err != cudaSuccess) return true;
  if (dev < 0 || dev >= 64) return false;
  if (!opted[dev]) {
    *err = cudaFuncSetAttribute(predator_block_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM);
    if (*err != cudaSuccess) return true;
    opted[dev] = true;
  }
  const int per_sm = std::max(1, std::min(8, (227 * 1024) / (PB_SMEM + 1024 + 256)));
  const int grid = (int)std::min<long long>(io.E, (long long)num_sms * per_sm);
  predator_block_async_kernel<<<grid, PB_THREADS, PB_SMEM, s>>>(io);
  *err = cudaGetLastError();
  return true;
}

}  // namespace swarm
