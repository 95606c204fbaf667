// swarm_abi.cu -- host side of the C ABI declared in include/swarm_abi.h (libswarm_b200.so).
// Plain pointers in, kernel launches on the caller's stream out; no torch, no CPU fallback.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "launchers.h"
#include "staged.cuh"

namespace swarm {

static thread_local std::string g_last_error;

struct Ctx {
  swarm_config cfg;
  swarm_state st;
  bool has_state = false;
  const int16_t* place = nullptr;
  const int16_t* ptape = nullptr;
  bool dev_reset = false;     // reset tapes drawn in-kernel (swarm_set_device_reset_tapes)
  uint64_t rst_seed = 0;
  std::string err;
  std::string path_name;
  bool fused = false;
  bool tiny = false;          // thread-per-env kernel (steps without a visual-field gate)
  int tiny_n = 8;             // its register capacity: 8 or 16 agents
  bool force_pairwise = false;  // SWARM_PATH_STAGED_PAIRWISE: never use the box-count reward path
  bool split = false;         // lane-group kernels: two launches per step (transition, reward), see fused_step_kernel
  bool box_near = false;      // ... and its reward launch may count {rep, att} by box popcounts
  bool box_one_launch = false;      // non-vf steps: single launch with box popcounts (N > 32 or 8 agents per lane)
  uint8_t* eff_scratch = nullptr;   // [E,N] effective actions between the two launches (when the caller binds no eff_action)
  int lpe = 0, apl = 0;
  int bitmap_words = 0;
  int launches = 0;
  int num_sms = 148;
  // staged scratch
  StagedScratch sc{};
  void* scratch_block = nullptr;
  int per_env_threads = 256;
  int coll_bm_words = 0;      // > 0: collide_bitmap_kernel (per-env occupancy bitmap fits in shared memory)
  int coll_smem_bytes = 0;
  BoxPlan box{};              // sub-quadratic reward counts inside collide_bitmap_kernel (steps without vf_size)
  // statistics: STAT_SLOTS results in flight (device slot + pinned host slot + completion event each); the reduction
  // runs on the caller's stream, the all-reduce and the copy to the host on the handle's side stream
  static constexpr int STAT_SLOTS = 4;
  double* d_stats = nullptr;        // [STAT_SLOTS][8]
  double* d_partial = nullptr;      // [stat_grid][8] per-CTA partial sums
  double* h_stats = nullptr;        // pinned [STAT_SLOTS][8]
  int stat_grid = 1;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_reduced[STAT_SLOTS] = {};
  cudaEvent_t ev_done[STAT_SLOTS] = {};
  unsigned long long stat_begun = 0, stat_ended = 0;
  // nccl
  void* nccl_lib = nullptr;
  void* nccl_comm = nullptr;
  bool nccl_owned = false;
  ~Ctx() {                     // owns only the library scratch; state / tapes / outputs belong to the caller
    if (scratch_block) cudaFree(scratch_block);
    if (eff_scratch) cudaFree(eff_scratch);
    if (d_stats) cudaFree(d_stats);
    if (d_partial) cudaFree(d_partial);
    if (h_stats) cudaFreeHost(h_stats);
    for (int q = 0; q < STAT_SLOTS; ++q) {
      if (ev_reduced[q]) cudaEventDestroy(ev_reduced[q]);
      if (ev_done[q]) cudaEventDestroy(ev_done[q]);
    }
    if (side) cudaStreamDestroy(side);
  }
};

static int fail(Ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(c, expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail((c), SWARM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void fill_thresholds(StepParams& p, int N, int S, int att, int rep, unsigned flags = 0u) {
  p.N = N;
  p.S = S;
  p.t_rep = clampi(rep, 0, S);
  p.t_att = clampi(att, 0, S);
  p.t_repf = clampi((long long)S - rep + 1 > S ? S : (S - rep + 1), 0, S);
  p.t_attf = clampi((long long)S - att + 1 > S ? S : (S - att + 1), 0, S);
  // literal diagonal terms of the N x N sums (D = 0, S - D = S): SDE:336-342
  p.diag_rep = (0 < rep ? 1 : 0) + (S < rep ? 1 : 0);
  p.diag_att = (((0 < att) == (0 >= rep)) ? 1 : 0) + (((S < att) == (S >= rep)) ? 1 : 0);
  p.inv_nn = 1.0 / ((double)N * (double)(N - 1));
  // far thresholds beyond half the grid: a pair at distance >= t needs both agents in opposite border bands
  p.far_border = (2 * p.t_repf > S && 2 * p.t_attf > S) ? 1 : 0;
  p.far_band = S - std::min(p.t_repf, p.t_attf);
  if (flags & SWARM_FLAG_NO_FAR_BORDER) p.far_border = 0;  // kernel A/B runs / tests of the 4-threshold sweep
}

// Box-count plan: per threshold the cheaper of the direct box and the complement (cost in bitmap word reads per
// agent).  The caller enables it when that cost is far below the pairwise sweep's.
static BoxPlan make_box_plan(int N, int S, int att, int rep, double& cost) {
  StepParams tp;
  memset(&tp, 0, sizeof(tp));
  fill_thresholds(tp, N, S, att, rep);
  const int ts[4] = {tp.t_rep, tp.t_att, tp.t_repf, tp.t_attf};
  BoxPlan bp{};
  cost = 0.0;
  bp.fast_rmax = -1;
  for (int q = 0; q < 4; ++q) {
    const int t = ts[q];
    bp.t[q] = t;
    if (t <= 0) bp.mode[q] = BOX_ZERO;
    else if (t >= S) bp.mode[q] = BOX_ALL;
    else {
      const double side = 2.0 * t - 1.0, direct = side * (side / 32.0 + 2.0);
      const double b = (double)(S - t), comp = 8.0 + 4.0 * b * (b / 32.0 + 2.0);
      if (direct <= comp) { bp.mode[q] = BOX_DIRECT; cost += direct; }
      else { bp.mode[q] = BOX_COMPLEMENT; bp.need_hist = 1; cost += comp; }
    }
  }
  // single-sweep variant for the fused kernel: word-aligned rows and small direct boxes
  int rmax = -1;
  bool ok = (S % 32) == 0;
  for (int q = 0; q < 4; ++q)
    if (bp.mode[q] == BOX_DIRECT) { rmax = std::max(rmax, bp.t[q] - 1); ok = ok && bp.t[q] - 1 <= 15; }
  if (ok && rmax >= 0) bp.fast_rmax = rmax;
  return bp;
}

static StepParams base_params(const Ctx* c) {
  StepParams p;
  memset(&p, 0, sizeof(p));
  const swarm_config& g = c->cfg;
  p.E = g.num_envs;
  fill_thresholds(p, g.num_agents, g.grid_size, g.attraction_thresh, g.repulsion_thresh, g.flags);
  p.t_vf = -1;
  p.env_offset = g.env_offset;
  p.vf_dense = (g.flags & SWARM_FLAG_VF_DENSE) ? 1 : 0;
  p.tiny_scalar = (g.flags & SWARM_FLAG_TINY_SCALAR) ? 1 : 0;
  p.w_att = p.w_rep = p.w_ali = 1.0;
  p.eat = g.predator_eat_rew;
  p.K = g.slab_len; p.PL = g.place_len; p.KP = g.pred_attempts; p.R = g.reset_slots;
  p.auto_reset = g.auto_reset; p.track_stats = g.track_stats;
  p.bitmap_words = c->bitmap_words;
  p.one = 1;
  const swarm_state& s = c->st;
  p.x = s.x; p.y = s.y; p.px = s.px; p.py = s.py; p.target = s.target; p.orient = s.pred_orient;
  p.frozen = s.frozen; p.err = s.err; p.reset_count = s.reset_count;
  p.tcell = c->fused ? nullptr : s.tcell;              // state of the staged path only
  p.ep_rec = s.ep_rec; p.fin_count = s.fin_count; p.fin_len = s.fin_len; p.fin_ret = s.fin_ret;
  p.place = c->place; p.ptape = c->ptape;
  p.rst_seed_lo = (unsigned)(c->rst_seed & 0xFFFFFFFFull); p.rst_seed_hi = (unsigned)(c->rst_seed >> 32);
  return p;
}

// ---- staged scratch ---------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int staged_alloc(Ctx* c) {
  const int E = c->cfg.num_envs, N = c->cfg.num_agents;
  StagedScratch& sc = c->sc;
  sc.NW = (N + 1) / 2;
  int log2hs = 4;
  while ((1 << log2hs) < 2 * N) ++log2hs;
  sc.log2hs = log2hs;
  sc.env_grid = std::min(E, 2 * c->num_sms);
  int threads = 128;
  while (threads < 1024 && threads * 4 < N) threads <<= 1;
  c->per_env_threads = threads;
  {
    const long long cells = (long long)c->cfg.grid_size * c->cfg.grid_size;
    const long long words = (cells + 31) / 32;
    const long long bytes = words * 4 + (long long)COLL_HC * 8;
    int dev_max = 0;
    cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->cfg.device);
    if (bytes + 1024 <= dev_max && (N + threads - 1) / threads <= 16) {
      c->coll_bm_words = (int)words;
      c->coll_smem_bytes = (int)bytes;
      const int S = c->cfg.grid_size;
      double cost = 0.0;
      BoxPlan bp = make_box_plan(N, S, c->cfg.attraction_thresh, c->cfg.repulsion_thresh, cost);
      const long long hist_bytes = bp.need_hist ? 2LL * S * 4 : 0;
      // (a box word-op is ~5 instructions, the symmetric pair sweep ~1.5 ALU instructions per agent pair)
      if (cost * 3.0 < (double)N && bytes + hist_bytes + 1024 <= dev_max) {
        bp.enabled = 1;
        c->coll_smem_bytes = (int)(bytes + hist_bytes);
      }
      c->box = bp;
      // per-function, per-device attribute shared by all handles: only ever raised (to the opt-in maximum), so that a
      // later handle with a smaller grid cannot lower it under this one's launches
      CUDA_TRY(c, smem_opt_in(collide_bitmap_kernel, c->coll_smem_bytes));
      CUDA_TRY(c, smem_opt_in(reset_bitmap_kernel, (int)(words * 4 + (long long)COLL_HC * 8)));
      CUDA_TRY(c, smem_opt_in(boxcount_kernel, (int)(words * 4 + 2LL * S * 4)));
    }
  }
  const size_t hs = (size_t)1 << log2hs;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_nx = take((size_t)E * N * 2), o_ny = take((size_t)E * N * 2), o_eff = take((size_t)E * N);
  const size_t o_hash = take((size_t)sc.env_grid * 2 * hs * 4), o_aslot = take((size_t)sc.env_grid * N * 4);
  const size_t o_pairw = take((size_t)E * sc.NW * 16), o_vfw = take((size_t)E * sc.NW * 16);
  const size_t o_pc = take((size_t)E * N * 32 + 64);
  const size_t o_cnt = take(256);
  const size_t o_ea = take((size_t)E * 12 * 4);
  CUDA_TRY(c, cudaMalloc(&c->scratch_block, off));
  char* b = (char*)c->scratch_block;
  sc.nx = (int16_t*)(b + o_nx); sc.ny = (int16_t*)(b + o_ny); sc.eff = (uint8_t*)(b + o_eff);
  sc.hash = (unsigned*)(b + o_hash); sc.aslot = (unsigned*)(b + o_aslot);
  sc.pairw = (uint4*)(b + o_pairw); sc.vfw = (uint2*)(b + o_vfw); sc.pcounts = (unsigned*)(b + o_pc);
  sc.counter = (unsigned*)(b + o_cnt);
  sc.envacc = (int*)(b + o_ea);
  CUDA_TRY(c, cudaMemset(b + o_hash, 0xFF, (size_t)sc.env_grid * hs * 4 * 2));   // keys = EMPTY ...
  // ... and counts = 0: keys and counts are interleaved per CTA ({keys[HS], cnt[HS]}), fix the counts
  for (int g = 0; g < sc.env_grid; ++g)
    CUDA_TRY(c, cudaMemset(b + o_hash + ((size_t)g * 2 * hs + hs) * 4, 0, hs * 4));
  return SWARM_OK;
}

static int pairwise_launch(const StepParams& p, const StagedScratch& sc, const int16_t* x, const int16_t* y,
                           const uint8_t* skip, bool vf, int num_sms, cudaStream_t s) {
  if (!vf) {                                   // symmetric tile sweep (every unordered tile pair once)
    PairSymIO io;
    io.E = p.E; io.N = p.N; io.S = p.S; io.NW = sc.NW;
    io.T = (p.N + 127) / 128;
    io.CH = (io.T + PS_CHUNK - 1) / PS_CHUNK;
    io.t_rep = p.t_rep; io.t_att = p.t_att; io.t_repf = p.t_repf; io.t_attf = p.t_attf;
    io.pairw = sc.pairw; io.skip = skip; io.pcounts = sc.pcounts; io.counter = sc.counter;
    io.items = (long long)p.E * io.T * io.CH;
    io.one = 1;
    const long long real_items = (io.items + 1) / 2 + (long long)p.E * io.T;   // upper bound of non-empty items
    long long grid = std::min<long long>((real_items + PS_WARPS - 1) / PS_WARPS, (long long)num_sms * 6);
    if (grid < 1) grid = 1;
    pairwise_sym_kernel<<<(int)grid, PS_WARPS * 32, 0, s>>>(io);
    return 0;
  }
  PairIO io;
  io.E = p.E; io.N = p.N; io.S = p.S; io.NW = sc.NW;
  io.t_rep = p.t_rep; io.t_att = p.t_att; io.t_repf = p.t_repf; io.t_attf = p.t_attf; io.t_vf = p.t_vf;
  io.pairw = sc.pairw; io.vfw = sc.vfw; io.x = x; io.y = y; io.skip = skip; io.pcounts = sc.pcounts;
  io.tiles_per_env = (p.N + PW_TILE - 1) / PW_TILE;
  io.chunks_per_env = (sc.NW + PW_JC - 1) / PW_JC;
  io.units = (long long)p.E * io.tiles_per_env * io.chunks_per_env;
  io.one = 1;
  // 4 resident CTAs per SM (128 threads, ~8-16 KB smem each); never more CTAs than units
  long long grid = std::min<long long>(io.units, (long long)num_sms * 4);
  if (grid < 1) grid = 1;
  pairwise_kernel<true><<<(int)grid, PW_THREADS, 0, s>>>(io);
  return 0;
}

// ---- statistics reduction: two stages, both in a fixed order (deterministic) ---------------------------
// stage 1: CTA b sums the envs b*256*k .. of its contiguous slice into partial[b][8]; stage 2: one CTA adds the partials.
constexpr int STAT_THREADS = 256;

__device__ __forceinline__ void stats_block_sum(double (&acc)[8], double* __restrict__ out) {
  __shared__ double sh[8][STAT_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double v = acc[q];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane == 0) sh[q][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double v = lane < STAT_THREADS / 32 ? sh[q][lane] : 0.0;
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
      if (lane == 0) out[q] = v;
    }
  }
}

__global__ void __launch_bounds__(STAT_THREADS) stats_partial_kernel(int E, int per_cta, const uint32_t* __restrict__ fin_count,
                                                                     const uint32_t* __restrict__ fin_len,
                                                                     const float* __restrict__ fin_ret,
                                                                     const uint32_t* __restrict__ ep_rec,
                                                                     double* __restrict__ partial) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int e0 = blockIdx.x * per_cta, e1 = min(E, e0 + per_cta);
  for (int e = e0 + threadIdx.x; e < e1; e += STAT_THREADS) {
    const float4 r = reinterpret_cast<const float4*>(fin_ret)[e];
    acc[0] += (double)fin_count[e];
    acc[1] += (double)fin_len[e];
    acc[3] += (double)r.x; acc[4] += (double)r.y; acc[5] += (double)r.z; acc[6] += (double)r.w;
    acc[7] += (double)ep_rec[4 * (size_t)e];
  }
  stats_block_sum(acc, partial + 8 * (size_t)blockIdx.x);
}

__global__ void __launch_bounds__(STAT_THREADS) stats_final_kernel(int parts, const double* __restrict__ partial,
                                                                   double* __restrict__ out) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < parts; b += STAT_THREADS)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += partial[8 * (size_t)b + q];
  __shared__ double tot[8];
  stats_block_sum(acc, tot);
  __syncthreads();
  if (threadIdx.x == 0) {
    tot[2] = tot[3] + tot[4] + tot[5] + tot[6];
#pragma unroll
    for (int q = 0; q < 8; ++q) out[q] = tot[q];
  }
}

// ---- INT32 pipe micro-benchmarks ----------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, unsigned seed, unsigned* __restrict__ out) {
  constexpr int CH = 8;
  unsigned a[CH], b[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { a[c] = seed + threadIdx.x * 7u + c; b[c] = seed * 3u + c * 11u + blockIdx.x; }
  const unsigned k1 = seed | 1u, k2 = (seed >> 3) | 0x00010001u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (KIND == 0) {            // IADD3 dependent chains
        a[c] = a[c] + b[c] + k1;
        b[c] = b[c] + a[c] + k2;
      } else if (KIND == 1) {     // IMNMX / VIMNMX
        a[c] = (unsigned)max((int)a[c], (int)(b[c] ^ k1));
        b[c] = (unsigned)min((int)b[c], (int)(a[c] ^ k2));
      } else if (KIND == 2) {     // IADD3 (alu pipe) + IMAD (fma pipe)
        a[c] = a[c] + b[c] + k1;
        b[c] = b[c] * k1 + a[c];
      } else if (KIND == 3) {     // DPX 16x2
        a[c] = __viaddmin_s16x2(a[c], k1, b[c]);
        b[c] = __viaddmin_s16x2_relu(b[c], k2, a[c]);
      } else {                    // DPX 16x2 + IMAD
        a[c] = __viaddmin_s16x2(a[c], k1, b[c]);
        b[c] = b[c] * k1 + a[c];
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) r ^= a[c] ^ b[c];
  if (r == 0x12345678u) out[0] = r;   // keep the chains alive
}

}  // namespace swarm

using namespace swarm;

typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);

extern "C" {

int swarm_abi_version(void) { return SWARM_ABI_VERSION; }

const char* swarm_last_error(swarm_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return c ? c->err.c_str() : g_last_error.c_str();
}

int swarm_create(const swarm_config* cfg, swarm_handle* out) {
  if (!cfg || !out) return fail(nullptr, SWARM_E_INVALID, "swarm_create: null argument");
  if (cfg->abi_version != SWARM_ABI_VERSION) return fail(nullptr, SWARM_E_INVALID, "swarm_create: ABI version mismatch");
  if (cfg->num_envs < 1) return fail(nullptr, SWARM_E_INVALID, "swarm_create: num_envs < 1");
  if (cfg->num_agents < 2 || cfg->num_agents > SWARM_MAX_AGENTS)
    return fail(nullptr, SWARM_E_INVALID, "swarm_create: num_agents must be in [2, 16384] (reward divides by N(N-1))");
  if (cfg->grid_size < 2 || cfg->grid_size > SWARM_MAX_GRID)
    return fail(nullptr, SWARM_E_INVALID, "swarm_create: grid_size must be in [2, 16384]");
  if ((long long)cfg->grid_size * cfg->grid_size <= cfg->num_agents)
    return fail(nullptr, SWARM_E_INVALID, "swarm_create: need more cells than agents + predator");
  if (cfg->slab_len < 0 || cfg->place_len < 2 * cfg->num_agents || cfg->pred_attempts < 1 || cfg->reset_slots < 1)
    return fail(nullptr, SWARM_E_INVALID, "swarm_create: bad tape geometry");
  if ((cfg->path == SWARM_PATH_FUSED || cfg->path == SWARM_PATH_GROUP) && cfg->num_agents > SWARM_FUSED_MAX_AGENTS)
    return fail(nullptr, SWARM_E_INVALID, "swarm_create: fused path needs num_agents <= 256");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(nullptr, SWARM_E_CUDA, "swarm_create: no CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SWARM_E_INVALID, "swarm_create: bad device ordinal");
  {
    cudaError_t de = cudaSetDevice(cfg->device);
    if (de != cudaSuccess) return fail(nullptr, SWARM_E_CUDA, std::string("swarm_create: cudaSetDevice: ") + cudaGetErrorString(de));
  }
  int num_sms = 148;
  if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess)
    return fail(nullptr, SWARM_E_CUDA, "swarm_create: cannot query the device");
  Ctx* c = new Ctx();          // every failure path below deletes it
  c->cfg = *cfg;
  c->num_sms = num_sms;
  c->force_pairwise = cfg->path == SWARM_PATH_STAGED_PAIRWISE;
  c->fused = cfg->path == SWARM_PATH_FUSED || cfg->path == SWARM_PATH_GROUP ||
             (cfg->path == SWARM_PATH_AUTO && cfg->num_agents <= SWARM_FUSED_MAX_AGENTS);
  // thread per env: always for N <= 8; for 9..16 agents only when the batch is large enough to fill the GPU with one
  // thread per env (the lane-group kernel spreads a small batch over 8x more threads and has the shorter latency)
  const int big_min_envs = cfg->tiny16_min_envs > 0 ? cfg->tiny16_min_envs : 65536;
  c->tiny = c->fused && cfg->path != SWARM_PATH_GROUP &&
            (cfg->num_agents <= TINY_N || (cfg->num_agents <= TINY_N_BIG && cfg->num_envs >= big_min_envs));
  c->tiny_n = cfg->num_agents <= TINY_N ? TINY_N : TINY_N_BIG;
  if (c->fused) {
    choose_fused_shape(cfg->num_agents, c->lpe, c->apl);
    const long long cells = (long long)cfg->grid_size * cfg->grid_size;
    const long long words = ((cells + 31) / 32 + 3) / 4 * 4;
    c->bitmap_words = (words <= 2048 && cfg->num_agents >= 16) ? (int)words : 0;
    {
      // the reward launch of the two-launch step counts {rep, att} by box popcounts on a padded copy of the bitmap
      // (fused_step.cuh, box_near_counts) when rows are word aligned, both radii fit a 32-bit window and the far
      // thresholds come from the border pass; the per-env region is sized for the larger of the two layouts
      StepParams tp;
      memset(&tp, 0, sizeof(tp));
      fill_thresholds(tp, cfg->num_agents, cfg->grid_size, cfg->attraction_thresh, cfg->repulsion_thresh, cfg->flags);
      const int S = cfg->grid_size;
      const int bw = box_near_words(S, std::max(std::max(tp.t_rep, tp.t_att) - 1, 0));
      c->box_near = c->bitmap_words > 0 && (S % 32) == 0 && tp.t_rep <= 16 && tp.t_att <= 16 && tp.far_border && bw <= 2048 &&
                    !(cfg->flags & SWARM_FLAG_NO_BOX_NEAR);
      if (c->box_near) c->bitmap_words = std::max(c->bitmap_words, bw);
      c->box_one_launch = c->box_near && (cfg->num_agents > 32) && !(cfg->flags & SWARM_FLAG_FUSED_SPLIT);
    }
    // (A box-count variant of the reward on this per-env bitmap was measured for the lane-group kernels in round 1:
    // 0.42-0.45 ms / step at cfg3 against 0.275 for the ring sweep -- the shared-memory gathers of 32 lanes into random
    // bitmap rows serialise on bank conflicts.  It pays off from N ~ 1000 up, which is the staged path; removed here.)
    char buf[64];
    if (c->tiny) snprintf(buf, sizeof(buf), "tiny<%d> (vf: fused<%d,%d>)", c->tiny_n, c->lpe, c->apl);
    else snprintf(buf, sizeof(buf), "fused<%d,%d>%s", c->lpe, c->apl, c->bitmap_words ? "+bitmap" : "");
    c->path_name = buf;
  } else {
    c->path_name = "staged";   // refined after staged_alloc (box-count plan)
  }
  // the staged scratch also backs swarm_autoreset / masks on the fused path only when needed
  if (!c->fused) {
    int rc = staged_alloc(c);
    if (rc != SWARM_OK) { std::string m = c->err; delete c; return fail(nullptr, rc, m); }
    if (c->coll_bm_words > 0 && c->box.enabled && !c->force_pairwise) c->path_name = "staged+box (vf: staged pairwise)";
    else if (c->coll_bm_words > 0) c->path_name = "staged+bitmap";
  }
  if (c->fused) {
    // One launch per step for small batches (latency, graph rollouts); two launches once the batch is many waves deep
    // and the instruction-cache footprint of the whole step is what limits the SMs (measured at cfg3).
    const long long agents = (long long)cfg->num_envs * cfg->num_agents;
    c->split = agents >= (1LL << 22);   // measured: 1 M agents (8,192 x 128) 0.042 ms in one launch vs 0.046 in two; 8.4 M 0.272 vs 0.248
    if (cfg->flags & SWARM_FLAG_FUSED_SPLIT) c->split = true;
    if (cfg->flags & SWARM_FLAG_FUSED_ONE_LAUNCH) c->split = false;
    if (c->split) {
      const cudaError_t ae = cudaMalloc(&c->eff_scratch, (size_t)agents + 16);
      if (ae != cudaSuccess) { delete c; return fail(nullptr, SWARM_E_CUDA, "swarm_create: cudaMalloc (effective actions)"); }
      if (!c->tiny) c->path_name += (c->box_one_launch && c->apl != 8) ? "+box (vf: x2 launches)" : " x2 launches";
    } else if (c->box_one_launch && !c->tiny) {
      c->path_name += "+box";
    }
    const StepParams p0 = base_params(c);
    const cudaError_t pe = fused_prepare(c->lpe, c->apl, p0);
    if (pe != cudaSuccess) {
      const std::string m = std::string("swarm_create: shared memory opt-in failed: ") + cudaGetErrorString(pe);
      delete c;
      return fail(nullptr, SWARM_E_CUDA, m);
    }
  }
  {
    // statistics plumbing: multi-CTA partial sums (a contiguous slice of envs per CTA), result slots, side stream
    c->stat_grid = std::max(1, std::min(c->num_sms * 4, (cfg->num_envs + STAT_THREADS - 1) / STAT_THREADS));
    cudaError_t e = cudaMalloc(&c->d_stats, Ctx::STAT_SLOTS * 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_partial, (size_t)c->stat_grid * 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_stats, Ctx::STAT_SLOTS * 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    for (int q = 0; q < Ctx::STAT_SLOTS && e == cudaSuccess; ++q) {
      e = cudaEventCreateWithFlags(&c->ev_reduced[q], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[q], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
      const std::string m = std::string("swarm_create: statistics buffers: ") + cudaGetErrorString(e);
      delete c;
      return fail(nullptr, SWARM_E_CUDA, m);
    }
  }
  *out = reinterpret_cast<swarm_handle>(c);
  return SWARM_OK;
}

int swarm_destroy(swarm_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return SWARM_OK;
  cudaSetDevice(c->cfg.device);
  if (c->side) cudaStreamSynchronize(c->side);
  if (c->nccl_comm && c->nccl_lib && c->nccl_owned) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(c->nccl_lib, "ncclCommDestroy");
    if (f) f(c->nccl_comm);
  }
  delete c;
  return SWARM_OK;
}

int swarm_bind_state(swarm_handle h, const swarm_state* st) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !st) return fail(c, SWARM_E_INVALID, "swarm_bind_state: null argument");
  if (!st->x || !st->y || !st->px || !st->py || !st->target || !st->pred_orient || !st->frozen || !st->err ||
      !st->reset_count)
    return fail(c, SWARM_E_INVALID, "swarm_bind_state: a required state pointer is null");
  if (c->cfg.track_stats && (!st->ep_rec || !st->fin_count || !st->fin_len || !st->fin_ret))
    return fail(c, SWARM_E_INVALID, "swarm_bind_state: track_stats needs the statistics pointers");
  if (((uintptr_t)st->x | (uintptr_t)st->y | (uintptr_t)st->ep_rec | (uintptr_t)st->fin_ret) & 15)
    return fail(c, SWARM_E_INVALID, "swarm_bind_state: x / y / ep_rec / fin_ret must be 16-byte aligned");
  if ((uintptr_t)st->tcell & 3) return fail(c, SWARM_E_INVALID, "swarm_bind_state: tcell must be 4-byte aligned");
  c->st = *st;
  c->has_state = true;
  return SWARM_OK;
}

int swarm_bind_reset_tapes(swarm_handle h, const int16_t* place, const int16_t* ptape) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !place || !ptape) return fail(c, SWARM_E_INVALID, "swarm_bind_reset_tapes: null argument");
  c->place = place;
  c->ptape = ptape;
  c->dev_reset = false;
  return SWARM_OK;
}

int swarm_set_device_reset_tapes(swarm_handle h, uint64_t seed) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return fail(c, SWARM_E_INVALID, "swarm_set_device_reset_tapes: null handle");
  c->place = nullptr;
  c->ptape = nullptr;
  c->dev_reset = true;
  c->rst_seed = seed;
  return SWARM_OK;
}

const char* swarm_path_name(swarm_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return c ? c->path_name.c_str() : "";
}

int swarm_last_launch_count(swarm_handle h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return c ? c->launches : 0;
}

static int do_reset(Ctx* c, const uint8_t* mask, int mode, cudaStream_t s) {
  if (!c->has_state || (!c->dev_reset && (!c->place || !c->ptape)))
    return fail(c, SWARM_E_STATE, "reset: bind state and reset tapes first");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  StepParams p = base_params(c);
  c->launches = 0;
  if (c->fused) {
    CUDA_TRY(c, fused_reset_launch(c->lpe, c->apl, p, mask, mode == 0 ? 1 : 0, s));
  } else {
    if (c->coll_bm_words > 0)
      reset_bitmap_kernel<<<c->sc.env_grid, c->per_env_threads, c->coll_bm_words * 4 + COLL_HC * 8, s>>>(
          p, c->sc, c->coll_bm_words, mask, mode);
    else
      reset_kernel<<<c->sc.env_grid, c->per_env_threads, 0, s>>>(p, c->sc, mask, mode);
    CUDA_TRY(c, cudaGetLastError());
  }
  c->launches = 1;
  return SWARM_OK;
}

int swarm_reset(swarm_handle h, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return fail(nullptr, SWARM_E_INVALID, "swarm_reset: null handle");
  return do_reset(c, nullptr, 0, (cudaStream_t)stream);
}

int swarm_autoreset(swarm_handle h, const uint8_t* mask, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return fail(nullptr, SWARM_E_INVALID, "swarm_autoreset: null handle");
  return do_reset(c, mask, 1, (cudaStream_t)stream);
}

int swarm_step(swarm_handle h, const swarm_step_in* in, const swarm_step_out* out, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !in || !out) return fail(c, SWARM_E_INVALID, "swarm_step: null argument");
  if (!c->has_state) return fail(c, SWARM_E_STATE, "swarm_step: bind state first");
  if (c->cfg.auto_reset && !c->dev_reset && (!c->place || !c->ptape))
    return fail(c, SWARM_E_STATE, "swarm_step: auto_reset needs bound reset tapes");
  const bool dev_tapes = !in->retarget && !in->slab;          // both step tapes drawn in-kernel (Philox)
  if (!in->actions || (!dev_tapes && (!in->retarget || (!in->slab && c->cfg.slab_len > 0))))
    return fail(c, SWARM_E_INVALID, "swarm_step: null input tape (pass retarget and slab, or neither for device tapes)");
  if (!out->done || !out->eaten || !out->reward || !out->counts || !out->pred_prev || !out->pred_next ||
      !out->step_orient || !out->step_target || !out->consumed)
    return fail(c, SWARM_E_INVALID, "swarm_step: a required output pointer is null");
  if ((out->term_x == nullptr) != (out->term_y == nullptr))
    return fail(c, SWARM_E_INVALID, "swarm_step: term_x and term_y go together");
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  StepParams p = base_params(c);
  p.actions = in->actions; p.retarget = in->retarget; p.slab = in->slab;
  p.rng_seed_lo = (unsigned)(in->rng_seed & 0xFFFFFFFFull); p.rng_seed_hi = (unsigned)(in->rng_seed >> 32);
  p.rng_step = in->rng_step;
  p.w_att = in->w_attraction; p.w_rep = in->w_repulsion; p.w_ali = in->w_alignment;
  p.t_vf = in->vf_size < 0 ? -1 : clampi(in->vf_size >= p.S ? p.S : in->vf_size + 1, 0, p.S);
  p.done = out->done; p.eaten = out->eaten; p.reward = out->reward; p.counts = out->counts;
  p.pred_prev = out->pred_prev; p.pred_next = out->pred_next; p.step_orient = out->step_orient;
  p.step_target = out->step_target; p.consumed = out->consumed; p.term_x = out->term_x; p.term_y = out->term_y;
  p.agent_reward = out->agent_reward; p.agent_counts = out->agent_counts; p.eff_action = out->eff_action;
  p.agent_terms = out->agent_terms;
  if (out->agent_terms && p.N > 4096)
    return fail(c, SWARM_E_INVALID, "swarm_step: agent_terms (int16) needs num_agents <= 4096");
  if ((uintptr_t)out->agent_terms & 15) return fail(c, SWARM_E_INVALID, "swarm_step: agent_terms must be 16-byte aligned");
  const bool vf = p.t_vf >= 0;
  c->launches = 0;
  if (c->tiny && !vf) {
    if (p.N == c->tiny_n && ((uintptr_t)in->actions & 7))
      return fail(c, SWARM_E_INVALID, "swarm_step: actions must be 8-byte aligned");
    CUDA_TRY(c, tiny_step_launch(c->tiny_n, p, s, &c->launches));
    return SWARM_OK;
  }
  if (c->fused) {
    // {rep, att} by box popcounts where they apply.  From 33 agents up they also pay in the single-launch kernel, and with
    // them the single launch is at least as fast as two at every batch size (cfg3: 0.200 vs 0.201 ms at 65,536 envs,
    // 0.033 vs 0.040 at 8,192), so such steps never split; vf_size steps and the ring-sweep shapes keep the split.
    // (8 agents per lane: the single launch is the slower one at large batches -- 32,768 x 256: 0.363 vs 0.267 ms -- so
    // that shape splits by size as before and takes the box counts in either mode)
    const bool box_single = c->box_one_launch && !vf;
    p.split = (c->split && !(box_single && c->apl != 8)) ? 1 : 0;
    p.eff_buf = out->eff_action ? out->eff_action : c->eff_scratch;
    p.box_near = ((p.split || box_single) && c->box_near && !vf) ? 1 : 0;
    CUDA_TRY(c, fused_step_launch(c->lpe, c->apl, p, s));
    c->launches = p.split ? 2 : 1;
    return SWARM_OK;
  }
  // ---- staged path -------------------------------------------------------------------------------
  const StagedScratch& sc = c->sc;
  const long long total = (long long)p.E * p.N;
  {
    const long long nvec = std::max<long long>(1, total >> 3);
    const int grid = (int)std::min<long long>((nvec + 255) / 256, (long long)c->num_sms * 16);
    move_kernel<<<grid, 256, 0, s>>>(total, p.N, p.S, p.x, p.y, p.actions, sc.nx, sc.ny, sc.eff, p.err);
  }
  const bool box = c->coll_bm_words > 0 && c->box.enabled && !vf && !c->force_pairwise;
  int box_parts = 1;                           // CTAs per env for the box counts: spread few big envs over the SMs
  if (box) box_parts = std::max(1, std::min(8, c->num_sms / std::max(1, p.E)));
  int extra_launches = 0;
  if (c->coll_bm_words > 0) {
    BoxPlan bp = c->box;
    bp.enabled = box ? (box_parts > 1 ? 2 : 1) : 0;
    collide_bitmap_kernel<<<sc.env_grid, c->per_env_threads, c->coll_smem_bytes, s>>>(p, sc, c->coll_bm_words, bp);
    if (box && box_parts > 1) {
      const int smem = c->coll_bm_words * 4 + (bp.need_hist ? 2 * p.S * 4 : 0);
      boxcount_kernel<<<p.E * box_parts, 1024, smem, s>>>(p, sc, c->coll_bm_words, bp, box_parts);
      extra_launches = 1;
    }
  } else {
    collide_kernel<<<sc.env_grid, c->per_env_threads, 0, s>>>(p, sc);
  }
  {
    PredIO io;
    io.E = p.E; io.N = p.N; io.S = p.S; io.xo = p.x; io.yo = p.y; io.xn = sc.nx; io.yn = sc.ny;
    io.retarget = p.retarget; io.frozen = p.frozen; io.px = p.px; io.py = p.py; io.target = p.target;
    io.rng_seed_lo = p.rng_seed_lo; io.rng_seed_hi = p.rng_seed_hi; io.rng_step = p.rng_step;
    io.env_offset = p.env_offset; io.tcell = p.tcell;
    io.orient = p.orient; io.pred_prev = p.pred_prev; io.pred_next = p.pred_next; io.step_orient = p.step_orient;
    io.step_target = p.step_target; io.done = p.done; io.eaten = p.eaten; io.err = p.err;
    launch_predator(io, c->num_sms, s);
  }
  c->launches = 5 + extra_launches;
  if (!box) {                                  // tiled pairwise kernels compute the threshold counts
    const long long words = (long long)p.E * sc.NW;
    pairprep_kernel<<<(int)((words + 255) / 256), 256, 0, s>>>(p.E, p.N, p.S, sc.NW, sc.nx, sc.ny, sc.eff, sc.pairw,
                                                              vf ? sc.vfw : nullptr, sc.pcounts, sc.counter, sc.envacc);
    pairwise_launch(p, sc, sc.nx, sc.ny, p.done, vf, c->num_sms, s);
    c->launches = 7;
  }
  {
    const int ga = (int)((total + 255) / 256), ge = (p.E + 127) / 128;
    if (vf) { finalize_agents_kernel<true><<<ga, 256, 0, s>>>(p, sc); finalize_env_kernel<true><<<ge, 128, 0, s>>>(p, sc); }
    else { finalize_agents_kernel<false><<<ga, 256, 0, s>>>(p, sc); finalize_env_kernel<false><<<ge, 128, 0, s>>>(p, sc); }
  }
  if (p.auto_reset) {
    if (c->coll_bm_words > 0)
      reset_bitmap_kernel<<<sc.env_grid, c->per_env_threads, c->coll_bm_words * 4 + COLL_HC * 8, s>>>(p, sc, c->coll_bm_words,
                                                                                                      p.done, 1);
    else
      reset_kernel<<<sc.env_grid, c->per_env_threads, 0, s>>>(p, sc, p.done, 1);
    c->launches += 1;
  }
  CUDA_TRY(c, cudaGetLastError());
  return SWARM_OK;
}

// ---- standalone kernels ------------------------------------------------------------------------------

int swarm_move(int32_t E, int32_t N, int32_t S, const int16_t* x, const int16_t* y, const uint8_t* actions,
               int16_t* out_x, int16_t* out_y, void* stream) {
  if (E < 1 || N < 1 || S < 2 || !x || !y || !actions || !out_x || !out_y)
    return fail(nullptr, SWARM_E_INVALID, "swarm_move: bad argument");
  if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)out_x | (uintptr_t)out_y) & 15 || ((uintptr_t)actions & 7))
    return fail(nullptr, SWARM_E_INVALID, "swarm_move: pointers must be 16-byte (positions) / 8-byte (actions) aligned");
  const long long total = (long long)E * N;
  const long long nvec = std::max<long long>(1, total >> 3);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<long long>((nvec + 255) / 256, (long long)sms * 16);
  move_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, N, S, x, y, actions, out_x, out_y, nullptr, nullptr);
  CUDA_TRY(nullptr, cudaGetLastError());
  return SWARM_OK;
}

int swarm_predator(int32_t E, int32_t N, int32_t S, const int16_t* x_old, const int16_t* y_old,
                   const int16_t* x_new, const int16_t* y_new, const uint8_t* retarget, int16_t* px, int16_t* py,
                   int32_t* target, uint8_t* pred_orient, int16_t* pred_prev, uint8_t* done, int32_t* eaten,
                   int16_t* tcell, void* stream) {
  if (E < 1 || N < 1 || S < 2 || !x_old || !y_old || !x_new || !y_new || !retarget || !px || !py || !target ||
      !pred_orient || !done || !eaten || ((uintptr_t)tcell & 3))
    return fail(nullptr, SWARM_E_INVALID, "swarm_predator: bad argument");
  PredIO io;
  memset(&io, 0, sizeof(io));
  io.E = E; io.N = N; io.S = S; io.xo = x_old; io.yo = y_old; io.xn = x_new; io.yn = y_new;
  io.retarget = retarget; io.px = px; io.py = py; io.target = target; io.orient = pred_orient;
  io.pred_prev = pred_prev; io.done = done; io.eaten = eaten; io.tcell = tcell;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  launch_predator(io, sms, (cudaStream_t)stream);
  CUDA_TRY(nullptr, cudaGetLastError());
  return SWARM_OK;
}

static void pairwise_scratch_layout(int E, int N, size_t& b_pairw, size_t& b_pc) {
  b_pairw = align_up((size_t)E * ((N + 1) / 2) * 16, 256);
  b_pc = align_up((size_t)E * N * 32 + 64, 256);
}

uint64_t swarm_pairwise_scratch_bytes(int32_t E, int32_t N) {
  if (E < 1 || N < 1) return 0;
  size_t b_pairw, b_pc;
  pairwise_scratch_layout(E, N, b_pairw, b_pc);
  return (uint64_t)(b_pairw + b_pc + 256);
}

int swarm_pairwise(int32_t E, int32_t N, int32_t S, int32_t attraction_thresh, int32_t repulsion_thresh,
                   const int16_t* x, const int16_t* y, const uint8_t* skip, int32_t* out_counts, void* scratch,
                   uint64_t scratch_bytes, void* stream) {
  if (E < 1 || N < 2 || N > SWARM_MAX_AGENTS || S < 2 || S > SWARM_MAX_GRID || !x || !y || !out_counts)
    return fail(nullptr, SWARM_E_INVALID, "swarm_pairwise: bad argument");
  if (!scratch || ((uintptr_t)scratch & 255) || scratch_bytes < swarm_pairwise_scratch_bytes(E, N))
    return fail(nullptr, SWARM_E_INVALID, "swarm_pairwise: scratch must be a 256-byte aligned device block of at least "
                                          "swarm_pairwise_scratch_bytes(E, N) bytes");
  cudaStream_t s = (cudaStream_t)stream;
  StepParams p;
  memset(&p, 0, sizeof(p));
  p.E = E;
  fill_thresholds(p, N, S, attraction_thresh, repulsion_thresh);
  p.t_vf = -1;
  StagedScratch sc;
  memset(&sc, 0, sizeof(sc));
  sc.NW = (N + 1) / 2;
  size_t b_pairw, b_pc;                       // caller-owned scratch: calls with different blocks may overlap freely
  pairwise_scratch_layout(E, N, b_pairw, b_pc);
  sc.pairw = (uint4*)scratch;
  sc.pcounts = (unsigned*)((char*)scratch + b_pairw);
  sc.counter = (unsigned*)((char*)scratch + b_pairw + b_pc);
  const long long words = (long long)E * sc.NW;
  pairprep_kernel<<<(int)((words + 255) / 256), 256, 0, s>>>(E, N, S, sc.NW, x, y, nullptr, sc.pairw, nullptr,
                                                            sc.pcounts, sc.counter, nullptr);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pairwise_launch(p, sc, x, y, skip, false, sms, s);
  const long long agents = (long long)E * N;
  pair_counts_kernel<<<(int)((agents + 255) / 256), 256, 0, s>>>(p, sc.pcounts, skip, out_counts);
  CUDA_TRY(nullptr, cudaGetLastError());
  return SWARM_OK;
}

int swarm_observe_vf(int32_t E, int32_t N, int32_t S, int32_t vf, const int16_t* x, const int16_t* y, const int16_t* px,
                     const int16_t* py, uint8_t* out, void* stream) {
  if (E < 1 || N < 1 || N > OBS_MAX_AGENTS || S < 2 || vf < 0 || vf > 7 || !x || !y || !px || !py || !out)
    return fail(nullptr, SWARM_E_INVALID, "swarm_observe_vf: bad argument (N <= 4096, 0 <= vf <= 7)");
  const int W = 2 * vf + 1;
  const long long cells = (long long)S * S;
  const long long words = (cells + 31) / 32 + 1;
  const bool bitmap = words * 4 <= 160 * 1024;                       // the env's occupancy bitmap fits in shared memory
  const int bm_words = bitmap ? (int)words : 0;
  const int smem = ((N + bm_words) * 4 + 15) / 16 * 16 + OBS_CHUNK * W * W + 16;
  static bool opted[64][2] = {};                                   // per device (the attribute is per function AND device)
  int devi = 0;
  CUDA_TRY(nullptr, cudaGetDevice(&devi));
  if (smem > 48 * 1024 && (devi < 0 || devi >= 64 || !opted[devi][bitmap])) {
    if (bitmap) CUDA_TRY(nullptr, smem_opt_in(observe_vf_kernel<true>, smem));
    else CUDA_TRY(nullptr, smem_opt_in(observe_vf_kernel<false>, smem));
    if (devi >= 0 && devi < 64) opted[devi][bitmap] = true;
  }
  {                                                                // lane-group shapes: a warp per env, no block barriers
    const int apl = N <= 128 ? 4 : 8;
    const long long pwords = ((long long)(S + 2 * vf) * S + 15) / 16 + 2;          // row-padded bitmap, overlapping words
    const int warp_bytes = (int)((pwords + 3) / 4 * 4) * 4 + (N * W * W + 15) / 16 * 16;
    const int wsmem = warp_bytes * OBSW_WARPS;
    if (N <= 256 && (N % apl) == 0 && vf >= 1 && vf <= 5 && pwords * 4 <= 16 * 1024 && wsmem <= 200 * 1024 &&
        ((((uintptr_t)x) | ((uintptr_t)y)) & 7) == 0) {
      static bool wopted[64][8][2] = {};
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devi);
      const int per_sm = std::max(1, std::min(16, (220 * 1024) / (wsmem + 1024)));
      const int wgrid = (int)std::min<long long>(((long long)E + OBSW_WARPS - 1) / OBSW_WARPS, (long long)sms * per_sm);
#define OBSW_LAUNCH(V, A)                                                                                              \
  do {                                                                                                                 \
    if (wsmem > 48 * 1024 && (devi < 0 || devi >= 64 || !wopted[devi][V][A == 8])) {                                   \
      CUDA_TRY(nullptr, smem_opt_in(observe_vf_warp_kernel<V, A>, wsmem));                                             \
      if (devi >= 0 && devi < 64) wopted[devi][V][A == 8] = true;                                                      \
    }                                                                                                                  \
    observe_vf_warp_kernel<V, A><<<wgrid, OBSW_WARPS * 32, wsmem, (cudaStream_t)stream>>>(E, N, S, (int)pwords, warp_bytes, \
                                                                                          x, y, px, py, out);          \
  } while (0)
#define OBSW_CASE(V)                                                                                                   \
  case V:                                                                                                              \
    if (apl == 4) OBSW_LAUNCH(V, 4); else OBSW_LAUNCH(V, 8);                                                           \
    break;
      switch (vf) { OBSW_CASE(1) OBSW_CASE(2) OBSW_CASE(3) OBSW_CASE(4) OBSW_CASE(5) }
#undef OBSW_CASE
#undef OBSW_LAUNCH
      CUDA_TRY(nullptr, cudaGetLastError());
      return SWARM_OK;
    }
  }
  const long long grid = (long long)E * ((N + OBS_CHUNK - 1) / OBS_CHUNK);
  if (grid > 0x7FFFFFFFLL) return fail(nullptr, SWARM_E_INVALID, "swarm_observe_vf: too many agent chunks");
  if (bitmap) observe_vf_kernel<true><<<(int)grid, OBS_CHUNK, smem, (cudaStream_t)stream>>>(E, N, S, vf, bm_words, x, y, px, py, out);
  else observe_vf_kernel<false><<<(int)grid, OBS_CHUNK, smem, (cudaStream_t)stream>>>(E, N, S, vf, 0, x, y, px, py, out);
  CUDA_TRY(nullptr, cudaGetLastError());
  return SWARM_OK;
}

// ---- statistics -----------------------------------------------------------------------------------------

int swarm_stats_begin(swarm_handle h, int32_t allreduce, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return fail(c, SWARM_E_INVALID, "swarm_stats_begin: null handle");
  if (!c->cfg.track_stats || !c->has_state) return fail(c, SWARM_E_STATE, "swarm_stats_begin: track_stats is off");
  if (allreduce && !c->nccl_comm) return fail(c, SWARM_E_STATE, "swarm_stats_begin: call swarm_nccl_init / swarm_nccl_adopt first");
  if (c->stat_begun - c->stat_ended >= (unsigned long long)Ctx::STAT_SLOTS)
    return fail(c, SWARM_E_STATE, "swarm_stats_begin: too many results in flight, collect one with swarm_stats_end");
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int slot = (int)(c->stat_begun % Ctx::STAT_SLOTS);
  double* d = c->d_stats + 8 * slot;
  const int E = c->cfg.num_envs;
  const int per_cta = (E + c->stat_grid - 1) / c->stat_grid;
  // the reduction is ordered with the steps on the caller's stream: it sees exactly the steps enqueued before it
  stats_partial_kernel<<<c->stat_grid, STAT_THREADS, 0, s>>>(E, per_cta, c->st.fin_count, c->st.fin_len, c->st.fin_ret,
                                                            c->st.ep_rec, c->d_partial);
  stats_final_kernel<<<1, STAT_THREADS, 0, s>>>(c->stat_grid, c->d_partial, d);
  CUDA_TRY(c, cudaGetLastError());
  CUDA_TRY(c, cudaEventRecord(c->ev_reduced[slot], s));
  // everything else happens behind that event on the side stream: the step loop never waits for it
  CUDA_TRY(c, cudaStreamWaitEvent(c->side, c->ev_reduced[slot], 0));
  if (allreduce) {
    nccl_allreduce_t f = (nccl_allreduce_t)dlsym(c->nccl_lib, "ncclAllReduce");
    if (!f) return fail(c, SWARM_E_NCCL, "ncclAllReduce missing");
    const int rc = f(d, d, 8, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->side);   // on the device buffer
    if (rc != 0) return fail(c, SWARM_E_NCCL, "ncclAllReduce failed, code " + std::to_string(rc));
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->h_stats + 8 * slot, d, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->side));
  CUDA_TRY(c, cudaEventRecord(c->ev_done[slot], c->side));
  // d_partial is reused by the next begin on the caller's stream: stream order already serialises the two reductions
  c->stat_begun += 1;
  return SWARM_OK;
}

int swarm_stats_end(swarm_handle h, double out[8], int32_t wait) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !out) return fail(c, SWARM_E_INVALID, "swarm_stats_end: null argument");
  if (c->stat_begun == c->stat_ended) return fail(c, SWARM_E_STATE, "swarm_stats_end: no swarm_stats_begin outstanding");
  const int slot = (int)(c->stat_ended % Ctx::STAT_SLOTS);
  if (wait) {
    CUDA_TRY(c, cudaEventSynchronize(c->ev_done[slot]));
  } else {
    const cudaError_t q = cudaEventQuery(c->ev_done[slot]);
    if (q == cudaErrorNotReady) return SWARM_E_NOT_READY;
    if (q != cudaSuccess) return fail(c, SWARM_E_CUDA, std::string("swarm_stats_end: ") + cudaGetErrorString(q));
  }
  memcpy(out, c->h_stats + 8 * slot, 8 * sizeof(double));
  c->stat_ended += 1;
  return SWARM_OK;
}

static int stats_sync(Ctx* c, double out[8], int allreduce, void* stream, const char* who) {
  if (!c || !out) return fail(c, SWARM_E_INVALID, std::string(who) + ": null argument");
  double tmp[8];
  while (c->stat_begun != c->stat_ended) {         // drain older asynchronous results first (kept in order)
    const int rc = swarm_stats_end(reinterpret_cast<swarm_handle>(c), tmp, 1);
    if (rc != SWARM_OK) return rc;
  }
  const int rc = swarm_stats_begin(reinterpret_cast<swarm_handle>(c), allreduce, stream);
  if (rc != SWARM_OK) return rc;
  return swarm_stats_end(reinterpret_cast<swarm_handle>(c), out, 1);
}

int swarm_stats_read(swarm_handle h, double out[8], void* stream) {
  return stats_sync(reinterpret_cast<Ctx*>(h), out, 0, stream, "swarm_stats_read");
}

int swarm_stats_allreduce(swarm_handle h, double out[8], void* stream) {
  return stats_sync(reinterpret_cast<Ctx*>(h), out, 1, stream, "swarm_stats_allreduce");
}

static void* nccl_open() {
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return lib;
}

int swarm_nccl_unique_id(uint8_t out_id[128]) {
  void* lib = nccl_open();
  if (!lib) return fail(nullptr, SWARM_E_NCCL, "libnccl.so.2 not found");
  typedef int (*getid_t)(void*);
  getid_t f = (getid_t)dlsym(lib, "ncclGetUniqueId");
  if (!f) return fail(nullptr, SWARM_E_NCCL, "ncclGetUniqueId missing");
  if (f(out_id) != 0) return fail(nullptr, SWARM_E_NCCL, "ncclGetUniqueId failed");
  return SWARM_OK;
}

int swarm_nccl_init(swarm_handle h, const uint8_t id[128], int32_t nranks, int32_t rank) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !id) return fail(c, SWARM_E_INVALID, "swarm_nccl_init: null argument");
  c->nccl_lib = nccl_open();
  if (!c->nccl_lib) return fail(c, SWARM_E_NCCL, "libnccl.so.2 not found");
  struct uid { char b[128]; } u;
  memcpy(u.b, id, 128);
  typedef int (*init_t)(void**, int, uid, int);
  init_t f = (init_t)dlsym(c->nccl_lib, "ncclCommInitRank");
  if (!f) return fail(c, SWARM_E_NCCL, "ncclCommInitRank missing");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int rc = f(&c->nccl_comm, nranks, u, rank);
  if (rc != 0) return fail(c, SWARM_E_NCCL, "ncclCommInitRank failed, code " + std::to_string(rc));
  c->nccl_owned = true;
  // NCCL connects its channels lazily, inside the first collective (hundreds of milliseconds): pay that here, on the
  // side stream, so that the first swarm_stats_begin of a step loop costs what every later one costs
  nccl_allreduce_t ar = (nccl_allreduce_t)dlsym(c->nccl_lib, "ncclAllReduce");
  if (!ar) return fail(c, SWARM_E_NCCL, "ncclAllReduce missing");
  CUDA_TRY(c, cudaMemsetAsync(c->d_partial, 0, 8 * sizeof(double), c->side));
  const int wrc = ar(c->d_partial, c->d_partial, 8, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->side);
  if (wrc != 0) return fail(c, SWARM_E_NCCL, "ncclAllReduce (warm-up) failed, code " + std::to_string(wrc));
  CUDA_TRY(c, cudaStreamSynchronize(c->side));
  return SWARM_OK;
}

int swarm_nccl_adopt(swarm_handle h, void* comm) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c || !comm) return fail(c, SWARM_E_INVALID, "swarm_nccl_adopt: null argument");
  if (c->nccl_comm) return fail(c, SWARM_E_STATE, "swarm_nccl_adopt: the handle already has a communicator");
  c->nccl_lib = nccl_open();
  if (!c->nccl_lib) return fail(c, SWARM_E_NCCL, "libnccl.so.2 not found");
  c->nccl_comm = comm;
  c->nccl_owned = false;
  return SWARM_OK;
}

// ---- INT32 peak ------------------------------------------------------------------------------------------

int swarm_int32_peak(int32_t device, int32_t kind, double* out_lane_ops_per_s, void* stream) {
  if (!out_lane_ops_per_s || kind < 0 || kind > 4) return fail(nullptr, SWARM_E_INVALID, "swarm_int32_peak: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(nullptr, cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  unsigned* d = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&d, 64));
  const int iters = 4096, grid = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, s);
    switch (kind) {
      case 0: int32_peak_kernel<0><<<grid, threads, 0, s>>>(iters, 12345u + rep, d); break;
      case 1: int32_peak_kernel<1><<<grid, threads, 0, s>>>(iters, 12345u + rep, d); break;
      case 2: int32_peak_kernel<2><<<grid, threads, 0, s>>>(iters, 12345u + rep, d); break;
      case 3: int32_peak_kernel<3><<<grid, threads, 0, s>>>(iters, 12345u + rep, d); break;
      default: int32_peak_kernel<4><<<grid, threads, 0, s>>>(iters, 12345u + rep, d); break;
    }
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaError_t e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail(nullptr, SWARM_E_CUDA, cudaGetErrorString(e));
  const double ops = (double)grid * threads * (double)iters * 8.0 * 2.0;   // CH chains x 2 ops per iteration
  *out_lane_ops_per_s = ops / ((double)best * 1e-3);
  return SWARM_OK;
}

}  // extern "C"
