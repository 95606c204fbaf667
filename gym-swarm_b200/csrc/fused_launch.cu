// fused_launch.cu -- lane-group step kernels without vf_size, and the lane-group reset kernel
#include "fused_dispatch.cuh"

namespace swarm {

void choose_fused_shape(int N, int& lpe, int& apl) {
  if (N <= 4) { lpe = 4; apl = 1; }
  else if (N <= 8) { lpe = 8; apl = 1; }
  else if (N <= 16) { lpe = 8; apl = 2; }
  else if (N <= 32) { lpe = 8; apl = 4; }
  else if (N <= 64) { lpe = 16; apl = 4; }
  else if (N <= 128) { lpe = 32; apl = 4; }
  else { lpe = 32; apl = 8; }
}

template <int LPE, int APL>
static cudaError_t prepare(const StepParams& p) {
  cudaError_t e = smem_opt_in(fused_step_kernel<LPE, APL, 0, 0>, fused_smem_bytes<LPE, APL>(p, 0));
  if (e == cudaSuccess) e = smem_opt_in(fused_step_kernel<LPE, APL, 0, 1>, fused_smem_bytes<LPE, APL>(p, 0));
  if (e == cudaSuccess) e = smem_opt_in(fused_step_kernel<LPE, APL, 0, 2>, fused_smem_bytes<LPE, APL>(p, 0));
  return e;
}

template <int LPE, int APL>
static cudaError_t launch_step(const StepParams& p, cudaStream_t s) {
  using C = FusedCfg<LPE, APL>;
  const int envs_per_cta = C::WARPS * C::EPW;
  const int grid = (p.E + envs_per_cta - 1) / envs_per_cta;
  const int smem = fused_smem_bytes<LPE, APL>(p, 0);
  if (p.split) {
    fused_step_kernel<LPE, APL, 0, 1><<<grid, C::WARPS * 32, smem, s>>>(p);
    fused_step_kernel<LPE, APL, 0, 2><<<grid, C::WARPS * 32, smem, s>>>(p);
  } else {
    fused_step_kernel<LPE, APL, 0, 0><<<grid, C::WARPS * 32, smem, s>>>(p);
  }
  return cudaGetLastError();
}

template <int LPE, int APL>
static cudaError_t launch_transition(const StepParams& p, cudaStream_t s) {
  using C = FusedCfg<LPE, APL>;
  const int envs_per_cta = C::WARPS * C::EPW;
  const int grid = (p.E + envs_per_cta - 1) / envs_per_cta;
  fused_step_kernel<LPE, APL, 0, 1><<<grid, C::WARPS * 32, fused_smem_bytes<LPE, APL>(p, 0), s>>>(p);
  return cudaGetLastError();
}

template <int LPE, int APL>
static cudaError_t launch_reset(const StepParams& p, const uint8_t* mask, int zero_counts, cudaStream_t s) {
  using C = FusedCfg<LPE, APL>;
  const int envs_per_cta = C::WARPS * C::EPW;
  const int grid = (p.E + envs_per_cta - 1) / envs_per_cta;
  const int smem = C::WARPS * C::EPW * C::NMAX * 4;
  fused_reset_kernel<LPE, APL><<<grid, C::WARPS * 32, smem, s>>>(p, mask, zero_counts);
  return cudaGetLastError();
}

cudaError_t fused_prepare(int lpe, int apl, const StepParams& p) {
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = prepare<L, A>(p)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e != cudaSuccess ? e : fused_vf_prepare(lpe, apl, p);
}

cudaError_t fused_step_launch(int lpe, int apl, const StepParams& p, cudaStream_t s) {
  if (p.t_vf >= 0) return fused_vf_step_launch(lpe, apl, p, s);
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = launch_step<L, A>(p, s)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e;
}

cudaError_t fused_transition_launch(int lpe, int apl, const StepParams& p, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = launch_transition<L, A>(p, s)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e;
}

cudaError_t fused_reset_launch(int lpe, int apl, const StepParams& p, const uint8_t* mask, int zero_counts, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = launch_reset<L, A>(p, mask, zero_counts, s)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e;
}

}  // namespace swarm
