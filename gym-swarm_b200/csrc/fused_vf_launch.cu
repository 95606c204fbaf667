// fused_vf_launch.cu -- lane-group step kernels for vf_size steps (SDE:349-355): dense and sparse visual-field modes
#include "fused_dispatch.cuh"

namespace swarm {

// 1 = dense (every ordered pair evaluated with the gate), 2 = sparse (the symmetric ring sweep records the visible
// pairs, alignment sums from those).  Sparse pays while few pairs are visible: expected fraction
// ((2 vf + 1) / S)^2 <= 1/16; measured at 65,536 x 128 on 128^2, vf = 5: DESIGN.md 5.2.
template <int LPE, int APL>
static int vf_mode(const StepParams& p) {
  if (!FusedCfg<LPE, APL>::ring_ok()) return 1;
  if (p.vf_dense) return 1;                                                // SWARM_FLAG_VF_DENSE (tests / A-B runs)
  const long long w = 2LL * (p.t_vf - 1) + 1;
  return (w * w * 16 <= (long long)p.S * p.S) ? 2 : 1;
}

template <int LPE, int APL>
static cudaError_t prepare(const StepParams& p) {
  cudaError_t e = smem_opt_in(fused_step_kernel<LPE, APL, 1, 0>, fused_smem_bytes<LPE, APL>(p, 1));
  if (e == cudaSuccess) e = smem_opt_in(fused_step_kernel<LPE, APL, 1, 2>, fused_smem_bytes<LPE, APL>(p, 1));
  if constexpr (FusedCfg<LPE, APL>::ring_ok()) {
    if (e == cudaSuccess) e = smem_opt_in(fused_step_kernel<LPE, APL, 2, 0>, fused_smem_bytes<LPE, APL>(p, 2));
    if (e == cudaSuccess) e = smem_opt_in(fused_step_kernel<LPE, APL, 2, 2>, fused_smem_bytes<LPE, APL>(p, 2));
  }
  return e;
}

template <int LPE, int APL>
static cudaError_t launch_step(const StepParams& p, cudaStream_t s) {
  using C = FusedCfg<LPE, APL>;
  const int envs_per_cta = C::WARPS * C::EPW;
  const int grid = (p.E + envs_per_cta - 1) / envs_per_cta;
  const int vfm = vf_mode<LPE, APL>(p);
  const int smem = fused_smem_bytes<LPE, APL>(p, vfm);
  if (p.split) {                                   // transition launch (no visual-field code in it), then the reward
    const cudaError_t e = fused_transition_launch(LPE, APL, p, s);
    if (e != cudaSuccess) return e;
    if (vfm == 1) fused_step_kernel<LPE, APL, 1, 2><<<grid, C::WARPS * 32, smem, s>>>(p);
    else if constexpr (C::ring_ok()) fused_step_kernel<LPE, APL, 2, 2><<<grid, C::WARPS * 32, smem, s>>>(p);
    return cudaGetLastError();
  }
  if (vfm == 1) fused_step_kernel<LPE, APL, 1, 0><<<grid, C::WARPS * 32, smem, s>>>(p);
  else if constexpr (C::ring_ok()) fused_step_kernel<LPE, APL, 2, 0><<<grid, C::WARPS * 32, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t fused_vf_prepare(int lpe, int apl, const StepParams& p) {
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = prepare<L, A>(p)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e;
}

cudaError_t fused_vf_step_launch(int lpe, int apl, const StepParams& p, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
#define CALL(L, A) e = launch_step<L, A>(p, s)
  FUSED_DISPATCH(lpe, apl, CALL);
#undef CALL
  return e;
}

}  // namespace swarm
