// launchers.h -- host entry points of the kernel translation units.  Every kernel family lives in its own .cu file
// (fused_launch.cu, fused_vf_launch.cu, tiny_launch.cu; the staged kernels stay with the ABI in swarm_abi.cu) so that
// the library builds in parallel; no device code is shared across translation units except through headers.
#pragma once
#include "swarm_common.cuh"

namespace swarm {

// Opt-in to dynamic shared memory: cudaFuncAttributeMaxDynamicSharedMemorySize is a per-function, per-device attribute
// shared by every handle, so it is only ever raised -- to the most the device allows this kernel (opt-in maximum minus
// the kernel's static shared memory) -- never set to one handle's size: a second, smaller handle would otherwise lower
// it under the first one's launches.
template <typename K>
static cudaError_t smem_opt_in(K kernel, int needed) {
  int dev = 0, dev_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  const int limit = dev_max - (int)fa.sharedSizeBytes;
  if (needed > limit) return cudaErrorInvalidConfiguration;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
}

constexpr int TINY_N = 8;        // largest N of the default thread-per-env kernel
constexpr int TINY_N_BIG = 16;   // second instantiation (9..16 agents), used for large batches only

// lane-group shape <LPE, APL> for N agents (N <= SWARM_FUSED_MAX_AGENTS)
void choose_fused_shape(int N, int& lpe, int& apl);

// fused_launch.cu: steps without vf_size and the (masked) reset.  fused_prepare opts in to the dynamic shared memory
// of every step kernel of the shape (both files) once per handle, so that swarm_step itself only launches.
cudaError_t fused_prepare(int lpe, int apl, const StepParams& p);
cudaError_t fused_step_launch(int lpe, int apl, const StepParams& p, cudaStream_t s);   // forwards vf_size steps
cudaError_t fused_transition_launch(int lpe, int apl, const StepParams& p, cudaStream_t s);   // first of two launches (p.split)
cudaError_t fused_reset_launch(int lpe, int apl, const StepParams& p, const uint8_t* mask, int zero_counts, cudaStream_t s);

// fused_vf_launch.cu: vf_size steps (dense and sparse visual-field modes)
cudaError_t fused_vf_prepare(int lpe, int apl, const StepParams& p);
cudaError_t fused_vf_step_launch(int lpe, int apl, const StepParams& p, cudaStream_t s);

// tiny_launch.cu: thread-per-env step (+ masked auto-reset launch); returns the number of launches in *launches
cudaError_t tiny_step_launch(int tiny_n, const StepParams& p, cudaStream_t s, int* launches);

}  // namespace swarm
