// fused_dispatch.cuh -- shape dispatch and shared-memory sizing shared by fused_launch.cu and fused_vf_launch.cu
#pragma once
#include "fused_step.cuh"
#include "launchers.h"

namespace swarm {

template <int LPE, int APL>
static int fused_smem_bytes(const StepParams& p, int vfm) {
  using C = FusedCfg<LPE, APL>;
  return C::WARPS * (C::pair_bytes() + C::vf_bytes(vfm) + C::far_bytes() + C::EPW * p.bitmap_words * 4);
}

#define FUSED_DISPATCH(_l, _a, CALL)                             \
  do {                                                           \
    if (_l == 4 && _a == 1) { CALL(4, 1); }                      \
    else if (_l == 8 && _a == 1) { CALL(8, 1); }                 \
    else if (_l == 8 && _a == 2) { CALL(8, 2); }                 \
    else if (_l == 8 && _a == 4) { CALL(8, 4); }                 \
    else if (_l == 16 && _a == 4) { CALL(16, 4); }               \
    else if (_l == 32 && _a == 4) { CALL(32, 4); }               \
    else { CALL(32, 8); }                                        \
  } while (0)

}  // namespace swarm
