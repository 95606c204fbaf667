// tiny_launch.cu -- thread-per-env step kernels (N <= 8; N <= 16 for large batches) and their masked reset
#include "launchers.h"
#include "tiny_step.cuh"
#include "tiny_packed.cuh"

namespace swarm {

cudaError_t tiny_step_launch(int tiny_n, const StepParams& p, cudaStream_t s, int* launches) {
  const int grid = (p.E + TINY_THREADS - 1) / TINY_THREADS;
  const bool agent_out = p.agent_reward || p.agent_counts || p.agent_terms;
  if (tiny_n == TINY_N && p.N == TINY_N && !agent_out && !p.tiny_scalar) {
    tiny8_packed_kernel<<<grid, TINY_THREADS, 0, s>>>(p);   // exactly 8 agents, env totals only: positions stay packed
  } else if (tiny_n == TINY_N) {
    if (agent_out) tiny_step_kernel<TINY_N, true><<<grid, TINY_THREADS, 0, s>>>(p);
    else tiny_step_kernel<TINY_N, false><<<grid, TINY_THREADS, 0, s>>>(p);
  } else {
    if (agent_out) tiny_step_kernel<TINY_N_BIG, true><<<grid, TINY_THREADS, 0, s>>>(p);
    else tiny_step_kernel<TINY_N_BIG, false><<<grid, TINY_THREADS, 0, s>>>(p);
  }
  *launches = 1;
  if (p.auto_reset) {
    const int rgrid = (p.E + TINY_RESET_SPAN - 1) / TINY_RESET_SPAN;
    if (tiny_n == TINY_N) tiny_reset_kernel<TINY_N><<<rgrid, TINY_THREADS, 0, s>>>(p, p.done);
    else tiny_reset_kernel<TINY_N_BIG><<<rgrid, TINY_THREADS, 0, s>>>(p, p.done);
    *launches = 2;
  }
  return cudaGetLastError();
}

}  // namespace swarm
