// Read-bandwidth ceiling of this GPU for read-dominated streaming kernels (context for the predator kernel's roofline):
// LDG.128 grid-stride sums with different unroll depths / CTA shapes, a copy kernel for comparison.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/readbw tools/micro/readbw.cu && build/readbw
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
template <int U>
__global__ void read_kernel(const uint4* __restrict__ a, size_t n, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(a + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n; i += stride) { uint4 v = __ldg(a + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[0] = acc;
}
// two arrays read in lockstep (like xn / yn)
template <int U>
__global__ void read2_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, size_t n, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U], w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { v[u] = __ldg(a + i + u * stride); w[u] = __ldg(b + i + u * stride); }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w ^ w[u].x ^ w[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = __ldg(a + i);
}
__global__ void write_kernel(uint4* __restrict__ b, size_t n, unsigned v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 w = make_uint4(v, v + 1, v + 2, v + 3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = w;
}
// one read stream, W write streams' worth of bytes (read 1 : write R, like the visual-field windows: 1 : 6)
template <int R>
__global__ void expand_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 v = __ldg(a + i);
#pragma unroll
    for (int r = 0; r < R; ++r) b[i * R + r] = make_uint4(v.x + r, v.y, v.z, v.w);
  }
}
template <typename F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
  const size_t bytes = (size_t)1 << 30, n = bytes / 16;
  uint4 *a, *b; unsigned* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int tpb : {128, 256, 512}) for (int cps : {2, 4, 8, 16}) {
    if (tpb * cps > 2048) continue;
    const int grid = sms * cps;
    float t1 = timeit([&] { read_kernel<1><<<grid, tpb>>>(a, n, out); });
    float t4 = timeit([&] { read_kernel<4><<<grid, tpb>>>(a, n, out); });
    float t8 = timeit([&] { read_kernel<8><<<grid, tpb>>>(a, n, out); });
    float t24 = timeit([&] { read2_kernel<4><<<grid, tpb>>>(a, b, n / 2, out); });
    float tc = timeit([&] { copy_kernel<<<grid, tpb>>>(a, b, n); });
    float tw = timeit([&] { write_kernel<<<grid, tpb>>>(b, n, 7u); });
    float te = timeit([&] { expand_kernel<6><<<grid, tpb>>>(a, b, n / 6); });
    printf("tpb %4d cta/sm %2d : write-only %.0f  read1:write6 %.0f GB/s\n", tpb, cps, bytes / tw / 1e6,
           (n / 6) * 16.0 * 7 / te / 1e6);
    printf("tpb %4d cta/sm %2d : read U1 %.0f  U4 %.0f  U8 %.0f  read2 U4 %.0f  copy(r+w) %.0f GB/s\n", tpb, cps,
           bytes / t1 / 1e6, bytes / t4 / 1e6, bytes / t8 / 1e6, bytes / t24 / 1e6, 2.0 * bytes / tc / 1e6);
  }
  { // non-persistent: one vector per thread
    const int grid = (int)(n / 256);
    float t = timeit([&] { read_kernel<1><<<grid, 256>>>(a, n, out); });
    printf("one LDG.128 per thread, %d CTAs: %.0f GB/s\n", grid, bytes / t / 1e6);
  }
  { float t = timeit([&] { cudaMemsetAsync(b, 3, bytes); }); printf("cudaMemsetAsync: %.0f GB/s\n", bytes / t / 1e6); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
