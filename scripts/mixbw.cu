// DRAM throughput vs read:write mix (tuning experiment, not part of the product).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// each thread: R loads from R planes, W stores to W planes, same index (plane-strided streams)
template <int R, int W>
__global__ void k_planes(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4 v = make_uint4(1, 2, 3, (uint32_t)i);
#pragma unroll
  for (int r = 0; r < R; ++r) { uint4 t = a[(size_t)r * n + i]; v.x ^= t.x; v.y += t.y; v.z ^= t.z; v.w += t.w; }
#pragma unroll
  for (int w = 0; w < W; ++w) b[(size_t)w * n + i] = v;
}
// contiguous output: each warp writes W*512 contiguous bytes (like a 32-env chunk), reads R*512 contiguous
template <int R, int W>
__global__ void k_chunks(const uint4* __restrict__ a, uint4* __restrict__ b, size_t nwarps) {
  size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  uint4 v = make_uint4(1, 2, 3, (uint32_t)warp);
#pragma unroll
  for (int r = 0; r < R; ++r) { uint4 t = a[(warp * R + r) * 32 + lane]; v.x ^= t.x; v.y += t.y; v.z ^= t.z; v.w += t.w; }
#pragma unroll
  for (int w = 0; w < W; ++w) b[(warp * W + w) * 32 + lane] = v;
}
template <typename F> float timeit(F f, int reps = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
template <int R, int W> void run(const uint4* a, uint4* b, size_t n) {
  float ms = timeit([&] { k_planes<R, W><<<(unsigned)((n + 255) / 256), 256>>>(a, b, n); });
  double bytes = (double)(R + W) * n * 16;
  printf("planes %d r : %d w   %8.1f us  total %6.0f GB/s  (read %5.0f, write %5.0f)\n", R, W, ms * 1e3, bytes / ms / 1e6,
         R * n * 16.0 / ms / 1e6, W * n * 16.0 / ms / 1e6);
  size_t nwarps = n / 32;
  ms = timeit([&] { k_chunks<R, W><<<(unsigned)((n + 255) / 256), 256>>>(a, b, nwarps); });
  printf("chunks %d r : %d w   %8.1f us  total %6.0f GB/s  (read %5.0f, write %5.0f)\n", R, W, ms * 1e3, bytes / ms / 1e6,
         R * n * 16.0 / ms / 1e6, W * n * 16.0 / ms / 1e6);
}
int main() {
  size_t n = (size_t)1 << 22;           // 4M threads; each plane 64 MB
  uint4 *a, *b;
  CK(cudaMalloc(&a, n * 16 * 8)); CK(cudaMalloc(&b, n * 16 * 16));
  CK(cudaMemset(a, 1, n * 16 * 8)); CK(cudaMemset(b, 2, n * 16 * 16));
  run<1, 0>(a, b, n); run<4, 0>(a, b, n); run<8, 0>(a, b, n);
  run<0, 1>(a, b, n); run<0, 4>(a, b, n); run<0, 8>(a, b, n); run<0, 16>(a, b, n);
  run<4, 4>(a, b, n); run<2, 4>(a, b, n); run<1, 4>(a, b, n); run<1, 6>(a, b, n); run<1, 8>(a, b, n); run<2, 12>(a, b, n);
  run<1, 12>(a, b, n); run<1, 16>(a, b, n); run<4, 2>(a, b, n); run<8, 1>(a, b, n);
  return 0;
}
