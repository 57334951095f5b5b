// Write-bandwidth microbenchmark (tuning experiment, not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw wbw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_fill_v4(uint4* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(1, 2, 3, 4);
  for (; i < n; i += st) p[i] = v;
}
__global__ void k_fill_cs(uint4* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p + i), "r"(7u) : "memory");
}
__global__ void k_fill_wt(uint4* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) asm volatile("st.global.wt.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p + i), "r"(7u) : "memory");
}
__global__ void k_fill_v8(uint4* p, size_t n) {   // 256-bit stores, n counts uint4
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (size_t)gridDim.x * blockDim.x * 2;
  for (; i < n; i += st) asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(p + i), "r"(7u) : "memory");
}
// contiguous chunk per block (like one warp = one contiguous slice), each block walks its own range
__global__ void k_fill_chunk(uint4* p, size_t n) {
  size_t per = (n + gridDim.x - 1) / gridDim.x;
  size_t lo = (size_t)blockIdx.x * per, hi = lo + per; if (hi > n) hi = n;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) p[i] = make_uint4(1, 2, 3, 4);
}
// TMA-style bulk stores from shared memory
__global__ void k_fill_bulk(uint4* p, size_t n, int chunk_u4) {
  extern __shared__ __align__(128) uint4 sm[];
  for (int i = threadIdx.x; i < chunk_u4; i += blockDim.x) sm[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    size_t nchunks = n / chunk_u4;
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
    int inflight = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + c * chunk_u4), "r"(saddr), "r"(chunk_u4 * 16) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// non-persistent TMA: one CTA = one (or a few) bulk stores, then exit
__global__ void k_fill_bulk_once(uint4* p, size_t n, int chunk_u4, int per_cta) {
  extern __shared__ __align__(128) uint4 sm[];
  for (int i = threadIdx.x; i < chunk_u4; i += blockDim.x) sm[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
    for (int k = 0; k < per_cta; ++k) {
      size_t c = (size_t)blockIdx.x * per_cta + k;
      if (c * chunk_u4 < n)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + c * chunk_u4), "r"(saddr), "r"(chunk_u4 * 16) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
// per-warp TMA: every warp of a persistent CTA owns a staging buffer and issues its own bulk stores
__global__ void k_fill_bulk_warp(uint4* p, size_t n, int chunk_u4) {
  extern __shared__ __align__(128) uint4 sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint4* my = sm + (size_t)warp * chunk_u4;
  for (int i = lane; i < chunk_u4; i += 32) my[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  size_t nchunks = n / chunk_u4;
  uint32_t saddr = (uint32_t)__cvta_generic_to_shared(my);
  for (size_t c = (size_t)blockIdx.x * nw + warp; c < nchunks; c += (size_t)gridDim.x * nw) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + c * chunk_u4), "r"(saddr), "r"(chunk_u4 * 16) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// persistent per-warp TMA with a dynamic (atomic) chunk queue: keeps the write front tight
__global__ void k_fill_bulk_dyn(uint4* p, size_t n, int chunk_u4, unsigned long long* counter, int grab) {
  extern __shared__ __align__(128) uint4 sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* my = sm + (size_t)warp * chunk_u4;
  for (int i = lane; i < chunk_u4; i += 32) my[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  size_t nchunks = n / chunk_u4;
  uint32_t saddr = (uint32_t)__cvta_generic_to_shared(my);
  while (true) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(counter, (unsigned long long)grab);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    if (c0 >= nchunks) break;
    for (int k = 0; k < grab && c0 + k < nchunks; ++k) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + (c0 + k) * chunk_u4), "r"(saddr), "r"(chunk_u4 * 16) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// persistent, static but BLOCKED assignment: warp w owns a contiguous range of chunks
__global__ void k_fill_bulk_blocked(uint4* p, size_t n, int chunk_u4) {
  extern __shared__ __align__(128) uint4 sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint4* my = sm + (size_t)warp * chunk_u4;
  for (int i = lane; i < chunk_u4; i += 32) my[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  size_t nchunks = n / chunk_u4, tw = (size_t)gridDim.x * nw, gw = (size_t)blockIdx.x * nw + warp;
  size_t per = (nchunks + tw - 1) / tw, lo = gw * per, hi = lo + per; if (hi > nchunks) hi = nchunks;
  uint32_t saddr = (uint32_t)__cvta_generic_to_shared(my);
  for (size_t c = lo; c < hi; ++c) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + c * chunk_u4), "r"(saddr), "r"(chunk_u4 * 16) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void k_copy(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
__global__ void k_read(const uint4* __restrict__ a, size_t n, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n; i += st) { uint4 v = a[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) *out = acc;
}
// 1 read : 6 writes, like the step kernel
__global__ void k_mix(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n / 6; i += st) { uint4 v = a[i]; for (int k = 0; k < 6; ++k) b[i + k * (n / 6)] = v; }
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

int main() {
  size_t bytes = (size_t)1 << 30, n = bytes / 16;
  uint4 *a, *b; uint32_t* out;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  float ms;
  ms = timeit([&] { cudaMemsetAsync(b, 3, bytes); }); printf("cudaMemset                      %8.1f us %7.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
  int grids[] = {148 * 2, 148 * 8, 148 * 32, 65536};
  for (int g : grids) {
    ms = timeit([&] { k_fill_v4<<<g, 256>>>(b, n); }); printf("fill v4 default   grid %6d    %8.1f us %7.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { k_fill_cs<<<g, 256>>>(b, n); }); printf("fill v4 .cs       grid %6d    %8.1f us %7.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { k_fill_wt<<<g, 256>>>(b, n); }); printf("fill v4 .wt       grid %6d    %8.1f us %7.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { k_fill_v8<<<g, 256>>>(b, n); }); printf("fill v8 (256-bit) grid %6d    %8.1f us %7.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { k_fill_chunk<<<g, 256>>>(b, n); }); printf("fill chunked      grid %6d    %8.1f us %7.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
  }
  int chunks[] = {256, 1024, 2048};   // uint4 per bulk copy: 4 KB, 16 KB, 32 KB
  for (int c : chunks)
    for (int g : {148, 148 * 2, 148 * 4}) {
      cudaFuncSetAttribute(k_fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, c * 16);
      ms = timeit([&] { k_fill_bulk<<<g, 128, c * 16>>>(b, n, c); });
      printf("fill bulk (TMA) %5d B grid %4d %8.1f us %7.0f GB/s\n", c * 16, g, ms * 1e3, bytes / ms / 1e6);
    }
  for (int c : {768, 1024}) for (int per : {1, 4}) {
    cudaFuncSetAttribute(k_fill_bulk_once, cudaFuncAttributeMaxDynamicSharedMemorySize, c * 16);
    unsigned g = (unsigned)((n / c + per - 1) / per);
    ms = timeit([&] { k_fill_bulk_once<<<g, 128, c * 16>>>(b, n, c, per); });
    printf("fill bulk once %5d B x%d grid %6u %8.1f us %7.0f GB/s\n", c * 16, per, g, ms * 1e3, bytes / ms / 1e6);
  }
  for (int c : {768, 1024}) for (int g : {148 * 3, 148 * 2}) {
    cudaFuncSetAttribute(k_fill_bulk_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, c * 16 * 4);
    ms = timeit([&] { k_fill_bulk_warp<<<g, 128, c * 16 * 4>>>(b, n, c); });
    printf("fill bulk per-warp %5d B grid %4d %8.1f us %7.0f GB/s\n", c * 16, g, ms * 1e3, bytes / ms / 1e6);
  }
  unsigned long long* counter; CK(cudaMalloc(&counter, 8));
  for (int c : {768}) for (int g : {148 * 3}) for (int grab : {1, 4, 16}) {
    cudaFuncSetAttribute(k_fill_bulk_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, c * 16 * 4);
    ms = timeit([&] { cudaMemsetAsync(counter, 0, 8); k_fill_bulk_dyn<<<g, 128, c * 16 * 4>>>(b, n, c, counter, grab); });
    printf("fill bulk dynamic %5d B grid %4d grab %2d %8.1f us %7.0f GB/s\n", c * 16, g, grab, ms * 1e3, bytes / ms / 1e6);
  }
  for (int c : {768}) for (int g : {148 * 3}) {
    cudaFuncSetAttribute(k_fill_bulk_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, c * 16 * 4);
    ms = timeit([&] { k_fill_bulk_blocked<<<g, 128, c * 16 * 4>>>(b, n, c); });
    printf("fill bulk blocked %5d B grid %4d %8.1f us %7.0f GB/s\n", c * 16, g, ms * 1e3, bytes / ms / 1e6);
  }
  ms = timeit([&] { k_copy<<<148 * 16, 256>>>(a, b, n); }); printf("copy v4                         %8.1f us %7.0f GB/s (r+w)\n", ms * 1e3, 2.0 * bytes / ms / 1e6);
  ms = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D                  %8.1f us %7.0f GB/s (r+w)\n", ms * 1e3, 2.0 * bytes / ms / 1e6);
  ms = timeit([&] { k_read<<<148 * 16, 256>>>(a, n, out); }); printf("read v4                         %8.1f us %7.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
  ms = timeit([&] { k_mix<<<148 * 16, 256>>>(a, b, n); }); printf("mix 1r:6w                       %8.1f us %7.0f GB/s (r+w)\n", ms * 1e3, (bytes + bytes / 6.0) / ms / 1e6);
  return 0;
}
