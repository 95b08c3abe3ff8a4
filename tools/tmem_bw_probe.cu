// Measures tcgen05.ld (TMEM -> registers) throughput per SM: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I shapemol_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "smb_tc.cuh"
using namespace smb::tc;

template <int X>
__global__ void probe(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) {
      uint32_t v[32];
      tmem_ld32(lane_addr + ((i * 32 + (warp >> 2) * 64) & 511), v);
      wait_ld();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= v[q];
    } else {
      uint32_t v[32], w[32];
      tmem_ld32(lane_addr + ((i * 64 + (warp >> 2) * 128) & 511), v);
      tmem_ld32(lane_addr + ((i * 64 + 32 + (warp >> 2) * 128) & 511), w);
      wait_ld();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= v[q] ^ w[q];
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

int main() {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int x : {32, 64}) {
      if (x == 32) probe<32><<<148, warps * 32>>>(iters, out, sink); else probe<64><<<148, warps * 32>>>(iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * warps * 32 * x * 4;
      printf("warps %2d  x%d per wait: %llu cycles, %.1f B/cycle/SM  (%s)\n", warps, x, h[0], bytes / (double)h[0], cudaGetErrorString(e));
    }
  }
  return 0;
}
