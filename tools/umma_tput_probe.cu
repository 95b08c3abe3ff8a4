// tcgen05.mma issue-to-completion throughput for the operand layouts the edge pipeline uses (no swizzle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I shapemol_b200/csrc tools/umma_tput_probe.cu -o build/umma_tput_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "smb_tc.cuh"
using namespace smb::tc;

// mode 0: GEMM1  SS  A K-major [128][96] (SBO 1536), B MN-major [32 k][128 n] x3 (SBO 512), N = 128, 6 MMAs / tile
// mode 1: GEMM2  TS  A from TMEM, B K-major [128 n][128 k] (SBO 2048), N = 128, 8 MMAs / tile
// mode 2: GEMM2V SS  A K-major W2, B K-major z^T, N = 128, 8 MMAs / tile
// mode 3: GEMM2XV TS N = 16, 8 MMAs / tile
// mode 4: GEMM1 + GEMM2(TS) interleaved as the pipeline issues them
// mode 5: query fold  SS  8 independent MMAs (N = 16, K = 16) into separate column blocks
// mode 6: GEMM2K TS  A from TMEM, B MN-major, N = 64, 8 MMAs / tile
// mode 7: GEMM2K TS  N = 128 B MN-major
// mode 8: GEMM1 + fold + GEMM2K(N = 64) as ROLE_K issues them
__global__ void __launch_bounds__(128, 1) probe(int mode, int tiles, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar, bar2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int p = tid; p < 120 * 1024 / 16; p += 128) reinterpret_cast<uint4*>(smem)[p] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
  if (warp == 0) tmem_alloc<512>(&slot);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init_fence(); }
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t a1 = smem_u32(smem), b1 = a1 + 24576, w2 = b1 + 24576, zt = w2 + 32768;
  constexpr uint32_t ID1 = idesc_bf16(128, true), ID2 = idesc_bf16(128, false), ID3 = idesc_bf16(16, false), ID4 = idesc_bf16(64, true);
  if (tid == 0) {
    const unsigned long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (t % 3) * 128;
      if (mode == 0 || mode == 4)
        for (int ks = 0; ks < 6; ++ks) mma_ss(d, smem_desc(a1 + ks * 256, 128, 1536), smem_desc(b1 + (ks >> 1) * 8192 + (ks & 1) * 256, 128, 512), ID1, ks > 0);
      if (mode == 1 || mode == 4)
        for (int ks = 0; ks < 8; ++ks) mma_ts(tmem + ((t + 1) % 3) * 128, tmem + 384 + (t & 1) * 64 + ks * 8, smem_desc(w2 + ks * 256, 128, 2048), ID2, ks > 0);
      if (mode == 2)
        for (int ks = 0; ks < 8; ++ks) mma_ss(d, smem_desc(w2 + ks * 256, 128, 2048), smem_desc(zt + ks * 256, 128, 2048), ID2, ks > 0);
      if (mode == 3)
        for (int ks = 0; ks < 8; ++ks) mma_ts(d, tmem + 384 + (t & 1) * 64 + ks * 8, smem_desc(w2 + ks * 256, 128, 2048), ID3, ks > 0);
      if (mode == 10)
        for (int ks = 0; ks < 8; ++ks) mma_ss(d, smem_desc(w2 + ks * 256, 128, 2048), smem_desc(zt + ks * 256, 128, 2048), ID2, ks > 0);
      if (mode == 5 || mode == 8 || mode == 9)
        for (int j = 0; j < 8; ++j) mma_ss(tmem + 384 + 16 * j, smem_desc(w2 + j * 256, 128, 2048), smem_desc(zt + j * 512, 128, 256), ID3, 0);
      if (mode == 6 || mode == 8)
        for (int ks = 0; ks < 8; ++ks) mma_ts(tmem + ((t + 1) % 2) * 128, tmem + 256 + (t & 1) * 64 + ks * 8, smem_desc(zt + ks * 256, 128, 2048), ID4, ks > 0);
      if (mode == 7)
        for (int ks = 0; ks < 8; ++ks) mma_ts(tmem + ((t + 1) % 2) * 128, tmem + 256 + (t & 1) * 64 + ks * 8, smem_desc(zt + ks * 256, 128, 2048), ID1, ks > 0);
      if (mode == 8)
        for (int ks = 0; ks < 6; ++ks) mma_ss(tmem + (t % 2) * 128, smem_desc(a1 + ks * 256, 128, 1536), smem_desc(b1 + (ks >> 1) * 8192 + (ks & 1) * 256, 128, 512), ID1, ks > 0);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  if (tid == 32 && (mode == 9 || mode == 10)) {   // a second issuing warp, same work, other accumulator columns
    for (int t = 0; t < tiles; ++t) {
      if (mode == 9)
        for (int j = 0; j < 8; ++j) mma_ss(tmem + 256 + 16 * j, smem_desc(w2 + j * 256, 128, 2048), smem_desc(zt + j * 512, 128, 256), ID3, 0);
      else
        for (int ks = 0; ks < 8; ++ks) mma_ss(tmem + 256, smem_desc(w2 + ks * 256, 128, 2048), smem_desc(zt + ks * 256, 128, 2048), ID2, ks > 0);
    }
    mma_commit(&bar2);
    mbar_wait(&bar2, 0);
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const int tiles = 500;
  const char* names[11] = {"GEMM1 SS (K=96, N=128, B MN-major)", "GEMM2 TS (K=128, N=128)", "GEMM2V SS (K=128, N=128, both K-major)", "GEMM2XV TS (K=128, N=16)", "GEMM1 + GEMM2 TS",
                          "fold SS 8 x (N=16, K=16)", "GEMM2K TS (K=128, N=64, B MN-major)", "GEMM2K TS (K=128, N=128, B MN-major)", "GEMM1 + fold + GEMM2K(N=64)", "fold from TWO issuing warps (per warp)", "GEMM2V N=128 from TWO issuing warps (per warp)"};
  const int mmas[11] = {6, 8, 8, 8, 14, 8, 8, 8, 22, 8, 8};
  for (int mode = 0; mode < 11; ++mode) {
    probe<<<148, 128, 120 * 1024>>>(mode, tiles, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-42s %8.1f cycles/tile  %6.1f cycles/MMA  (%s)\n", names[mode], (double)h[0] / tiles, (double)h[0] / tiles / mmas[mode], cudaGetErrorString(e));
  }
  return 0;
}
