// Shared device helpers for the shapemol_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define SMB_WARP 32

#define SMB_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) { smb::set_error(#expr, _e); return (int)_e; }            \
  } while (0)

// Timing-ablation switches (EdgeArgs::dbg / NodeArgs::dbg bit masks, results are wrong when set) and the role trace exist
// only in -DSMB_DEBUG builds (SMB_NVCC_EXTRA=-DSMB_DEBUG python -m shapemol_b200.build --force); release builds compile
// them out.
#ifdef SMB_DEBUG
#define SMB_DBG(args, bit) (((args).dbg & (bit)) != 0)
#else
#define SMB_DBG(args, bit) false
#endif

namespace smb {

void set_error(const char* what, cudaError_t e);
void set_error_msg(const char* msg);

// ---- per-device launch configuration ----------------------------------------------------
// cudaFuncSetAttribute and the SM count belong to the CURRENT device, so both are cached per device ordinal
// (idempotent values: a racing first use from two host threads is benign).
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
template <class Kernel>
inline int ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&cache)[kMaxDevices]) {
  const int d = current_device_slot();
  if (cache[d] < bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    cache[d] = bytes;
  }
  return 0;
}
inline int device_sm_count() {
  static int cache[kMaxDevices] = {};
  const int d = current_device_slot();
  if (cache[d] == 0) {
    int n = 0;
    cache[d] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) == cudaSuccess && n > 0) ? n : 148;
  }
  return cache[d];
}

// ---- bf16 split helpers -------------------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significant bits ("bf16x3" products
// hi*hi + lo*hi + hi*lo reproduce an fp32 product to ~2^-16 relative).
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b (high half)
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  float ar = a - __bfloat162float(ah), br = b - __bfloat162float(bh);
  __nv_bfloat162 h; h.x = ah; h.y = bh;
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_bf16x2(ar, br);
}

// ---- mma.sync m16n8k16 bf16 (fp32 accumulate) ---------------------------------------
// A (16x16, row): a0=(g,2t..2t+1) a1=(g+8,2t..) a2=(g,2t+8..) a3=(g+8,2t+8..)
// B (16x8,  col): b0=(k=2t..2t+1, n=g) b1=(k=2t+8..2t+9, n=g)
// C (16x8):       c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)      g=lane>>2, t=lane&3
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  // not volatile: the scheduler may interleave independent accumulator chains
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One K=16 step of C += A*B at the selected precision.  bw = {hi01, hi23, lo01, lo23}.
template <bool X3>
__device__ __forceinline__ void mma_step(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                         const uint4& bw) {
  if (X3) {
    mma_bf16(c, alo, bw.x, bw.y);
    mma_bf16(c, ahi, bw.z, bw.w);
  }
  mma_bf16(c, ahi, bw.x, bw.y);
}

// One K=16 step for NC independent accumulators sharing the A operand (one B fragment each); the
// three split terms are issued term-major so that consecutive MMAs never depend on each other.
template <bool X3, int NC>
__device__ __forceinline__ void mma_step_n(float (&c)[NC][4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           const uint4 (&bw)[NC]) {
  if (X3) {
#pragma unroll
    for (int q = 0; q < NC; ++q) mma_bf16(c[q], alo, bw[q].x, bw[q].y);
#pragma unroll
    for (int q = 0; q < NC; ++q) mma_bf16(c[q], ahi, bw[q].z, bw[q].w);
  }
#pragma unroll
  for (int q = 0; q < NC; ++q) mma_bf16(c[q], ahi, bw[q].x, bw[q].y);
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float group_sum(float v) {   // across the 8 row-groups (lane>>2)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cp.async 16-byte copy global -> shared
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// RBF centres of GaussianSmearing (reference models/common.py:19); coeff = -0.5.
__device__ __forceinline__ float rbf_centre(int g) {
  // 0, 1, 1.25 .. 3 (step .25), 3.5 .. 6 (step .5), 7, 8, 9, 10
  if (g == 0) return 0.f;
  if (g <= 9) return 1.f + 0.25f * (float)(g - 1);
  if (g <= 15) return 3.f + 0.5f * (float)(g - 9);
  return 6.f + (float)(g - 15);
}

}  // namespace smb
