// tcgen05 / TMEM / mbarrier building blocks for the sm_100a kernels (inline PTX; no CUTLASS).
//
// Shared-memory operand layouts used throughout (no swizzle; validated on hardware by
// tools/umma_probe.cu):
//   K-major  [rows][K] bf16 : byte(r, k) = (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2      (LBO = 128)
//   MN-major [K][cols] bf16 : byte(k, c) = (c/8)*SBO + k*16 + (c%8)*2                       (LBO = 128)
//     -> along K the MN-major layout is linear with a 16-byte stride, so a K=16 step may start at any k.
// A-from-TMEM: lane = row, 32-bit column j holds (k=2j, k=2j+1), k even in the low half.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version 1 (sm_100)
  return d;          // base offset 0, no swizzle
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M = 128, A K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- MMA issue (one thread) -----------------------------------------------------------------
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ---- fences -----------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM allocation (one warp) -----------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "power of two >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(NCOLS) : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
// A failed try_wait returns after a few cycles, so a bare retry loop re-issues twice per ~8 cycles and takes issue slots
// from the working warps of its scheduler; with a fixed 20 ns back-off the polling (try_wait + nanosleep + branch) was still a
// third of all executed instructions of the edge pipeline (ncu source page, round 2).  Exponential back-off: short waits stay
// responsive, a warp that waits a whole pipeline stage polls a handful of times.  MAX_NS bounds the wake-up latency a waiter adds
// to the hand-off (the single MMA-issuer warps use a small cap: their latency is on every tile's critical path).
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
template <uint32_t MAX_NS = 256>
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  if (mbar_try(b, parity)) return;
  uint32_t ns = 32;
#pragma unroll 1
  while (true) {
    asm volatile("nanosleep.u32 %0;" :: "r"(ns));
    if (mbar_try(b, parity)) return;
    ns = ns * 2 < MAX_NS ? ns * 2 : MAX_NS;
  }
}
// busy-polling variant for the single MMA-issuer warps: their wake-up latency is on every tile's critical path
__device__ __forceinline__ void mbar_wait_spin(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\t"
               "WAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@!P1 bra WAIT_%=;\n\t"
               "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine); completion is signalled on the mbarrier as `bytes` of transaction count
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {   // adds pending transaction bytes, no arrival
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
// named barrier over `n` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// ---- TMEM <-> registers, 32x32b shape: thread t of warp w <-> lane 32*(w%4)+t, consecutive columns ------
#define SMB_R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define SMB_W4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : SMB_R4(v, 0), SMB_R4(v, 4), SMB_R4(v, 8), SMB_R4(v, 12), SMB_R4(v, 16), SMB_R4(v, 20), SMB_R4(v, 24), SMB_R4(v, 28)
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : SMB_R4(v, 0), SMB_R4(v, 4), SMB_R4(v, 8), SMB_R4(v, 12) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : SMB_R4(v, 0), SMB_R4(v, 4) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : SMB_R4(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t (&v)[1]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               :: "r"(taddr), SMB_W4(v, 0), SMB_W4(v, 4), SMB_W4(v, 8), SMB_W4(v, 12), SMB_W4(v, 16), SMB_W4(v, 20), SMB_W4(v, 24), SMB_W4(v, 28)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), SMB_W4(v, 0), SMB_W4(v, 4), SMB_W4(v, 8), SMB_W4(v, 12) : "memory");
}
#undef SMB_R4
#undef SMB_W4

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two fp32 operations per issued instruction) ---------------
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float sum2(uint64_t v) { float lo, hi; upk2(v, lo, hi); return lo + hi; }

// LayerNorm pieces shared by the tcgen05 kernels (affine part folded into the operands, smb_host.cu):
// sum of squares of 64 accumulator columns, and  z = relu(v * rstd + beta') packed to bf16 pairs.
__device__ __forceinline__ float ln_sumsq64(const uint32_t (&v)[64]) {
  uint64_t q0 = 0ull, q1 = 0ull, q2 = 0ull, q3 = 0ull;
#pragma unroll
  for (int e = 0; e < 64; e += 8) {
    const uint64_t p0 = pk2u(v[e], v[e + 1]), p1 = pk2u(v[e + 2], v[e + 3]), p2 = pk2u(v[e + 4], v[e + 5]), p3 = pk2u(v[e + 6], v[e + 7]);
    q0 = fma2(p0, p0, q0); q1 = fma2(p1, p1, q1); q2 = fma2(p2, p2, q2); q3 = fma2(p3, p3, q3);
  }
  return (sum2(q0) + sum2(q1)) + (sum2(q2) + sum2(q3));
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b);
// 32 columns starting at v[c0] -> 16 packed words; beta points at the 32 floats of these columns (16-byte aligned, smem)
__device__ __forceinline__ void ln_apply32(const uint32_t (&v)[64], int c0, float rstd, const float* beta, uint32_t (&zp)[16]) {
  const uint64_t r2 = pk2(rstd, rstd);
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 bb = *reinterpret_cast<const float4*>(beta + e);
    float y0, y1, y2, y3;
    upk2(fma2(pk2u(v[c0 + e], v[c0 + e + 1]), r2, pk2(bb.x, bb.y)), y0, y1);
    upk2(fma2(pk2u(v[c0 + e + 2], v[c0 + e + 3]), r2, pk2(bb.z, bb.w)), y2, y3);
    zp[e / 2] = pack_bf16_relu(y0, y1);
    zp[e / 2 + 1] = pack_bf16_relu(y2, y3);
  }
}

// ---- packed bf16 conversions -----------------------------------------------------------------------
// (lo16 = bf16(a), hi16 = bf16(b)), round to nearest even, optional fused ReLU
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// hi/lo split of two (optionally ReLU'd) values: v ~= hi + lo, 16 significant bits
template <bool RELU>
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = RELU ? pack_bf16_relu(a, b) : pack_bf16(a, b);
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  if (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
  lo = pack_bf16(a - ah, b - bh);
}

}  // namespace tc
}  // namespace smb
