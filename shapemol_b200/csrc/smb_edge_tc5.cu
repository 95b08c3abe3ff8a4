// tcgen05 / TMEM implementation of the fused edge kernels (plain-bf16 contraction mode, molecules of
// <= 32 atoms: the MOSES workload).  Same math and same EdgeArgs contract as smb_edge_attn.cu
// (BaseX2HAttLayer / BaseH2XAttLayer, models/uni_transformer.py:48-162); this file only changes HOW
// the two Linear layers of an edge MLP are contracted.
//
// Tile = 128 edge rows = a run of whole destination atoms of ONE molecule (deg <= 31 each).
//
//   GEMM1 (SS)  D1[128 x 128] = A1[128 x 96] . B1[96 x 128]
//       A1 row (i <- j) = [ rbf(|x_i - x_j|) (20, padded to 32) | onehot_32(i) | onehot_32(j) ]   (bf16, smem)
//       B1              = [ W1r (32 x 128)  ;  A-tile (32 atoms x 128) ; B-tile (32 atoms x 128) ] (bf16, smem)
//     i.e. the gather of the node-level projections A_i + B_j is done BY the tensor core:
//     the one-hot columns select the rows of the molecule's A/B tiles, no per-edge SIMT gather/add.
//   E1: thread = row: LayerNorm over the row's 128 TMEM columns is in-thread, ReLU, bf16.
//   GEMM2
//     ROLE_K  (TS) D2[128 x 128] = z (A operand from TMEM) . W2^T     -> <Q_i, .> per head, softmax over
//                                                                       the destination's rows -> alpha
//     ROLE_XV (TS) D2[128 x 16]  = z . W2xv^T                         -> alpha * w * (x_i - x_j) -> VN maps
//     ROLE_V  (SS) D2^T[128 ch x 128 rows] = W2 . z^T  (z in smem as the B operand): lane = channel,
//              column = edge row, so the per-destination sum over neighbour rows is IN-THREAD.
//
// One tile in flight per CTA, two CTAs per SM (<= 113 KB smem, 256 TMEM columns each): the SM overlaps
// one CTA's SIMT epilogue with the other's MMAs.
#include "smb_common.cuh"
#include "smb_kernels.h"
#include "smb_tc.cuh"

#include <cstdlib>

namespace smb {

namespace {

using namespace tc;

constexpr int H = 128;
constexpr int G = 32;                  // atoms per molecule supported by the one-hot operand
constexpr int TM = 128;                // rows per tile
constexpr int THREADS = 256;           // 8 warps: (row quarter = warp & 3, column half = warp >> 2)
constexpr int WARPS5 = THREADS / 32;
constexpr int K1 = 32 + 2 * G;         // 96
constexpr int A1_SBO = (K1 / 8) * 128; // 1536
constexpr int LS = 17;                 // padded row stride of per-row head scratch
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t Z_COL = 128;        // z (A-from-TMEM operand) lives in columns [128, 192)

template <int ROLE>
struct Plan {
  static constexpr int o_bar = 0;                 // mbarrier (8)
  static constexpr int o_tmem = 8;                // TMEM base (4)
  static constexpr int o_last = 16;               // 4 x u32 "last row of its destination" masks
  static constexpr int o_pref = 64;               // int[33] prefix of deg over the molecule's atoms
  static constexpr int o_rowdst = 256;            // u8[128]
  static constexpr int o_rowslot = 384;           // u8[128]
  static constexpr int o_x = 512;                 // float[32][4]
  static constexpr int o_vec = 1024;              // ln_g | ln_b | b2   (3 x 128 floats)
  static constexpr int o_stat = 2560;             // float2[2][128]
  static constexpr int o_w1r = 4608;              // 8192
  static constexpr int o_w2 = o_w1r + 8192;
  static constexpr int w2_bytes = ROLE == ROLE_XV ? kHeads * H * 2 : H * H * 2;
  static constexpr int o_sa = o_w2 + w2_bytes;    // A-tile  [32 k][128 n] MN-major, 8192
  static constexpr int o_sb = o_sa + 8192;        // B-tile
  static constexpr int o_a1 = o_sb + 8192;        // A1 (24576); ROLE_V: reused for z^T operand (32768)
  static constexpr int a1_bytes = ROLE == ROLE_V ? TM * H * 2 : TM * K1 * 2;
  static constexpr int o_role = o_a1 + a1_bytes;
  // ROLE_K : logits float[128][17]
  // ROLE_V : alpha float[128][16] | part float2[128]
  // ROLE_XV: w float[128][17] | rel float4[128] | o float[32][16][4] | vnw float[2][16][49] | shape float[96]
  static constexpr int o_r0 = o_role;
  static constexpr int o_r1 = o_r0 + (ROLE == ROLE_V ? TM * 16 * 4 : TM * LS * 4);
  static constexpr int o_r2 = o_r1 + (ROLE == ROLE_V ? TM * 8 : ROLE == ROLE_XV ? TM * 16 : 0);
  static constexpr int o_r3 = o_r2 + (ROLE == ROLE_XV ? G * 16 * 4 * 4 : 0);
  static constexpr int o_r4 = o_r3 + (ROLE == ROLE_XV ? 2 * kHeads * kVnStride * 4 : 0);
  static constexpr int total = o_r4 + (ROLE == ROLE_XV ? kShape * 3 * 4 : 0);
  static_assert(total <= 113 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
  return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

template <int ROLE>
__global__ void __launch_bounds__(THREADS, 2) edge5_kernel(EdgeArgs a) {
  using P = Plan<ROLE>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P::o_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::o_tmem);
  uint32_t* s_last = reinterpret_cast<uint32_t*>(smem + P::o_last);
  int* s_pref = reinterpret_cast<int*>(smem + P::o_pref);
  uint8_t* s_rowdst = smem + P::o_rowdst;
  uint8_t* s_rowslot = smem + P::o_rowslot;
  float* s_x = reinterpret_cast<float*>(smem + P::o_x);
  float* s_g = reinterpret_cast<float*>(smem + P::o_vec);
  float* s_be = s_g + H;
  float* s_b2 = s_be + H;
  float2* s_stat = reinterpret_cast<float2*>(smem + P::o_stat);
  unsigned char* s_w1r = smem + P::o_w1r;
  unsigned char* s_w2 = smem + P::o_w2;
  unsigned char* s_sa = smem + P::o_sa;
  unsigned char* s_sb = smem + P::o_sb;
  unsigned char* s_a1 = smem + P::o_a1;
  float* s_r0 = reinterpret_cast<float*>(smem + P::o_r0);            // logits / alpha / w
  float2* s_part = reinterpret_cast<float2*>(smem + P::o_r1);        // ROLE_V
  float4* s_rel = reinterpret_cast<float4*>(smem + P::o_r1);         // ROLE_XV
  float* s_o = reinterpret_cast<float*>(smem + P::o_r2);             // ROLE_XV
  float* s_vnw = reinterpret_cast<float*>(smem + P::o_r3);           // ROLE_XV
  float* s_shape = reinterpret_cast<float*>(smem + P::o_r4);         // ROLE_XV

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;
  const int KSTR = a.k + 1;

  // ---- once per CTA: weights -> smem, TMEM allocation, mbarrier ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.w1r_u);
    uint4* dst = reinterpret_cast<uint4*>(s_w1r);
    for (int p = tid; p < 8192 / 16; p += THREADS) dst[p] = src[p];
    const uint4* s2 = reinterpret_cast<const uint4*>(a.w2_u);
    uint4* d2 = reinterpret_cast<uint4*>(s_w2);
    for (int p = tid; p < P::w2_bytes / 16; p += THREADS) d2[p] = s2[p];
    if (tid < H) {
      s_g[tid] = a.ln_g[tid];
      s_be[tid] = a.ln_b[tid];
      s_b2[tid] = ROLE == ROLE_XV ? (tid < kHeads ? a.b2[tid] : 0.f) : a.b2[tid];
    }
    if (ROLE == ROLE_XV)
      for (int p = tid; p < kHeads * kVnStride; p += THREADS) {
        s_vnw[p] = a.vn_feat[p];
        s_vnw[kHeads * kVnStride + p] = a.vn_dir[p];
      }
    // the A/B tiles are zero-initialised once: rows >= n of later molecules keep finite stale data that
    // no one-hot column selects
    uint4* z0 = reinterpret_cast<uint4*>(s_sa);
    for (int p = tid; p < 2 * 8192 / 16; p += THREADS) z0[p] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (tid == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  float bn_s = 0.f, bn_q = 0.f;   // ROLE_XV: per-warp BatchNorm partial sums (lane & 15 = channel, lanes < 16)

  constexpr uint32_t IDESC1 = idesc_bf16(H, true);
  constexpr uint32_t IDESC2 = idesc_bf16(ROLE == ROLE_XV ? kHeads : H, false);

#pragma unroll 1
  for (int m = blockIdx.x; m < a.n_mols; m += gridDim.x) {
    const int a0 = a.mol_ptr[m];
    const int n = a.mol_ptr[m + 1] - a0;
    if (n <= 0) continue;
    if (n == 1 && ROLE != ROLE_XV) {      // no edges: empty neighbour sum (ROLE_K has nothing to write)
      if (ROLE == ROLE_V && tid < H) a.agg[(size_t)a0 * H + tid] = 0.f;
      continue;
    }
    // ---- per molecule: coordinates, degree prefix, A/B tiles (fp32 -> bf16, MN-major UMMA layout) ----
    if (tid < n) {
      s_x[tid * 4] = a.x[(size_t)(a0 + tid) * 3];
      s_x[tid * 4 + 1] = a.x[(size_t)(a0 + tid) * 3 + 1];
      s_x[tid * 4 + 2] = a.x[(size_t)(a0 + tid) * 3 + 2];
    }
    if (warp == 1) {
      int dg = lane < n ? a.deg[a0 + lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, dg, o);
        if (lane >= o) dg += up;
      }
      s_pref[lane + 1] = dg;
      if (lane == 0) s_pref[0] = 0;
    }
    for (int p = tid; p < G * 16 * 2; p += THREADS) {
      const int kk = p & 31, nc = (p >> 5) & 15, which = p >> 9;
      if (kk < n) {
        const float* src = a.ab + (size_t)(a0 + kk) * (4 * H) + (which ? a.col_b : a.col_a) + nc * 8;
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(src));
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(src + 4));
        *reinterpret_cast<uint4*>((which ? s_sb : s_sa) + nc * 512 + kk * 16) = pack8(f0, f1);
      }
    }
    if (ROLE == ROLE_XV)
      for (int p = tid; p < kShape * 3; p += THREADS) s_shape[p] = a.shape[(size_t)m * kShape * 3 + p];
    __syncthreads();

    int d0 = 0;
#pragma unroll 1
    while (d0 < n) {
      const int base = s_pref[d0];
      int d1 = d0 + 1;
      while (d1 < n && s_pref[d1 + 1] - base <= TM) ++d1;
      const int rows = s_pref[d1] - base;
      const int nd = d1 - d0;

      // =========================== P: build the A1 operand ===========================
      const bool valid = r < rows;
      int i = 0, s = 0, j = 0;
      float rx = 0.f, ry = 0.f, rz = 0.f, dist = 0.f;
      if (valid) {
        i = d0;
        while (s_pref[i + 1] - base <= r) ++i;
        s = r - (s_pref[i] - base);
        j = a.nbr[(size_t)(a0 + i) * KSTR + s];
        rx = s_x[i * 4] - s_x[j * 4]; ry = s_x[i * 4 + 1] - s_x[j * 4 + 1]; rz = s_x[i * 4 + 2] - s_x[j * 4 + 2];
        dist = sqrtf(rx * rx + ry * ry + rz * rz);
      }
      {
        unsigned char* arow = s_a1 + (r >> 3) * A1_SBO + (r & 7) * 16;
        if (half == 0) {
          const bool last = valid && (s_pref[i + 1] - s_pref[i] == s + 1);
          const uint32_t lm = __ballot_sync(0xffffffffu, last);
          if (lane == 0) s_last[warp] = lm;
          s_rowdst[r] = valid ? (uint8_t)i : (uint8_t)255;
          s_rowslot[r] = (uint8_t)s;
          if (ROLE == ROLE_XV) s_rel[r] = make_float4(rx, ry, rz, 0.f);
          // rbf columns 0..15
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float dd = dist - rbf_centre(c * 8 + q);
              e[q] = valid ? exp2f(-0.72134752044448170f * dd * dd) : 0.f;
            }
            *reinterpret_cast<uint4*>(arow + c * 128) =
                make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
          }
        } else {
          // rbf columns 16..19 (20..31 are zero padding)
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float dd = dist - rbf_centre(16 + q);
            e[q] = valid ? exp2f(-0.72134752044448170f * dd * dd) : 0.f;
          }
          *reinterpret_cast<uint4*>(arow + 2 * 128) = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), 0u, 0u);
          *reinterpret_cast<uint4*>(arow + 3 * 128) = make_uint4(0u, 0u, 0u, 0u);
          // one-hot(dst) in k = 32..63, one-hot(src) in k = 64..95
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int idx = w ? j : i;
            const uint32_t one = (idx & 1) ? 0x3F800000u : 0x00003F80u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (valid && (idx >> 3) == c) {
                const int wd = (idx & 7) >> 1;
                v.x = wd == 0 ? one : 0u; v.y = wd == 1 ? one : 0u; v.z = wd == 2 ? one : 0u; v.w = wd == 3 ? one : 0u;
              }
              *reinterpret_cast<uint4*>(arow + (4 + 4 * w + c) * 128) = v;
            }
          }
        }
      }
      if (ROLE == ROLE_V) {   // stage this tile's alpha (softmax x gate, written by ROLE_K) : [row][16]
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (valid) {
          const float4* al = reinterpret_cast<const float4*>(a.alpha + ((size_t)(a0 + i) * KSTR + s) * kHeads + half * 8);
          v0 = al[0]; v1 = al[1];
        }
        float4* dst = reinterpret_cast<float4*>(s_r0 + r * 16 + half * 8);
        dst[0] = v0; dst[1] = v1;
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();

      // =========================== GEMM1 ===========================
      if (tid == 0) {
        fence_after_sync();
        const uint32_t a1 = smem_u32(s_a1);
        const uint32_t b1[3] = {smem_u32(s_w1r), smem_u32(s_sa), smem_u32(s_sb)};
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks)
          mma_ss(tmem, smem_desc(a1 + ks * 256, 128, A1_SBO), smem_desc(b1[ks >> 1] + (ks & 1) * 256, 128, 512), IDESC1, ks > 0);
        mma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();

      // =========================== E1: LayerNorm + ReLU -> z (bf16) ===========================
      {
        uint32_t v[64];
        tmem_ld32(lane_addr + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(lane_addr + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        wait_ld();
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const float f = __uint_as_float(v[e]);
          sum += f;
          sq = fmaf(f, f, sq);
        }
        s_stat[half * TM + r] = make_float2(sum, sq);
        __syncthreads();
        const float2 ot = s_stat[(half ^ 1) * TM + r];
        const float mean = (sum + ot.x) * (1.f / H);
        const float var = fmaxf((sq + ot.y) * (1.f / H) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        const float shift = -mean * rstd;
        uint32_t zp[32];
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          const float4 gg = *reinterpret_cast<const float4*>(s_g + half * 64 + e);
          const float4 bb = *reinterpret_cast<const float4*>(s_be + half * 64 + e);
          const float y0 = fmaf(fmaf(__uint_as_float(v[e]), rstd, shift), gg.x, bb.x);
          const float y1 = fmaf(fmaf(__uint_as_float(v[e + 1]), rstd, shift), gg.y, bb.y);
          const float y2 = fmaf(fmaf(__uint_as_float(v[e + 2]), rstd, shift), gg.z, bb.z);
          const float y3 = fmaf(fmaf(__uint_as_float(v[e + 3]), rstd, shift), gg.w, bb.w);
          zp[e / 2] = pack_bf16_relu(y0, y1);
          zp[e / 2 + 1] = pack_bf16_relu(y2, y3);
        }
        if (ROLE == ROLE_V) {
          // z^T operand: K-major [row][k]; this thread owns k = 64 half .. 64 half + 63 of row r
          unsigned char* zrow = s_a1 + (r >> 3) * 2048 + (r & 7) * 16 + half * 8 * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(zrow + q * 128) = make_uint4(zp[4 * q], zp[4 * q + 1], zp[4 * q + 2], zp[4 * q + 3]);
          fence_async_smem();
        } else {
          tmem_st32(lane_addr + Z_COL + half * 32, zp);
          wait_st();
        }
      }
      fence_before_sync();
      __syncthreads();

      // =========================== GEMM2 ===========================
      if (tid == 0) {
        fence_after_sync();
        const uint32_t w2 = smem_u32(s_w2);
        if (ROLE == ROLE_V) {
          const uint32_t zt = smem_u32(s_a1);
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks)
            mma_ss(tmem, smem_desc(w2 + ks * 256, 128, 2048), smem_desc(zt + ks * 256, 128, 2048), IDESC2, ks > 0);
        } else {
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks)
            mma_ts(tmem, tmem + Z_COL + ks * 8, smem_desc(w2 + ks * 256, 128, 2048), IDESC2, ks > 0);
        }
        mma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();

      // =========================== E2 ===========================
      if (ROLE == ROLE_K) {
        {
          uint32_t v[64];
          tmem_ld32(lane_addr + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(lane_addr + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          wait_ld();
          // b2 shifts every logit of (i, head) by the same <Q_i, b2>: softmax-invariant, dropped
          const float4* qrow = reinterpret_cast<const float4*>(a.q + (size_t)(a0 + i) * H + half * 64);
          const float scale = 0.35355339059327373f;   // 1/sqrt(dh), dh = 8
#pragma unroll
          for (int hh = 0; hh < 8; ++hh) {
            const float4 q0 = __ldg(qrow + 2 * hh), q1 = __ldg(qrow + 2 * hh + 1);
            float l = __uint_as_float(v[8 * hh]) * q0.x;
            l = fmaf(__uint_as_float(v[8 * hh + 1]), q0.y, l);
            l = fmaf(__uint_as_float(v[8 * hh + 2]), q0.z, l);
            l = fmaf(__uint_as_float(v[8 * hh + 3]), q0.w, l);
            l = fmaf(__uint_as_float(v[8 * hh + 4]), q1.x, l);
            l = fmaf(__uint_as_float(v[8 * hh + 5]), q1.y, l);
            l = fmaf(__uint_as_float(v[8 * hh + 6]), q1.z, l);
            l = fmaf(__uint_as_float(v[8 * hh + 7]), q1.w, l);
            s_r0[r * LS + half * 8 + hh] = l * scale;
          }
        }
        __syncthreads();
        // softmax over the rows of each (destination, head); alpha * e_w -> global
        for (int p = tid; p < nd * kHeads; p += THREADS) {
          const int dl = p >> 4, hd = p & 15;
          const int r0 = s_pref[d0 + dl] - base, dg = s_pref[d0 + dl + 1] - s_pref[d0 + dl];
          float mx = -INFINITY;
          for (int q = 0; q < dg; ++q) mx = fmaxf(mx, s_r0[(r0 + q) * LS + hd]);
          float se = 0.f;
          for (int q = 0; q < dg; ++q) {
            const float e = __expf(s_r0[(r0 + q) * LS + hd] - mx);
            s_r0[(r0 + q) * LS + hd] = e;
            se += e;
          }
          const float inv = 1.f / se;
          const size_t gi = (size_t)(a0 + d0 + dl);
          float* al = a.alpha + gi * KSTR * kHeads;
          const float* ew = a.ew_in + gi * KSTR;
          for (int q = 0; q < dg; ++q) al[q * kHeads + hd] = s_r0[(r0 + q) * LS + hd] * inv * ew[q];
        }
      } else if (ROLE == ROLE_V) {
        // thread = output channel c (TMEM lane), columns = edge rows; this half covers rows [64 half, 64 half + 64)
        const int c = r, hq = c >> 3;
        const bool straddle = rows > 64 && s_rowdst[63] == s_rowdst[64];
        float acc = 0.f, asum = 0.f, first_acc = 0.f, first_asum = 0.f;
        int first_i = -1;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int col0 = half * 64 + ch * 32;
          if (col0 >= rows) break;
          uint32_t v[32];
          tmem_ld32(lane_addr + col0, v);
          wait_ld();
          const uint32_t lm = s_last[col0 >> 5];
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int rr = col0 + q;
            if (rr < rows) {
              const float al = s_r0[rr * 16 + hq];
              acc = fmaf(al, __uint_as_float(v[q]), acc);
              asum += al;
              if ((lm >> q) & 1u) {
                const int di = s_rowdst[rr];
                if (half == 1 && straddle && first_i < 0) {
                  first_i = di; first_acc = acc; first_asum = asum;
                } else {
                  a.agg[(size_t)(a0 + di) * H + c] = fmaf(s_b2[c], asum, acc);
                }
                acc = 0.f; asum = 0.f;
              }
            }
          }
        }
        if (half == 0 && straddle) s_part[c] = make_float2(acc, asum);
        __syncthreads();
        if (half == 1 && first_i >= 0) {
          const float2 pp = s_part[c];
          a.agg[(size_t)(a0 + first_i) * H + c] = fmaf(s_b2[c], first_asum + pp.y, first_acc + pp.x);
        }
      } else {   // ROLE_XV
        if (half == 0) {
          uint32_t v[16];
          tmem_ld16(lane_addr, v);
          wait_ld();
          if (valid) {
            const float4* al = reinterpret_cast<const float4*>(a.alpha + ((size_t)(a0 + i) * KSTR + s) * kHeads);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 av = al[q];
              s_r0[r * LS + 4 * q] = av.x * (__uint_as_float(v[4 * q]) + s_b2[4 * q]);
              s_r0[r * LS + 4 * q + 1] = av.y * (__uint_as_float(v[4 * q + 1]) + s_b2[4 * q + 1]);
              s_r0[r * LS + 4 * q + 2] = av.z * (__uint_as_float(v[4 * q + 2]) + s_b2[4 * q + 2]);
              s_r0[r * LS + 4 * q + 3] = av.w * (__uint_as_float(v[4 * q + 3]) + s_b2[4 * q + 3]);
            }
          }
        }
        __syncthreads();
        // o_i^a = sum_j alpha e_w w (x_i - x_j)
        for (int p = tid; p < nd * kHeads; p += THREADS) {
          const int dl = p >> 4, hd = p & 15;
          const int r0 = s_pref[d0 + dl] - base, dg = s_pref[d0 + dl + 1] - s_pref[d0 + dl];
          float ox = 0.f, oy = 0.f, oz = 0.f;
          for (int q = 0; q < dg; ++q) {
            const float w = s_r0[(r0 + q) * LS + hd];
            const float4 rl = s_rel[r0 + q];
            ox = fmaf(w, rl.x, ox); oy = fmaf(w, rl.y, oy); oz = fmaf(w, rl.z, oz);
          }
          float* o = s_o + (dl * kHeads + hd) * 4;
          o[0] = ox; o[1] = oy; o[2] = oz;
        }
        __syncthreads();
        // VN linear maps (shape_vn_layers.py:100,105): lanes 0..15 map_to_feat channel, 16..31 map_to_dir channel
        for (int dl = warp; dl < nd; dl += WARPS5) {
          const int ch = lane & 15, which = lane >> 4;
          const float* w = s_vnw + (which * kHeads + ch) * kVnStride;
          const float* so = s_o + dl * kHeads * 4;
          const float xi = s_x[(d0 + dl) * 4], yi = s_x[(d0 + dl) * 4 + 1], zi = s_x[(d0 + dl) * 4 + 2];
          float vx = w[0] * xi, vy = w[0] * yi, vz = w[0] * zi;
#pragma unroll
          for (int cc = 0; cc < kHeads; ++cc) {
            const float wc = w[1 + cc];
            vx = fmaf(wc, so[cc * 4], vx); vy = fmaf(wc, so[cc * 4 + 1], vy); vz = fmaf(wc, so[cc * 4 + 2], vz);
          }
#pragma unroll 8
          for (int cc = 0; cc < kShape; ++cc) {
            const float wc = w[1 + kHeads + cc];
            vx = fmaf(wc, s_shape[cc * 3], vx); vy = fmaf(wc, s_shape[cc * 3 + 1], vy); vz = fmaf(wc, s_shape[cc * 3 + 2], vz);
          }
          float* row = a.vn + (size_t)(a0 + d0 + dl) * kVnRow;
          row[3 + which * 48 + ch * 3] = vx; row[4 + which * 48 + ch * 3] = vy; row[5 + which * 48 + ch * 3] = vz;
          if (lane < 3) {
            float sm = 0.f;
#pragma unroll
            for (int cc = 0; cc < kHeads; ++cc) sm += so[cc * 4 + lane];
            row[lane] = sm * (1.f / kHeads);
          }
          if (which == 0) {
            const float nu = sqrtf(vx * vx + vy * vy + vz * vz) + 1e-6f;
            bn_s += nu; bn_q = fmaf(nu, nu, bn_q);
          }
        }
      }
      fence_before_sync();
      __syncthreads();   // tile scratch, A1 and the TMEM accumulator are free again
      d0 = d1;
    }   // tiles
  }     // molecules

  if (ROLE == ROLE_XV && lane < 16) {
    float* part = a.bn_partial + (size_t)(blockIdx.x * WARPS5 + warp) * 32;
    part[lane] = bn_s;
    part[16 + lane] = bn_q;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<TMEM_COLS>(tmem);
}

int g_sms5 = 0;
int sms5() {
  if (g_sms5 == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sms5 = n;
    else
      g_sms5 = 148;
  }
  return g_sms5;
}

template <int ROLE>
int launch5(const EdgeArgs& a, int* bn_rows_out, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(edge5_kernel<ROLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<ROLE>::total);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int grid = 2 * sms5();
  if (grid > kEdgeMaxCtas) grid = kEdgeMaxCtas;
  if (grid > a.n_mols) grid = a.n_mols;
  if (bn_rows_out) *bn_rows_out = grid * WARPS5;
  edge5_kernel<ROLE><<<grid, THREADS, Plan<ROLE>::total, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

bool edge_tc5_supported(const smb_model_dims& d, int role, const EdgeArgs& a) {
  static const bool legacy = getenv("SMB_EDGE_LEGACY") != nullptr;   // debugging aid: force the mma.sync kernels
  return !legacy && d.precision == SMB_PREC_BF16 && d.hidden == H && role != ROLE_GATE && a.n_max >= 1 && a.n_max <= G;
}

int launch_edge_tc5(int role, const EdgeArgs& a, int* bn_rows_out, cudaStream_t st) {
  switch (role) {
    case ROLE_K: return launch5<ROLE_K>(a, bn_rows_out, st);
    case ROLE_V: return launch5<ROLE_V>(a, bn_rows_out, st);
    case ROLE_XV: return launch5<ROLE_XV>(a, bn_rows_out, st);
    default: set_error_msg("launch_edge_tc5: unsupported role"); return SMB_E_UNSUPPORTED;
  }
}

}  // namespace smb
