// Node-level chain GEMM on tensor cores (mma.sync bf16 / split-bf16, fp32 accumulate).
//
//   stage 1:  Y1[128 x N1] = X[128 x K1] W1^T + b1      X assembled on the fly from (h | inv[mol]) ,
//                                                        (agg | h) or h
//             the first n_pass columns are stored as they are (dst/src projections of the edge MLPs'
//             first Linear), the last H columns stay in registers
//   act    :  LayerNorm(eps 1e-5) + ReLU   (models/common.py:50-64)   or  softplus - ln2 (:39-45)
//   stage 2:  Y2[128 x N2] = Z W2^T + b2 (+ residual)
//
// One CTA = 128 atoms, 8 warps x 16 rows; weights stream L2 -> smem through a 2-deep cp.async ring
// and are shared by the 8 warps.  Bound: tensor pipe / smem bandwidth (weights re-read per 128 rows).
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

namespace {

constexpr int H = 128;
constexpr int NT_CHUNK = 8;          // n-tiles (of 8 columns) per weight tile
constexpr int KS2 = H / 16;

template <bool X3> struct Frag { using type = uint4; };
template <> struct Frag<false> { using type = uint2; };

template <bool X3>
__device__ __forceinline__ uint4 ld_frag(const typename Frag<X3>::type* p) {
  if constexpr (X3) return *p;
  else { const uint2 v = *p; return make_uint4(v.x, v.y, 0u, 0u); }
}

// copy `nt` n-tiles x `ks` k-steps of fragments (source n-tile stride = ks_total k-steps) into smem
template <bool X3>
__device__ __forceinline__ void load_tile(typename Frag<X3>::type* dst, const typename Frag<X3>::type* src, int nt0,
                                          int nt, int ks_total, int ks0, int ks) {
  constexpr int FB = X3 ? 16 : 8;
  const int per_nt = ks * 32 * FB / 16;     // 16-byte pieces per n-tile
  const int total = nt * per_nt;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int n = p / per_nt, r = p % per_nt;
    const char* s = reinterpret_cast<const char*>(src + ((size_t)(nt0 + n) * ks_total + ks0) * 32) + (size_t)r * 16;
    char* d = reinterpret_cast<char*>(dst + (size_t)n * ks * 32) + (size_t)r * 16;
    cp_async16(d, s);
  }
}

// C[nt][4] += A(ks k-steps) * Btile
template <bool X3, int KSN, int NTN>
__device__ __forceinline__ void gemm_tile(float (&c)[NTN][4], const uint32_t (&ahi)[KSN][4], const uint32_t (&alo)[KSN][4],
                                          const typename Frag<X3>::type* bt, int lane) {
#pragma unroll
  for (int ks = 0; ks < KSN; ++ks) {
    uint4 bw[NTN];
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt) bw[nt] = ld_frag<X3>(bt + (nt * KSN + ks) * 32 + lane);
    mma_step_n<X3, NTN>(c, ahi[ks], alo[ks], bw);
  }
}

// MODE 0: pre  (X = [h | inv], K1 = 160, n_pass = 512)     -> out1 [N][512], out2 = Q [N][128]
// MODE 1: out  (X = [agg | h], K1 = 256, two K blocks)     -> out2 = h' = MLP + residual
// MODE 2: head (X = h, K1 = 128, shifted softplus, N2=16)  -> out2 = logits [N][n2_valid]
template <int MODE, bool X3>
__global__ void __launch_bounds__(256, 1) node_mlp_kernel(NodeArgs a) {
  using F = typename Frag<X3>::type;
  constexpr int KB = MODE == 0 ? 160 : 128;      // K block held in registers
  constexpr int NKB = MODE == 1 ? 2 : 1;
  constexpr int KSB = KB / 16;                   // k-steps per block
  constexpr int KS1 = KSB * NKB;                 // k-steps of W1 in total
  constexpr int NPASS_CH = MODE == 0 ? 8 : 0;    // pass-through chunks of 64 columns
  constexpr int N2T = MODE == 2 ? 2 : 16;        // stage-2 n-tiles
  constexpr int TILE_ELEMS = NT_CHUNK * (KSB > KS2 ? KSB : KS2) * 32;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  F* buf0 = reinterpret_cast<F*>(smem_raw);
  F* buf1 = buf0 + TILE_ELEMS;
  float* s_g = reinterpret_cast<float*>(buf1 + TILE_ELEMS);
  float* s_b = s_g + H;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 128 + warp * 16 + g, row1 = row0 + 8;
  const bool ok0 = row0 < a.n_atoms, ok1 = row1 < a.n_atoms;
  const F* w1 = reinterpret_cast<const F*>(a.w1);
  const F* w2 = reinterpret_cast<const F*>(a.w2);

  if (threadIdx.x < H) {
    s_g[threadIdx.x] = MODE == 2 ? 0.f : a.ln_g[threadIdx.x];
    s_b[threadIdx.x] = MODE == 2 ? 0.f : a.ln_b[threadIdx.x];
  }

  auto xval = [&](int row, bool ok, int col) -> float2 {
    if (!ok) return make_float2(0.f, 0.f);
    if (col < H) return *reinterpret_cast<const float2*>(a.xa + (size_t)row * H + col);
    if (MODE == 0) return *reinterpret_cast<const float2*>(a.xb + (size_t)a.atom_mol[row] * kShape + (col - H));
    return *reinterpret_cast<const float2*>(a.xb + (size_t)row * H + (col - H));
  };

  uint32_t ahi[KSB][4], alo[KSB][4];
  auto load_a = [&](int kb) {
#pragma unroll
    for (int ks = 0; ks < KSB; ++ks) {
      const int c0 = kb * KB + ks * 16 + 2 * t;
      const float2 v00 = xval(row0, ok0, c0), v10 = xval(row1, ok1, c0);
      const float2 v01 = xval(row0, ok0, c0 + 8), v11 = xval(row1, ok1, c0 + 8);
      split_bf16x2(v00.x, v00.y, ahi[ks][0], alo[ks][0]);
      split_bf16x2(v10.x, v10.y, ahi[ks][1], alo[ks][1]);
      split_bf16x2(v01.x, v01.y, ahi[ks][2], alo[ks][2]);
      split_bf16x2(v11.x, v11.y, ahi[ks][3], alo[ks][3]);
    }
  };

  // molecule of each row (only for the operand-image output)
  int ma0 = 0, mn0 = 1, ma1 = 0, mn1 = 1;
  if (MODE == 0 && a.out1_h) {
    if (ok0) { const int m = a.atom_mol[row0]; ma0 = a.mol_ptr[m]; mn0 = a.mol_ptr[m + 1] - ma0; }
    if (ok1) { const int m = a.atom_mol[row1]; ma1 = a.mol_ptr[m]; mn1 = a.mol_ptr[m + 1] - ma1; }
  }

  float hid[2][NT_CHUNK][4];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int nt = 0; nt < NT_CHUNK; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) hid[c][nt][e] = 0.f;

  // ---- job list: [pass chunks (kb 0)] [hidden chunk 0/1 for each kb] [stage-2 tiles] ----
  constexpr int N_S1 = NPASS_CH + 2 * NKB;
  constexpr int N_S2 = (N2T + NT_CHUNK - 1) / NT_CHUNK;
  constexpr int N_JOBS = N_S1 + N_S2;
  auto issue = [&](int job) {
    F* dst = (job & 1) ? buf1 : buf0;
    if (job < N_S1) {
      int chunk, kb;
      if (job < NPASS_CH) { chunk = job; kb = 0; }
      else { const int j = job - NPASS_CH; kb = j / 2; chunk = NPASS_CH + (j & 1); }
      load_tile<X3>(dst, w1, chunk * NT_CHUNK, NT_CHUNK, KS1, kb * KSB, KSB);
    } else {
      const int c2 = job - N_S1;
      const int nt = N2T < NT_CHUNK ? N2T : NT_CHUNK;
      load_tile<X3>(dst, w2, c2 * NT_CHUNK, nt, KS2, 0, KS2);
    }
    cp_async_commit();
  };

  issue(0);
  load_a(0);
  uint32_t zhi[KS2][4], zlo[KS2][4];
  int job = 0;

  // ---- pass-through chunks: store the dst/src projections ----
#pragma unroll 1
  for (int pc = 0; pc < NPASS_CH; ++pc, ++job) {
    if (job + 1 < N_JOBS) { issue(job + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const F* bt = (job & 1) ? buf1 : buf0;
    float c[NT_CHUNK][4];
#pragma unroll
    for (int nt = 0; nt < NT_CHUNK; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
    gemm_tile<X3, KSB, NT_CHUNK>(c, ahi, alo, bt, lane);
#pragma unroll
    for (int nt = 0; nt < NT_CHUNK; ++nt) {
      const int col = pc * 64 + nt * 8 + 2 * t;
      const float2 bb = *reinterpret_cast<const float2*>(a.b1 + col);
      if (a.out1_h) {   // bf16, per-molecule MN-major operand image for the tcgen05 edge pipeline (one bulk copy per tile)
        const int part = col >> 7, cc = col & 127;
        if (ok0) a.out1_h[((size_t)ma0 * 1024 + (size_t)part * mn0 * 256 + (cc >> 3) * mn0 * 16 + (row0 - ma0) * 16 + (cc & 7) * 2) >> 2] =
            pack_bf16x2(c[nt][0] + bb.x, c[nt][1] + bb.y);
        if (ok1) a.out1_h[((size_t)ma1 * 1024 + (size_t)part * mn1 * 256 + (cc >> 3) * mn1 * 16 + (row1 - ma1) * 16 + (cc & 7) * 2) >> 2] =
            pack_bf16x2(c[nt][2] + bb.x, c[nt][3] + bb.y);
      } else {
        if (ok0) *reinterpret_cast<float2*>(a.out1 + (size_t)row0 * a.n_pass + col) = make_float2(c[nt][0] + bb.x, c[nt][1] + bb.y);
        if (ok1) *reinterpret_cast<float2*>(a.out1 + (size_t)row1 * a.n_pass + col) = make_float2(c[nt][2] + bb.x, c[nt][3] + bb.y);
      }
    }
    __syncthreads();
  }

  // ---- hidden chunks (accumulated over the K blocks) ----
#pragma unroll
  for (int j = 0; j < 2 * NKB; ++j, ++job) {
    if (job + 1 < N_JOBS) { issue(job + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const F* bt = (job & 1) ? buf1 : buf0;
    if (NKB == 2 && j == 2) load_a(1);      // second K block (the h part of [agg | h])
    if ((j & 1) == 0) gemm_tile<X3, KSB, NT_CHUNK>(hid[0], ahi, alo, bt, lane);
    else gemm_tile<X3, KSB, NT_CHUNK>(hid[1], ahi, alo, bt, lane);
    __syncthreads();
  }

  // ---- bias + activation, then re-pack as A fragments of stage 2 ----
  {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int nt = 0; nt < NT_CHUNK; ++nt) {
        const int col = NPASS_CH * 64 + c * 64 + nt * 8 + 2 * t;
        const float2 bb = *reinterpret_cast<const float2*>(a.b1 + col);
        hid[c][nt][0] += bb.x; hid[c][nt][1] += bb.y; hid[c][nt][2] += bb.x; hid[c][nt][3] += bb.y;
        s0 += hid[c][nt][0] + hid[c][nt][1];
        s1 += hid[c][nt][2] + hid[c][nt][3];
      }
    if (MODE != 2) {
      const float m0 = quad_sum(s0) * (1.f / H), m1 = quad_sum(s1) * (1.f / H);
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int nt = 0; nt < NT_CHUNK; ++nt) {
          float d;
          d = hid[c][nt][0] - m0; v0 = fmaf(d, d, v0); d = hid[c][nt][1] - m0; v0 = fmaf(d, d, v0);
          d = hid[c][nt][2] - m1; v1 = fmaf(d, d, v1); d = hid[c][nt][3] - m1; v1 = fmaf(d, d, v1);
        }
      const float r0 = 1.f / sqrtf(quad_sum(v0) * (1.f / H) + 1e-5f), r1 = 1.f / sqrtf(quad_sum(v1) * (1.f / H) + 1e-5f);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int nt = 0; nt < NT_CHUNK; ++nt) {
          const int col = c * 64 + nt * 8 + 2 * t;
          const float2 gg = *reinterpret_cast<const float2*>(s_g + col), be = *reinterpret_cast<const float2*>(s_b + col);
          hid[c][nt][0] = fmaxf(fmaf((hid[c][nt][0] - m0) * r0, gg.x, be.x), 0.f);
          hid[c][nt][1] = fmaxf(fmaf((hid[c][nt][1] - m0) * r0, gg.y, be.y), 0.f);
          hid[c][nt][2] = fmaxf(fmaf((hid[c][nt][2] - m1) * r1, gg.x, be.x), 0.f);
          hid[c][nt][3] = fmaxf(fmaf((hid[c][nt][3] - m1) * r1, gg.y, be.y), 0.f);
        }
    } else {
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int nt = 0; nt < NT_CHUNK; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = hid[c][nt][e];
            hid[c][nt][e] = (v > 20.f ? v : log1pf(expf(v))) - 0.69314718055994531f;
          }
    }
#pragma unroll
    for (int ks = 0; ks < KS2; ++ks) {
      const int c = ks / 4, n0 = (ks % 4) * 2;
      split_bf16x2(hid[c][n0][0], hid[c][n0][1], zhi[ks][0], zlo[ks][0]);
      split_bf16x2(hid[c][n0][2], hid[c][n0][3], zhi[ks][1], zlo[ks][1]);
      split_bf16x2(hid[c][n0 + 1][0], hid[c][n0 + 1][1], zhi[ks][2], zlo[ks][2]);
      split_bf16x2(hid[c][n0 + 1][2], hid[c][n0 + 1][3], zhi[ks][3], zlo[ks][3]);
    }
  }

  // ---- stage 2 ----
  constexpr int NTN = N2T < NT_CHUNK ? N2T : NT_CHUNK;
#pragma unroll 1
  for (int c2 = 0; c2 < N_S2; ++c2, ++job) {
    if (job + 1 < N_JOBS) { issue(job + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const F* bt = (job & 1) ? buf1 : buf0;
    float c[NTN][4];
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
    gemm_tile<X3, KS2, NTN>(c, zhi, zlo, bt, lane);
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt) {
      const int col = c2 * 64 + nt * 8 + 2 * t;
      float2 o0 = make_float2(c[nt][0], c[nt][1]), o1 = make_float2(c[nt][2], c[nt][3]);
      if (MODE == 2) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = col + e;
          if (cc < a.n2_valid) {
            const float bb = a.b2[cc];
            if (ok0) a.out2[(size_t)row0 * a.n2_valid + cc] = (e ? o0.y : o0.x) + bb;
            if (ok1) a.out2[(size_t)row1 * a.n2_valid + cc] = (e ? o1.y : o1.x) + bb;
          }
        }
      } else {
        const float2 bb = *reinterpret_cast<const float2*>(a.b2 + col);
        o0.x += bb.x; o0.y += bb.y; o1.x += bb.x; o1.y += bb.y;
        if (a.residual) {
          if (ok0) { const float2 r = *reinterpret_cast<const float2*>(a.residual + (size_t)row0 * H + col); o0.x += r.x; o0.y += r.y; }
          if (ok1) { const float2 r = *reinterpret_cast<const float2*>(a.residual + (size_t)row1 * H + col); o1.x += r.x; o1.y += r.y; }
        }
        if (ok0) *reinterpret_cast<float2*>(a.out2 + (size_t)row0 * H + col) = o0;
        if (ok1) *reinterpret_cast<float2*>(a.out2 + (size_t)row1 * H + col) = o1;
      }
    }
    __syncthreads();
  }
}

template <int MODE, bool X3>
int launch_mode(const NodeArgs& a, cudaStream_t st) {
  constexpr int KSB = (MODE == 0 ? 160 : 128) / 16;
  constexpr int FB = X3 ? 16 : 8;
  const size_t smem = (size_t)2 * NT_CHUNK * (KSB > KS2 ? KSB : KS2) * 32 * FB + 2 * H * 4;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem(node_mlp_kernel<MODE, X3>, smem, configured)) return rc;
  const int grid = (a.n_atoms + 127) / 128;
  node_mlp_kernel<MODE, X3><<<grid, 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

int launch_node_mlp(const smb_model_dims& d, const NodeArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 0) return 0;
  const bool x3 = d.precision == SMB_PREC_BF16X3;
  switch (a.x_mode) {
    case XMODE_H_INV: return x3 ? launch_mode<0, true>(a, st) : launch_mode<0, false>(a, st);
    case XMODE_AGG_H: return x3 ? launch_mode<1, true>(a, st) : launch_mode<1, false>(a, st);
    default: return x3 ? launch_mode<2, true>(a, st) : launch_mode<2, false>(a, st);
  }
}

}  // namespace smb
