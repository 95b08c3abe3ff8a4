// Fused edge kernels of BaseX2HAttLayer / BaseH2XAttLayer (models/uni_transformer.py:48-162) and of
// the global edge gate (_pred_ew, :475-481), on a dense per-destination neighbour table.
//
// For a destination atom i with neighbour slots j = nbr[i][0..deg):
//   pre_ij = W1r . rbf(|x_i - x_j|)  +  A_i  +  B_j          (first Linear, split: SURVEY 0.6)
//   z_ij   = ReLU(LayerNorm(pre_ij))
//   ROLE_GATE : e_w    = sigmoid(w2 . z + b2)
//   ROLE_K    : logit  = <Q_i^a , (W2 z)^a> / sqrt(dh)   -> softmax over the slots of i (per head)
//                        alpha*e_w written to [N][k+1][16]
//   ROLE_V    : agg_i  = sum_j alpha_ij^a e_w (W2 z + b2)^a                      (X2H messages)
//   ROLE_XV   : o_i^a  = sum_j alpha_ij^a e_w (w2^a . z + b2^a) (x_i - x_j)      (H2X messages)
//               then the VN linear maps of shape_linear (shape_vn_layers.py:100,105) and the
//               per-channel partial sums of the VN batch-norm statistics.
//
// Mapping: persistent CTA (one per SM) loops over molecules; the molecule's A/B projections are
// staged in shared memory; one warp owns one destination: its <=32 neighbour slots are the 32 rows of
// a [32 x 128] tile held in mma.sync accumulator fragments, so LayerNorm is a quad shuffle, the
// per-destination softmax is a warp-local reduction over rows (no atomics, deterministic), and the
// first GEMM's accumulator fragments are re-packed in registers as the A operand of the second GEMM.
// Weights (bf16 hi/lo B fragments) stay resident in shared memory for the CTA's lifetime.
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

namespace {

constexpr int H = 128;
constexpr int NT = H / 8;       // 16 n-tiles
constexpr int KS2 = H / 16;     // 8 k-steps of the second GEMM
constexpr int HP = H + 8;       // padded row stride of the A/B tiles (conflict-free float2 fragment reads)
constexpr int LS = 17;          // padded row stride of the per-warp logit / alpha scratch
constexpr int MAXSLOT = 64;
constexpr int WARPS = kEdgeWarps;
constexpr int MT = 1;          // m-tiles (16 rows) per sub-tile: 1 keeps the warp at <=128 registers (16 warps/SM)
constexpr int R = 16 * MT;     // neighbour slots (rows) per sub-tile
constexpr int NR = 2 * MT;     // rows per thread
constexpr int NGK = 4;         // heads (n-tiles) of the second GEMM in flight: independent MMA chains
constexpr int NGV = 4;

template <bool X3> struct Frag { using type = uint4; };
template <> struct Frag<false> { using type = uint2; };
template <bool X3>
__device__ __forceinline__ uint4 ld_frag(const typename Frag<X3>::type* p) {
  if constexpr (X3) return *p;
  else { const uint2 v = *p; return make_uint4(v.x, v.y, 0u, 0u); }
}

struct SmemPlan {
  size_t w1r, w2, vecs, tiles, xs, warp, vnw, shape, total;
  size_t warp_stride;
  int rows;       // atom rows of the staged group tile
  int mols;       // molecules staged together (a "group")
  int slot_cap;   // neighbour slots held in the per-warp scratch
};
constexpr int GROUP_ROWS = 64;
constexpr int MAX_GROUP_MOLS = 8;
__host__ __device__ inline SmemPlan plan_smem(int role, bool x3, int n_max, int k) {
  const size_t fb = x3 ? 16 : 8;
  SmemPlan p;
  size_t o = 0;
  p.w1r = o; o += (size_t)NT * 2 * 32 * fb;
  p.w2 = o;
  if (role == ROLE_K || role == ROLE_V) o += (size_t)NT * KS2 * 32 * fb;
  else if (role == ROLE_XV) o += (size_t)2 * KS2 * 32 * fb;
  else o += H * 4;
  p.vecs = o; o += 3 * H * 4;                        // ln_g | ln_b | b2 (gate: b1)
  if (n_max < 1) n_max = 1;
  p.mols = GROUP_ROWS / n_max;
  if (p.mols < 1) p.mols = 1;
  if (p.mols > MAX_GROUP_MOLS) p.mols = MAX_GROUP_MOLS;
  p.rows = p.mols * n_max;
  int deg_max = k + 1 < n_max ? k + 1 : n_max;       // deg <= min(k+1, n-1) (k+1 only for coincident duplicates)
  p.slot_cap = (deg_max + 31) / 32 * 32;
  if (p.slot_cap > MAXSLOT) p.slot_cap = MAXSLOT;
  p.tiles = o; if (role != ROLE_GATE) o += (size_t)2 * p.rows * HP * 4;
  p.xs = o; o += (size_t)p.rows * 4 * 4;
  p.vnw = o; if (role == ROLE_XV) o += (size_t)2 * kHeads * kVnStride * 4;
  o = (o + 15) / 16 * 16;
  p.shape = o; if (role == ROLE_XV) o += (size_t)p.mols * kShape * 3 * 4;
  o = (o + 15) / 16 * 16;
  p.warp = o;
  size_t ws = 0;
  if (role == ROLE_K) ws = H * 4 + (size_t)p.slot_cap * LS * 4;
  else if (role == ROLE_V) ws = (size_t)p.slot_cap * LS * 4 + H * 4;
  else if (role == ROLE_XV) ws = (size_t)p.slot_cap * LS * 4 + kHeads * 4 * 4;
  ws = (ws + 15) / 16 * 16;
  p.warp_stride = ws;
  o += ws * WARPS;
  p.total = o;
  return p;
}

template <int ROLE, bool X3>
__global__ void __launch_bounds__(WARPS * 32, 1) edge_kernel(EdgeArgs a) {
  static_assert(WARPS * 32 * 128 <= 65536, "register budget");
  using F = typename Frag<X3>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemPlan P = plan_smem(ROLE, X3, a.n_max, a.k);
  F* s_w1r = reinterpret_cast<F*>(smem + P.w1r);
  F* s_w2 = reinterpret_cast<F*>(smem + P.w2);
  float* s_w2vec = reinterpret_cast<float*>(smem + P.w2);     // gate only
  float* s_g = reinterpret_cast<float*>(smem + P.vecs);
  float* s_be = s_g + H;
  float* s_b2 = s_be + H;
  float* s_A = reinterpret_cast<float*>(smem + P.tiles);
  float* s_B = s_A + (size_t)P.rows * HP;
  float* s_x = reinterpret_cast<float*>(smem + P.xs);
  float* s_vnw = reinterpret_cast<float*>(smem + P.vnw);
  float* s_shape = reinterpret_cast<float*>(smem + P.shape);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  float* s_warp = reinterpret_cast<float*>(smem + P.warp + (size_t)warp * P.warp_stride);
  float* s_q = s_warp;                                        // ROLE_K
  float* s_l = ROLE == ROLE_K ? s_warp + H : s_warp;          // logits (K) / alpha (V, XV)
  float* s_o = s_warp + P.slot_cap * LS;                      // ROLE_XV: o[16][4]
  float* s_out = s_warp + P.slot_cap * LS;                    // ROLE_V: per-destination accumulator [H]
  const int KSTR = a.k + 1;

  // ---- stage the weights once per CTA ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.w1r);
    uint4* dst = reinterpret_cast<uint4*>(s_w1r);
    const int n16 = NT * 2 * 32 * (X3 ? 16 : 8) / 16;
    for (int p = tid; p < n16; p += blockDim.x) dst[p] = src[p];
    if (ROLE == ROLE_GATE) {
      for (int p = tid; p < H; p += blockDim.x) s_w2vec[p] = reinterpret_cast<const float*>(a.w2)[p];
    } else {
      const int nt2 = ROLE == ROLE_XV ? 2 : NT;
      const uint4* s2 = reinterpret_cast<const uint4*>(a.w2);
      uint4* d2 = reinterpret_cast<uint4*>(s_w2);
      const int m16 = nt2 * KS2 * 32 * (X3 ? 16 : 8) / 16;
      for (int p = tid; p < m16; p += blockDim.x) d2[p] = s2[p];
    }
    for (int p = tid; p < H; p += blockDim.x) {
      s_g[p] = a.ln_g[p];
      s_be[p] = a.ln_b[p];
      if (ROLE == ROLE_GATE) s_b2[p] = a.b1[p];
      else if (ROLE == ROLE_XV) s_b2[p] = p < kHeads ? a.b2[p] : 0.f;
      else s_b2[p] = a.b2[p];
    }
    if (ROLE == ROLE_XV) {
      for (int p = tid; p < kHeads * kVnStride; p += blockDim.x) {
        s_vnw[p] = a.vn_feat[p];
        s_vnw[kHeads * kVnStride + p] = a.vn_dir[p];
      }
    }
  }
  const float gate_b2 = ROLE == ROLE_GATE ? a.b2[0] : 0.f;
  float bn_s = 0.f, bn_q = 0.f;   // ROLE_XV: per-warp partial sums (lane = channel)

  const int n_groups = (a.n_mols + P.mols - 1) / P.mols;
#pragma unroll 1
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int m0 = grp * P.mols;
    const int m1 = min(m0 + P.mols, a.n_mols);
    const int a0 = a.mol_ptr[m0];
    const int n = a.mol_ptr[m1] - a0;            // atoms of the group (<= P.rows)
    __syncthreads();   // previous group's tiles are free (also orders the weight staging)
    for (int p = tid; p < n * 3; p += blockDim.x) s_x[(p / 3) * 4 + (p % 3)] = a.x[(size_t)a0 * 3 + p];
    if (ROLE != ROLE_GATE) {
      const int n4 = n * (H / 4);
      for (int p = tid; p < n4; p += blockDim.x) {
        const int row = p / (H / 4), c4 = p % (H / 4);
        const float* src = a.ab + (size_t)(a0 + row) * (4 * H);
        *reinterpret_cast<float4*>(s_A + row * HP + c4 * 4) = *reinterpret_cast<const float4*>(src + a.col_a + c4 * 4);
        *reinterpret_cast<float4*>(s_B + row * HP + c4 * 4) = *reinterpret_cast<const float4*>(src + a.col_b + c4 * 4);
      }
    }
    if (ROLE == ROLE_XV)
      for (int p = tid; p < (m1 - m0) * kShape * 3; p += blockDim.x) s_shape[p] = a.shape[(size_t)m0 * kShape * 3 + p];
    __syncthreads();

#pragma unroll 1
    for (int i = warp; i < n; i += WARPS) {        // i: group-local atom index
      const int gi = a0 + i;
      int mloc = 0;                                 // molecule of atom i within the group
      while (m0 + mloc + 1 < m1 && a.mol_ptr[m0 + mloc + 1] <= gi) ++mloc;
      const int moff = a.mol_ptr[m0 + mloc] - a0;   // group-local index of the molecule's first atom
      const int dg = min(a.deg[gi], P.slot_cap);
      const int nsub = (dg + R - 1) / R;
      const int* nb = a.nbr + (size_t)gi * KSTR;
      const float xi = s_x[i * 4], yi = s_x[i * 4 + 1], zi = s_x[i * 4 + 2];

      if (ROLE == ROLE_K) {
        *reinterpret_cast<float4*>(s_q + lane * 4) = *reinterpret_cast<const float4*>(a.q + (size_t)gi * H + lane * 4);
      }
      if (ROLE == ROLE_V || ROLE == ROLE_XV) {
        const float* al = a.alpha + (size_t)gi * KSTR * kHeads;
        for (int p = lane; p < nsub * R * kHeads; p += 32) {
          const int slot = p >> 4, hd = p & 15;
          s_l[slot * LS + hd] = slot < dg ? al[p] : 0.f;
        }
      }
      if (ROLE == ROLE_V) *reinterpret_cast<float4*>(s_out + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();

      float oacc[ROLE == ROLE_XV ? 4 : 1][3];
      if (ROLE == ROLE_XV) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { oacc[q][0] = 0.f; oacc[q][1] = 0.f; oacc[q][2] = 0.f; }
      }

#pragma unroll 1
      for (int s = 0; s < nsub; ++s) {
        // ---- rows of this thread: slot = R s + g + 8 r2 ----
        int jr[NR];
        float rel[NR][3], dist[NR];
#pragma unroll
        for (int r2 = 0; r2 < NR; ++r2) {
          const int slot = s * R + g + 8 * r2;
          const int j = slot < dg ? nb[slot] + moff : -1;
          jr[r2] = j;
          const int jj = j < 0 ? 0 : j;
          rel[r2][0] = xi - s_x[jj * 4]; rel[r2][1] = yi - s_x[jj * 4 + 1]; rel[r2][2] = zi - s_x[jj * 4 + 2];
          dist[r2] = sqrtf(rel[r2][0] * rel[r2][0] + rel[r2][1] * rel[r2][1] + rel[r2][2] * rel[r2][2]);
        }
        // ---- A fragments of GEMM1: rbf(dist), K = 20 padded to 32 ----
        uint32_t a1hi[MT][2][4], a1lo[MT][2][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r2 = mt * 2 + (e & 1);
              const int k0 = ks * 16 + 2 * t + ((e >> 1) ? 8 : 0);
              float v0 = 0.f, v1 = 0.f;
              if (k0 < kRbf && jr[r2] >= 0) {
                // exp(-0.5 d^2) = 2^(-0.5 log2(e) d^2)
                const float d0 = dist[r2] - rbf_centre(k0), d1 = dist[r2] - rbf_centre(k0 + 1);
                v0 = exp2f(-0.72134752044448170f * d0 * d0);
                v1 = exp2f(-0.72134752044448170f * d1 * d1);
              }
              split_bf16x2(v0, v1, a1hi[mt][ks][e], a1lo[mt][ks][e]);
            }
        // ---- GEMM1 (4 n-tiles = 4*MT independent accumulator chains in flight) ----
        float acc[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
        for (int nb4 = 0; nb4 < NT; nb4 += 4) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            uint4 bw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) bw[q] = ld_frag<X3>(s_w1r + ((nb4 + q) * 2 + ks) * 32 + lane);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              mma_step_n<X3, 4>(*reinterpret_cast<float(*)[4][4]>(&acc[mt][nb4]), a1hi[mt][ks], a1lo[mt][ks], bw);
          }
        }
        // ---- + A_i + B_j (gate: + b1), LayerNorm statistics ----
        float sum[NR];
#pragma unroll
        for (int r2 = 0; r2 < NR; ++r2) sum[r2] = 0.f;
        {
          const float* Ai = s_A + i * HP + 2 * t;
          const float* Bj[NR];
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) Bj[r2] = s_B + (jr[r2] < 0 ? 0 : jr[r2]) * HP + 2 * t;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float2 av;
            if (ROLE == ROLE_GATE) av = *reinterpret_cast<const float2*>(s_b2 + nt * 8 + 2 * t);
            else av = *reinterpret_cast<const float2*>(Ai + nt * 8);
#pragma unroll
            for (int r2 = 0; r2 < NR; ++r2) {
              float2 bv = make_float2(0.f, 0.f);
              if (ROLE != ROLE_GATE) bv = *reinterpret_cast<const float2*>(Bj[r2] + nt * 8);
              float& c0 = acc[r2 >> 1][nt][(r2 & 1) * 2];
              float& c1 = acc[r2 >> 1][nt][(r2 & 1) * 2 + 1];
              c0 += av.x + bv.x; c1 += av.y + bv.y;
              sum[r2] += c0 + c1;
            }
          }
        }
        float mean[NR], rstd[NR];
#pragma unroll
        for (int r2 = 0; r2 < NR; ++r2) mean[r2] = quad_sum(sum[r2]) * (1.f / H);
        {
          float var[NR];
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) var[r2] = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r2 = 0; r2 < NR; ++r2) {
              const float d0 = acc[r2 >> 1][nt][(r2 & 1) * 2] - mean[r2];
              const float d1 = acc[r2 >> 1][nt][(r2 & 1) * 2 + 1] - mean[r2];
              var[r2] = fmaf(d0, d0, var[r2]); var[r2] = fmaf(d1, d1, var[r2]);
            }
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) rstd[r2] = 1.f / sqrtf(quad_sum(var[r2]) * (1.f / H) + 1e-5f);
        }
        // ---- normalise + affine + ReLU (in place) ----
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float2 gg = *reinterpret_cast<const float2*>(s_g + nt * 8 + 2 * t);
          const float2 be = *reinterpret_cast<const float2*>(s_be + nt * 8 + 2 * t);
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) {
            float& c0 = acc[r2 >> 1][nt][(r2 & 1) * 2];
            float& c1 = acc[r2 >> 1][nt][(r2 & 1) * 2 + 1];
            c0 = fmaxf(fmaf((c0 - mean[r2]) * rstd[r2], gg.x, be.x), 0.f);
            c1 = fmaxf(fmaf((c1 - mean[r2]) * rstd[r2], gg.y, be.y), 0.f);
          }
        }

        if (ROLE == ROLE_GATE) {
          // ---- e_w = sigmoid(w2 . z + b2) ----
          float dot[NR];
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) dot[r2] = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float2 w = *reinterpret_cast<const float2*>(s_w2vec + nt * 8 + 2 * t);
#pragma unroll
            for (int r2 = 0; r2 < NR; ++r2)
              dot[r2] = fmaf(acc[r2 >> 1][nt][(r2 & 1) * 2], w.x, fmaf(acc[r2 >> 1][nt][(r2 & 1) * 2 + 1], w.y, dot[r2]));
          }
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) {
            dot[r2] = quad_sum(dot[r2]);
            const int slot = s * R + g + 8 * r2;
            if (t == (r2 & 3) && slot < dg) a.ew_out[(size_t)gi * KSTR + slot] = 1.f / (1.f + expf(-(dot[r2] + gate_b2)));
          }
        } else {
          // ---- re-pack z as A fragments of GEMM2 (accumulator layout == A layout, in registers) ----
          uint32_t zhi[MT][KS2][4], zlo[MT][KS2][4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int ks = 0; ks < KS2; ++ks) {
              split_bf16x2(acc[mt][2 * ks][0], acc[mt][2 * ks][1], zhi[mt][ks][0], zlo[mt][ks][0]);
              split_bf16x2(acc[mt][2 * ks][2], acc[mt][2 * ks][3], zhi[mt][ks][1], zlo[mt][ks][1]);
              split_bf16x2(acc[mt][2 * ks + 1][0], acc[mt][2 * ks + 1][1], zhi[mt][ks][2], zlo[mt][ks][2]);
              split_bf16x2(acc[mt][2 * ks + 1][2], acc[mt][2 * ks + 1][3], zhi[mt][ks][3], zlo[mt][ks][3]);
            }

          if (ROLE == ROLE_K || ROLE == ROLE_V) {
            constexpr int NG = ROLE == ROLE_K ? NGK : NGV;
#pragma unroll 1
            for (int nb2 = 0; nb2 < NT; nb2 += NG) {   // one n-tile == one head (dh = 8); NG heads in flight
              float c[MT][NG][4];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int q = 0; q < NG; ++q)
#pragma unroll
                  for (int e = 0; e < 4; ++e) c[mt][q][e] = 0.f;
#pragma unroll
              for (int ks = 0; ks < KS2; ++ks) {
                uint4 bw[NG];
#pragma unroll
                for (int q = 0; q < NG; ++q) bw[q] = ld_frag<X3>(s_w2 + ((nb2 + q) * KS2 + ks) * 32 + lane);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_step_n<X3, NG>(c[mt], zhi[mt][ks], zlo[mt][ks], bw);
              }
              if (ROLE == ROLE_K) {
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                  const float2 qv = *reinterpret_cast<const float2*>(s_q + (nb2 + q) * 8 + 2 * t);
                  // b2 shifts every logit of (i, head) by the same <Q_i, b2>: softmax-invariant, dropped
#pragma unroll
                  for (int r2 = 0; r2 < NR; ++r2) {
                    const float l = quad_sum(fmaf(c[r2 >> 1][q][(r2 & 1) * 2], qv.x, c[r2 >> 1][q][(r2 & 1) * 2 + 1] * qv.y));
                    if (t == (r2 & 3)) s_l[(s * R + g + 8 * r2) * LS + nb2 + q] = l;
                  }
                }
              } else {
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                  const int nt2 = nb2 + q;
                  const float2 bb = *reinterpret_cast<const float2*>(s_b2 + nt2 * 8 + 2 * t);
                  float v0 = 0.f, v1 = 0.f;
#pragma unroll
                  for (int r2 = 0; r2 < NR; ++r2) {
                    const float al = s_l[(s * R + g + 8 * r2) * LS + nt2];
                    v0 = fmaf(al, c[r2 >> 1][q][(r2 & 1) * 2] + bb.x, v0);
                    v1 = fmaf(al, c[r2 >> 1][q][(r2 & 1) * 2 + 1] + bb.y, v1);
                  }
                  v0 = group_sum(v0); v1 = group_sum(v1);
                  if (g == 0) {
                    float2* o = reinterpret_cast<float2*>(s_out + nt2 * 8 + 2 * t);
                    float2 cur = *o;
                    cur.x += v0; cur.y += v1;
                    *o = cur;
                  }
                }
              }
            }
          } else {   // ROLE_XV: N2 = 16 heads = 2 n-tiles
            float c[MT][2][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) c[mt][q][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS2; ++ks) {
              uint4 bw[2];
#pragma unroll
              for (int q = 0; q < 2; ++q) bw[q] = ld_frag<X3>(s_w2 + (q * KS2 + ks) * 32 + lane);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) mma_step_n<X3, 2>(c[mt], zhi[mt][ks], zlo[mt][ks], bw);
            }
#pragma unroll
            for (int nt2 = 0; nt2 < 2; ++nt2) {
              const float2 bb = *reinterpret_cast<const float2*>(s_b2 + nt2 * 8 + 2 * t);
#pragma unroll
              for (int r2 = 0; r2 < NR; ++r2) {
                const float* al = s_l + (s * R + g + 8 * r2) * LS + nt2 * 8 + 2 * t;
                const float w0 = al[0] * (c[r2 >> 1][nt2][(r2 & 1) * 2] + bb.x);
                const float w1 = al[1] * (c[r2 >> 1][nt2][(r2 & 1) * 2 + 1] + bb.y);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  oacc[nt2 * 2][d] = fmaf(w0, rel[r2][d], oacc[nt2 * 2][d]);
                  oacc[nt2 * 2 + 1][d] = fmaf(w1, rel[r2][d], oacc[nt2 * 2 + 1][d]);
                }
              }
            }
          }
        }
      }   // sub-tiles

      // ---- per-destination epilogues ----
      if (ROLE == ROLE_K) {
        __syncwarp();
        const int hd = lane & 15, part = lane >> 4;
        const float scale = 0.35355339059327373f;   // 1/sqrt(dh), dh = 8
        float mx = -INFINITY;
        for (int slot = part; slot < dg; slot += 2) mx = fmaxf(mx, s_l[slot * LS + hd] * scale);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        float se = 0.f;
        for (int slot = part; slot < dg; slot += 2) {
          const float e = expf(s_l[slot * LS + hd] * scale - mx);
          s_l[slot * LS + hd] = e;
          se += e;
        }
        se += __shfl_xor_sync(0xffffffffu, se, 16);
        const float inv = 1.f / se;
        float* al = a.alpha + (size_t)gi * KSTR * kHeads;
        const float* ew = a.ew_in + (size_t)gi * KSTR;
        for (int slot = part; slot < dg; slot += 2) al[slot * kHeads + hd] = s_l[slot * LS + hd] * inv * ew[slot];
        __syncwarp();
      } else if (ROLE == ROLE_V) {
        __syncwarp();
        *reinterpret_cast<float4*>(a.agg + (size_t)gi * H + lane * 4) = *reinterpret_cast<const float4*>(s_out + lane * 4);
        __syncwarp();
      } else if (ROLE == ROLE_XV) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float v = group_sum(oacc[q][d]);
            // head of (q): n-tile q>>1, column 2t + (q&1)
            if (g == 0) s_o[((q >> 1) * 8 + 2 * t + (q & 1)) * 4 + d] = v;
          }
        __syncwarp();
        // VN linear maps: lanes 0..15 -> map_to_feat channel, lanes 16..31 -> map_to_dir channel
        const int ch = lane & 15, which = lane >> 4;
        const float* w = s_vnw + (which * kHeads + ch) * kVnStride;
        const float* shp = s_shape + mloc * kShape * 3;
        float vx = w[0] * xi, vy = w[0] * yi, vz = w[0] * zi;
#pragma unroll
        for (int c = 0; c < kHeads; ++c) {
          const float wc = w[1 + c];
          vx = fmaf(wc, s_o[c * 4], vx); vy = fmaf(wc, s_o[c * 4 + 1], vy); vz = fmaf(wc, s_o[c * 4 + 2], vz);
        }
#pragma unroll 8
        for (int c = 0; c < kShape; ++c) {
          const float wc = w[1 + kHeads + c];
          vx = fmaf(wc, shp[c * 3], vx); vy = fmaf(wc, shp[c * 3 + 1], vy); vz = fmaf(wc, shp[c * 3 + 2], vz);
        }
        float* row = a.vn + (size_t)gi * kVnRow;
        row[3 + which * 48 + ch * 3] = vx; row[4 + which * 48 + ch * 3] = vy; row[5 + which * 48 + ch * 3] = vz;
        if (lane < 3) {
          float sm = 0.f;
#pragma unroll
          for (int c = 0; c < kHeads; ++c) sm += s_o[c * 4 + lane];
          row[lane] = sm * (1.f / kHeads);
        }
        if (which == 0) {
          const float nu = sqrtf(vx * vx + vy * vy + vz * vz) + 1e-6f;
          bn_s += nu; bn_q = fmaf(nu, nu, bn_q);
        }
        __syncwarp();
      }
    }   // destinations
  }     // groups

  if (ROLE == ROLE_XV && lane < 16) {
    float* part = a.bn_partial + (size_t)(blockIdx.x * WARPS + warp) * 32;
    part[lane] = bn_s;
    part[16 + lane] = bn_q;
  }
}


template <int ROLE, bool X3>
int launch_role(const EdgeArgs& a, int* bn_rows_out, cudaStream_t st) {
  const SmemPlan P = plan_smem(ROLE, X3, a.n_max, a.k);
  if (P.total > 227 * 1024) { set_error_msg("edge kernel: shared memory plan exceeds 227 KB"); return SMB_E_TOOBIG; }
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem(edge_kernel<ROLE, X3>, P.total, configured)) return rc;
  int grid = device_sm_count();
  if (grid > kEdgeMaxCtas) grid = kEdgeMaxCtas;
  const int n_groups = (a.n_mols + P.mols - 1) / P.mols;
  if (grid > n_groups) grid = n_groups;
  if (bn_rows_out) *bn_rows_out = grid * WARPS;
  edge_kernel<ROLE, X3><<<grid, WARPS * 32, P.total, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

int launch_edge(const smb_model_dims& d, int role, const EdgeArgs& a, int* grid_out, cudaStream_t st) {
  if (a.n_mols <= 0) { if (grid_out) *grid_out = 0; return 0; }
  const bool x3 = d.precision == SMB_PREC_BF16X3;
  switch (role) {
    case ROLE_GATE: return x3 ? launch_role<ROLE_GATE, true>(a, grid_out, st) : launch_role<ROLE_GATE, false>(a, grid_out, st);
    case ROLE_K: return x3 ? launch_role<ROLE_K, true>(a, grid_out, st) : launch_role<ROLE_K, false>(a, grid_out, st);
    case ROLE_V: return x3 ? launch_role<ROLE_V, true>(a, grid_out, st) : launch_role<ROLE_V, false>(a, grid_out, st);
    default: return x3 ? launch_role<ROLE_XV, true>(a, grid_out, st) : launch_role<ROLE_XV, false>(a, grid_out, st);
  }
}

}  // namespace smb
