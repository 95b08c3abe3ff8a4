// VN-DGCNN shape encoder (reference models/shape_pointcloud_modelAE.py:207-255 with the layers of
// models/shape_vn_layers.py:41-110,257-292), once per condition shape.
//
// Algebra used (exact; the reference materialises [B,2C,3,P,k] edge features instead):
//   edge feature of (i, j) = [f_j - f_i | f_i]  =>  W [f_j - f_i | f_i] = Wa f_j + (Wb - Wa) f_i
//   so per layer ONE node-level GEMM produces  U = Wa f, V = (Wb - Wa) f, Ud, Vd  (feat and dir maps)
//   and the per-edge work is  p_ij = U_j + V_i,  d_ij = Ud_j + Vd_i  followed by the VN batch-norm /
//   leaky-ReLU and the mean over the k neighbours.
//
// Layout: activations are [point][3][channels] (channel contiguous), i.e. a point's feature vector is
// three contiguous 128-float segments; the four block outputs are written side by side into one
// [point][3][128*n_blocks] buffer, which is the concatenation conv_c consumes.
//
// Kernels (fp32 data; the contractions run on tcgen05 with split-bf16 products, ~2^-16 per product, smb_tc_gemm.cu):
//   enc_sqnorm_kernel      |f|^2 per point
//   tc_gemm_kernel         Gram matrices F F^T of the hidden features, one cloud per grid.z (P <= 1024)
//   enc_topk_kernel        pd = -|fj|^2 + 2<fi,fj> - |fi|^2 from the Gram row, held in registers -> top-k per row
//   enc_knn_kernel         fp32 SIMT Gram tile + top-k: the coordinate graph of conv_pos (K = 3) and clouds of P > 1024
//   tc_gemm_kernel         node GEMM  [3*B*P x K] x [N x K]^T
//   enc_edge_stats_kernel  sum / sum of squares of |p_ij| per channel (BatchNorm2d batch statistics)
//   enc_bn_final_kernel    fp64 reduction of the partials, running-stat update, scale | shift
//   enc_edge_apply_kernel  VN batch-norm + directional leaky-ReLU + mean over k
//   enc_c_stats / enc_c_apply   conv_c (shared direction, BatchNorm1d) + mean over the points
#include <math.h>
#include <string.h>

#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {
namespace enc {

constexpr int HS = 128;          // vector channels of every hidden layer
constexpr int UVW = 4 * HS;      // U | V | Ud | Vd
constexpr int PCW = 64;          // row stride of the conv_c projection (latent feat channels | shared dir | pad)
constexpr float VN_EPS = 1e-6f;  // EPS of models/shape_vn_layers.py:6
constexpr int STAT_CTAS = 148 * 8;

struct Ws {
  size_t xx, idx, h0, hc, uv, pc, w4, wc, part, bnp, gram, wimg, wimg_bytes, total;
};
constexpr int GRAM_MAX_P = 1024;   // clouds up to this size take the tensor-core Gram path (the top-k holds a row in registers)

static Ws plan(int n_blocks, int latent, int k, size_t n_points, int P) {
  (void)latent;
  Ws w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  w.xx = take(n_points * 4);
  w.idx = take(n_points * k * 4);
  w.h0 = take(n_points * 3 * HS * 4);
  w.hc = take(n_points * 3 * HS * n_blocks * 4);
  w.uv = take(n_points * 3 * UVW * 4);
  w.pc = take(n_points * 3 * PCW * 4);
  w.w4 = take((size_t)UVW * HS * 4);
  w.wc = take((size_t)PCW * HS * n_blocks * 4);
  w.part = take((size_t)STAT_CTAS * 2 * HS * 8);
  w.bnp = take(2 * HS * 4);
  w.gram = take(P <= GRAM_MAX_P ? n_points * (size_t)P * 4 : 0);   // [B][P][P] fp32
  {   // pre-split weight image of the node GEMM in flight (smb_tc_gemm.cu)
    const int k_blk[1] = {HS}, k_c[1] = {HS * n_blocks};
    const long long a = tc_gemm_w_img_bytes(UVW, k_blk, 1, true), b = tc_gemm_w_img_bytes(PCW, k_c, 1, true);
    w.wimg_bytes = (size_t)(a > b ? a : b);
    w.wimg = take(w.wimg_bytes);
  }
  w.total = o;
  return w;
}

// ---- |f|^2 ------------------------------------------------------------------------------------
// one warp per point; feature vector = segs segments of seg_len floats, seg_stride apart
__global__ void __launch_bounds__(256) enc_sqnorm_kernel(const float* __restrict__ feat, size_t row_stride, int segs, int seg_len,
                                                         int seg_stride, size_t n_points, float* __restrict__ xx) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (warp >= n_points) return;
  const float* f = feat + warp * row_stride;
  float s = 0.f;
  for (int sg = 0; sg < segs; ++sg)
    for (int c = lane; c < seg_len; c += 32) { const float v = f[(size_t)sg * seg_stride + c]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) xx[warp] = s;
}

// ---- kNN on the Gram matrix (knn(), shape_vn_layers.py:286-292; the point itself is included) ----
// CTA = (cloud, block of R = 8*RM query rows); pd row block kept in shared memory, then one warp per
// row extracts the k largest entries (ties: smaller index first).
template <int RM>
__global__ void __launch_bounds__(256) enc_knn_kernel(const float* __restrict__ feat, size_t row_stride, int segs, int seg_len,
                                                      int seg_stride, const float* __restrict__ xx, int P, int k,
                                                      int* __restrict__ idx) {
  constexpr int R = 8 * RM;
  constexpr int BSTR = 132;
  extern __shared__ __align__(16) float sm[];
  const int Ppad = (P + 127) / 128 * 128;
  float* pd = sm;                      // [R][Ppad]
  float* As = pd + (size_t)R * Ppad;   // [32][R + 1]
  float* Bs = As + 32 * (R + 1);       // [32][BSTR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blocks_per_cloud = (P + R - 1) / R;
  const int b = blockIdx.x / blocks_per_cloud, row0 = (blockIdx.x % blocks_per_cloud) * R;
  const size_t base = (size_t)b * P;
  const int tr = warp, tc = lane;      // thread tile: rows tr*RM.., cols tc*4..

  for (int ct = 0; ct < Ppad / 128; ++ct) {
    float acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int sg = 0; sg < segs; ++sg)
      for (int k0 = 0; k0 < seg_len; k0 += 32) {
        const int kk = lane;
        const bool kv = k0 + kk < seg_len;
        const size_t koff = (size_t)sg * seg_stride + k0 + kk;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const int r = warp + 8 * i, row = row0 + r;
          As[kk * (R + 1) + r] = (kv && row < P) ? feat[(base + row) * row_stride + koff] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = warp + 8 * i, col = ct * 128 + c;
          Bs[kk * BSTR + c] = (kv && col < P) ? feat[(base + col) * row_stride + koff] : 0.f;
        }
        __syncthreads();
        const int kn = min(32, seg_len - k0);
        for (int q = 0; q < kn; ++q) {
          const float4 bv = *reinterpret_cast<const float4*>(Bs + q * BSTR + tc * 4);
#pragma unroll
          for (int i = 0; i < RM; ++i) {
            const float av = As[q * (R + 1) + tr * RM + i];
            acc[i][0] = fmaf(av, bv.x, acc[i][0]); acc[i][1] = fmaf(av, bv.y, acc[i][1]);
            acc[i][2] = fmaf(av, bv.z, acc[i][2]); acc[i][3] = fmaf(av, bv.w, acc[i][3]);
          }
        }
      }
    // pd[i][j] = -xx[j] - (-2 <fi,fj>) - xx[i]   (same association as the reference expression)
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int r = tr * RM + i, row = row0 + r;
      const float xr = row < P ? xx[base + row] : 0.f;
      float4 o;
      float* op = &o.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = ct * 128 + tc * 4 + j;
        op[j] = col < P ? (-xx[base + col] - (-2.f * acc[i][j])) - xr : -INFINITY;
      }
      *reinterpret_cast<float4*>(pd + (size_t)r * Ppad + ct * 128 + tc * 4) = o;
    }
  }
  __syncthreads();
  // ---- top-k per row ----
  for (int r = warp; r < R; r += 8) {
    const int row = row0 + r;
    if (row >= P) break;
    float* prow = pd + (size_t)r * Ppad;
    for (int s = 0; s < k; ++s) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < Ppad; c += 32) {
        const float v = prow[c];
        if (v > best) { best = v; bi = c; }     // ascending scan keeps the smallest index among equals
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (bi >= P) bi = row;                    // fewer than k finite candidates (never for k <= P)
      if (lane == 0) { idx[(base + row) * k + s] = bi; prow[bi] = -INFINITY; }
      __syncwarp();
    }
  }
}

// ---- node GEMM: C[M][ldc](N) = A[M][lda](K) * W[N][ldw](K)^T on tcgen05 (split-bf16 products, smb_tc_gemm.cu) ----------
static int enc_gemm(const float* A, int lda, const float* W, int ldw, float* C, int ldc, size_t M, int N, int K, void* wimg, size_t wimg_bytes,
                    cudaStream_t st) {
  TcGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.seg[0].a = A; g.seg[0].lda = lda; g.seg[0].k = K;
  g.n_segs = 1; g.W = W; g.ldw = ldw; g.M = (int)M; g.N = N; g.C = C; g.ldc = ldc;
  g.split3 = 1;   // fp32-level products: the next layer's kNN graph is selected on these features
  g.w_img = wimg; g.w_img_bytes = (long long)wimg_bytes;
  return launch_tc_gemm(g, 1, st);
}

// ---- kNN from a Gram matrix (P <= GRAM_MAX_P): G = F F^T per cloud on tcgen05, then one warp per row keeps the row's
// pd[j] = (-|f_j|^2 + 2 G_ij) - |f_i|^2 in registers and extracts the K2 = k + 8 largest (ties: smaller index first).
// The three-piece Gram entries are fp32-level (operand split ~2^-23, fp32 accumulation over K = 384); whenever the k-th and
// (k+1)-th candidates are closer than 3e-5 of |f_i|^2 (exact ties included) the K2 candidates are re-evaluated with an fp64-accumulated dot product of the fp32
// features and the k nearest are selected from those values. ----
template <int NV>
__global__ void __launch_bounds__(256) enc_topk_kernel(const float* __restrict__ gram, const float* __restrict__ xx,
                                                       const float* __restrict__ feat, size_t row_stride, int segs, int seg_len,
                                                       int seg_stride, int P, int k, size_t n_points, int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_points) return;
  const size_t base = row / P * P;
  const float* g = gram + row * P;
  const float xr = xx[row];
  float v[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const int c = q * 32 + lane;
    v[q] = c < P ? (-xx[base + c] - (-2.f * g[c])) - xr : -INFINITY;     // same association as the reference expression
  }
  const int K2 = min(min(k + 8, 32), P);
  int my_idx = 0x7fffffff;
  float my_pd = -INFINITY;
  for (int s = 0; s < K2; ++s) {
    float best = -INFINITY;
    int bq = 0;
#pragma unroll
    for (int q = 0; q < NV; ++q)
      if (v[q] > best) { best = v[q]; bq = q; }      // ascending scan keeps the smallest index among equals
    int bi = best > -INFINITY ? bq * 32 + lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (bi >= P) bi = (int)(row - base);          // fewer than K2 finite candidates (never for K2 <= P)
    if (lane == s) { my_idx = bi; my_pd = best; }
    if ((bi & 31) == lane) {
#pragma unroll
      for (int q = 0; q < NV; ++q)
        if (q == (bi >> 5)) v[q] = -INFINITY;
    }
  }
  // unambiguous boundary: the approximate order of the first k is the answer
  const float tk = __shfl_sync(0xffffffffu, my_pd, k - 1), tk1 = K2 > k ? __shfl_sync(0xffffffffu, my_pd, k) : -INFINITY;
  const float tol = 3e-5f * (fabsf(tk) + fabsf(xr) + 1e-30f);
  if (!(tk - tk1 <= tol)) {
    if (lane < k) idx[row * k + lane] = my_idx;
    return;
  }
  // exact re-evaluation of the K2 candidates
  const float* fi = feat + row * row_stride;
  for (int c = 0; c < K2; ++c) {
    const int j = __shfl_sync(0xffffffffu, my_idx, c);
    const float* fj = feat + (base + j) * row_stride;
    double acc = 0.0;
    for (int sg = 0; sg < segs; ++sg)
      for (int e = lane; e < seg_len; e += 32)
        acc = fma((double)fi[(size_t)sg * seg_stride + e], (double)fj[(size_t)sg * seg_stride + e], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float ex = (-xx[base + j] - (-2.f * (float)acc)) - xr;
    if (lane == c) my_pd = ex;
  }
  if (lane >= K2) { my_pd = -INFINITY; my_idx = 0x7fffffff; }
  for (int s = 0; s < k; ++s) {
    float best = my_pd;
    int bi = my_idx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) idx[row * k + s] = bi;
    if (my_idx == bi) my_pd = -INFINITY;
  }
}

// ---- weight preparation ----------------------------------------------------------------------------
// W4[0:128] = Wa, [128:256] = Wb - Wa, [256:384] = Wda, [384:512] = Wdb - Wda   (W = [Wa | Wb], [128][2*cin])
__global__ void enc_prep_w4_kernel(const float* __restrict__ feat, const float* __restrict__ dir, int cin, float* __restrict__ w4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HS * cin) return;
  const int o = i / cin, c = i % cin;
  const float fa = feat[(size_t)o * 2 * cin + c], fb = feat[(size_t)o * 2 * cin + cin + c];
  const float da = dir[(size_t)o * 2 * cin + c], db = dir[(size_t)o * 2 * cin + cin + c];
  w4[(size_t)o * cin + c] = fa;
  w4[(size_t)(HS + o) * cin + c] = fb - fa;
  w4[(size_t)(2 * HS + o) * cin + c] = da;
  w4[(size_t)(3 * HS + o) * cin + c] = db - da;
}
// conv_pos has one input vector channel (the coordinate): UV[pt][d][c] = x[pt][d] * w4[c]
__global__ void __launch_bounds__(256) enc_pos_uv_kernel(const float* __restrict__ clouds, const float* __restrict__ w4,
                                                         size_t n_rows, float* __restrict__ uv) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of one (point, d) row
  if (i >= n_rows * (UVW / 4)) return;
  const size_t row = i / (UVW / 4);
  const int c4 = (int)(i % (UVW / 4));
  const float x = clouds[row];
  const float4 w = *reinterpret_cast<const float4*>(w4 + c4 * 4);
  *reinterpret_cast<float4*>(uv + row * UVW + c4 * 4) = make_float4(x * w.x, x * w.y, x * w.z, x * w.w);
}

// ---- VNLinearLeakyReLU(dim=5) on the dynamic graph ----------------------------------------------------
__global__ void __launch_bounds__(HS) enc_edge_stats_kernel(const float* __restrict__ uv, const int* __restrict__ idx, int P, int k,
                                                            size_t n_points, double* __restrict__ part) {
  const int o = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (size_t pt = blockIdx.x; pt < n_points; pt += gridDim.x) {
    const size_t cloud0 = pt / P * P;
    const float* vi = uv + pt * 3 * UVW + HS;
    const float v0 = vi[o], v1 = vi[UVW + o], v2 = vi[2 * UVW + o];
    const int* nb = idx + pt * k;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int s = 0; s < k; ++s) {
      const float* uj = uv + (cloud0 + nb[s]) * 3 * UVW;
      const float p0 = uj[o] + v0, p1 = uj[UVW + o] + v1, p2 = uj[2 * UVW + o] + v2;
      const float nrm = sqrtf(p0 * p0 + p1 * p1 + p2 * p2) + VN_EPS;
      a1 += nrm;
      a2 = fmaf(nrm, nrm, a2);
    }
    s1 += (double)a1;
    s2 += (double)a2;
  }
  part[(size_t)blockIdx.x * 2 * HS + o] = s1;
  part[(size_t)blockIdx.x * 2 * HS + HS + o] = s2;
}

// BatchNorm over every axis but the channel: batch statistics (biased variance for the normalisation,
// unbiased for the running estimate, momentum 0.1) or the running statistics.
__global__ void enc_bn_final_kernel(const double* __restrict__ part, int rows, int channels, double count, int use_batch,
                                    const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ rm,
                                    float* __restrict__ rv, float* __restrict__ param) {
  const int o = threadIdx.x;
  if (o >= channels) return;
  float mean, var;
  if (use_batch) {
    double s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < rows; ++r) { s1 += part[(size_t)r * 2 * channels + o]; s2 += part[(size_t)r * 2 * channels + channels + o]; }
    const double m = s1 / count;
    double v = s2 / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    if (rm && rv) {
      const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
      rm[o] = 0.9f * rm[o] + 0.1f * mean;
      rv[o] = 0.9f * rv[o] + 0.1f * (float)unb;
    }
  } else {
    mean = rm[o]; var = rv[o];
  }
  const float scale = w[o] / sqrtf(var + 1e-5f);
  param[o] = scale;
  param[channels + o] = bias[o] - mean * scale;
}

__device__ __forceinline__ void vn_leaky(float& p0, float& p1, float& p2, float d0, float d1, float d2) {
  // shape_vn_layers.py:106-109, negative_slope 0.2
  const float dot = p0 * d0 + p1 * d1 + p2 * d2;
  if (dot < 0.f) {
    const float dn = d0 * d0 + d1 * d1 + d2 * d2;
    const float c = dot / (dn + VN_EPS);
    p0 = 0.2f * p0 + 0.8f * (p0 - c * d0);
    p1 = 0.2f * p1 + 0.8f * (p1 - c * d1);
    p2 = 0.2f * p2 + 0.8f * (p2 - c * d2);
  } else {
    p0 = 0.2f * p0 + 0.8f * p0; p1 = 0.2f * p1 + 0.8f * p1; p2 = 0.2f * p2 + 0.8f * p2;
  }
}

__global__ void __launch_bounds__(HS) enc_edge_apply_kernel(const float* __restrict__ uv, const int* __restrict__ idx, int P, int k,
                                                            size_t n_points, const float* __restrict__ param,
                                                            float* __restrict__ out, int ldo) {
  const int o = threadIdx.x;
  const float scale = param[o], shift = param[HS + o];
  const float inv_k = 1.f / (float)k;
  for (size_t pt = blockIdx.x; pt < n_points; pt += gridDim.x) {
    const size_t cloud0 = pt / P * P;
    const float* ci = uv + pt * 3 * UVW;
    const float v0 = ci[HS + o], v1 = ci[UVW + HS + o], v2 = ci[2 * UVW + HS + o];
    const float e0 = ci[3 * HS + o], e1 = ci[UVW + 3 * HS + o], e2 = ci[2 * UVW + 3 * HS + o];
    const int* nb = idx + pt * k;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int s = 0; s < k; ++s) {
      const float* uj = uv + (cloud0 + nb[s]) * 3 * UVW;
      float p0 = uj[o] + v0, p1 = uj[UVW + o] + v1, p2 = uj[2 * UVW + o] + v2;
      const float d0 = uj[2 * HS + o] + e0, d1 = uj[UVW + 2 * HS + o] + e1, d2 = uj[2 * UVW + 2 * HS + o] + e2;
      const float nrm = sqrtf(p0 * p0 + p1 * p1 + p2 * p2) + VN_EPS;
      const float f = (nrm * scale + shift) / nrm;
      p0 *= f; p1 *= f; p2 *= f;
      vn_leaky(p0, p1, p2, d0, d1, d2);
      a0 += p0; a1 += p1; a2 += p2;
    }
    float* op = out + pt * 3 * ldo;
    op[o] = a0 * inv_k; op[ldo + o] = a1 * inv_k; op[2 * ldo + o] = a2 * inv_k;
  }
}

// ---- conv_c: VNLinearLeakyReLU(dim=4, share_nonlinearity) + mean over the points -------------------------
__global__ void __launch_bounds__(256) enc_c_stats_kernel(const float* __restrict__ pc, int latent, size_t n_points,
                                                          double* __restrict__ part) {
  __shared__ double red[2][8][32];
  const int o = threadIdx.x & 31, pl = threadIdx.x >> 5;
  double s1 = 0.0, s2 = 0.0;
  if (o < latent)
    for (size_t pt = (size_t)blockIdx.x * 8 + pl; pt < n_points; pt += (size_t)gridDim.x * 8) {
      const float* r = pc + pt * 3 * PCW;
      const float p0 = r[o], p1 = r[PCW + o], p2 = r[2 * PCW + o];
      const float nrm = sqrtf(p0 * p0 + p1 * p1 + p2 * p2) + VN_EPS;
      s1 += (double)nrm;
      s2 += (double)nrm * (double)nrm;
    }
  red[0][pl][o] = s1; red[1][pl][o] = s2;
  __syncthreads();
  if (pl == 0 && o < latent) {
    for (int q = 1; q < 8; ++q) { s1 += red[0][q][o]; s2 += red[1][q][o]; }
    part[(size_t)blockIdx.x * 2 * latent + o] = s1;
    part[(size_t)blockIdx.x * 2 * latent + latent + o] = s2;
  }
}

__global__ void __launch_bounds__(256) enc_c_apply_kernel(const float* __restrict__ pc, int latent, int P,
                                                          const float* __restrict__ param, float* __restrict__ lat) {
  __shared__ float red[3][8][32];
  const int o = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * P;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (o < latent) {
    const float scale = param[o], shift = param[latent + o];
    for (int p = pl; p < P; p += 8) {
      const float* r = pc + (base + p) * 3 * PCW;
      float p0 = r[o], p1 = r[PCW + o], p2 = r[2 * PCW + o];
      const float d0 = r[latent], d1 = r[PCW + latent], d2 = r[2 * PCW + latent];
      const float nrm = sqrtf(p0 * p0 + p1 * p1 + p2 * p2) + VN_EPS;
      const float f = (nrm * scale + shift) / nrm;
      p0 *= f; p1 *= f; p2 *= f;
      vn_leaky(p0, p1, p2, d0, d1, d2);
      a0 += p0; a1 += p1; a2 += p2;
    }
  }
  red[0][pl][o] = a0; red[1][pl][o] = a1; red[2][pl][o] = a2;
  __syncthreads();
  if (pl == 0 && o < latent) {
    for (int q = 1; q < 8; ++q) { a0 += red[0][q][o]; a1 += red[1][q][o]; a2 += red[2][q][o]; }
    const float inv = 1.f / (float)P;
    float* op = lat + ((size_t)blockIdx.x * latent + o) * 3;
    op[0] = a0 * inv; op[1] = a1 * inv; op[2] = a2 * inv;
  }
}

static int check(const smb_encoder_weights* w, int n_clouds, int n_points) {
  if (!w) { set_error_msg("encoder: null weights"); return SMB_E_BADARG; }
  if (w->hidden != HS) { set_error_msg("encoder: hidden_dim must be 128"); return SMB_E_UNSUPPORTED; }
  if (w->n_blocks < 1 || w->n_blocks > 8) { set_error_msg("encoder: layer_num must be 1..8"); return SMB_E_UNSUPPORTED; }
  if (w->latent < 1 || w->latent > 32) { set_error_msg("encoder: latent_dim must be <= 32"); return SMB_E_UNSUPPORTED; }
  if (n_clouds < 0 || n_points < 0) { set_error_msg("encoder: negative size"); return SMB_E_BADARG; }
  if (n_clouds > 0 && (w->num_k < 1 || w->num_k > n_points)) { set_error_msg("encoder: num_k must be in 1..n_points"); return SMB_E_BADARG; }
  if (n_points > 6144) { set_error_msg("encoder: more than 6144 points per cloud"); return SMB_E_TOOBIG; }
  return 0;
}

static int knn_rows(int P) {   // query rows per CTA so that the pd row block fits in shared memory
  const int Ppad = (P + 127) / 128 * 128;
  if (Ppad <= 1536) return 32;
  if (Ppad <= 3072) return 16;
  return 8;
}

template <int RM>
static int launch_knn_t(const float* feat, size_t row_stride, int segs, int seg_len, int seg_stride, const float* xx, int B, int P,
                        int k, int* idx, cudaStream_t st) {
  constexpr int R = 8 * RM;
  const int Ppad = (P + 127) / 128 * 128;
  const size_t smem = ((size_t)R * Ppad + 32 * (R + 1) + 32 * 132) * 4;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem(enc_knn_kernel<RM>, smem, configured)) return rc;
  const int grid = B * ((P + R - 1) / R);
  enc_knn_kernel<RM><<<grid, 256, smem, st>>>(feat, row_stride, segs, seg_len, seg_stride, xx, P, k, idx);
  return (int)cudaGetLastError();
}

static int launch_knn(const float* feat, size_t row_stride, int segs, int seg_len, int seg_stride, float* xx, int B, int P, int k,
                      int* idx, float* gram, cudaStream_t st) {
  const size_t n_points = (size_t)B * P;
  enc_sqnorm_kernel<<<(unsigned)((n_points + 7) / 8), 256, 0, st>>>(feat, row_stride, segs, seg_len, seg_stride, n_points, xx);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  if (gram && P <= GRAM_MAX_P && seg_len >= 32 && segs <= 4 && k <= 24) {
    // hidden features (R^384): Gram matrices on the tensor pipe, the three coordinate segments as concatenated operands
    TcGemmArgs g;
    memset(&g, 0, sizeof(g));
    for (int sg = 0; sg < segs; ++sg) {
      g.seg[sg].a = feat + (size_t)sg * seg_stride; g.seg[sg].lda = (long long)row_stride; g.seg[sg].k = seg_len;
      g.seg[sg].w_off = sg * seg_stride;
    }
    g.n_segs = segs; g.W = feat; g.ldw = (long long)row_stride; g.M = P; g.N = P; g.C = gram; g.ldc = P;
    g.a_batch = g.w_batch = (long long)P * (long long)row_stride; g.c_batch = (long long)P * P;
    g.split3 = 1;
    rc = launch_tc_gemm(g, B, st);
    if (rc) return rc;
    const unsigned grid = (unsigned)((n_points + 7) / 8);
    if (P <= 256) enc_topk_kernel<8><<<grid, 256, 0, st>>>(gram, xx, feat, row_stride, segs, seg_len, seg_stride, P, k, n_points, idx);
    else if (P <= 512) enc_topk_kernel<16><<<grid, 256, 0, st>>>(gram, xx, feat, row_stride, segs, seg_len, seg_stride, P, k, n_points, idx);
    else enc_topk_kernel<32><<<grid, 256, 0, st>>>(gram, xx, feat, row_stride, segs, seg_len, seg_stride, P, k, n_points, idx);
    return (int)cudaGetLastError();
  }
  switch (knn_rows(P)) {
    case 32: return launch_knn_t<4>(feat, row_stride, segs, seg_len, seg_stride, xx, B, P, k, idx, st);
    case 16: return launch_knn_t<2>(feat, row_stride, segs, seg_len, seg_stride, xx, B, P, k, idx, st);
    default: return launch_knn_t<1>(feat, row_stride, segs, seg_len, seg_stride, xx, B, P, k, idx, st);
  }
}

#define ENC_LAUNCH(expr)                                                                  \
  do {                                                                                    \
    int _rc = (expr);                                                                     \
    if (_rc != 0) { if (_rc > 0) set_error(#expr, (cudaError_t)_rc); return _rc; }        \
  } while (0)
#define ENC_KERNEL(...)                                                                   \
  do {                                                                                    \
    __VA_ARGS__;                                                                          \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) { set_error("encoder kernel launch", _e); return (int)_e; }    \
  } while (0)

static int encode(const smb_encoder_weights& w, const float* clouds, int B, int P, float* latent, void* ws_base, size_t ws_bytes,
                  cudaStream_t st) {
  const size_t n_points = (size_t)B * P, n_rows = n_points * 3;
  const int k = w.num_k, nb = w.n_blocks, HC = HS * nb;
  const Ws L = plan(nb, w.latent, k, n_points, P);
  if (!ws_base || ws_bytes < L.total) { set_error_msg("smb_vn_dgcnn_encode: workspace too small"); return SMB_E_BADARG; }
  unsigned char* base = reinterpret_cast<unsigned char*>(ws_base);
  float* xx = reinterpret_cast<float*>(base + L.xx);
  int* idx = reinterpret_cast<int*>(base + L.idx);
  float* h0 = reinterpret_cast<float*>(base + L.h0);
  float* hc = reinterpret_cast<float*>(base + L.hc);
  float* uv = reinterpret_cast<float*>(base + L.uv);
  float* pc = reinterpret_cast<float*>(base + L.pc);
  float* w4 = reinterpret_cast<float*>(base + L.w4);
  float* wc = reinterpret_cast<float*>(base + L.wc);
  double* part = reinterpret_cast<double*>(base + L.part);
  float* bnp = reinterpret_cast<float*>(base + L.bnp);
  float* gram = P <= GRAM_MAX_P ? reinterpret_cast<float*>(base + L.gram) : nullptr;
  void* wimg = base + L.wimg;
  const int stat_grid = (int)(n_points < (size_t)STAT_CTAS ? n_points : (size_t)STAT_CTAS);
  const double edge_count = (double)n_points * (double)k;

  // ---- conv_pos on the kNN graph of the coordinates (shape_pointcloud_modelAE.py:241-243) ----
  ENC_LAUNCH(launch_knn(clouds, 3, 1, 3, 0, xx, B, P, k, idx, nullptr, st));
  ENC_KERNEL(enc_prep_w4_kernel<<<1, HS, 0, st>>>(w.conv_pos_feat, w.conv_pos_dir, 1, w4));
  ENC_KERNEL(enc_pos_uv_kernel<<<(unsigned)((n_rows * (UVW / 4) + 255) / 256), 256, 0, st>>>(clouds, w4, n_rows, uv));
  ENC_KERNEL(enc_edge_stats_kernel<<<stat_grid, HS, 0, st>>>(uv, idx, P, k, n_points, part));
  ENC_KERNEL(enc_bn_final_kernel<<<1, HS, 0, st>>>(part, stat_grid, HS, edge_count, w.training, w.conv_pos_bn_w, w.conv_pos_bn_b,
                                                   w.conv_pos_bn_rm, w.conv_pos_bn_rv, bnp));
  ENC_KERNEL(enc_edge_apply_kernel<<<stat_grid, HS, 0, st>>>(uv, idx, P, k, n_points, bnp, h0, HS));

  // ---- DGCNN blocks on the dynamic graph of the hidden features (:247-250) ----
  for (int i = 0; i < nb; ++i) {
    const float* in = i == 0 ? h0 : hc + (size_t)(i - 1) * HS;
    const int ldi = i == 0 ? HS : HC;
    ENC_LAUNCH(launch_knn(in, (size_t)3 * ldi, 3, HS, ldi, xx, B, P, k, idx, gram, st));
    ENC_KERNEL(enc_prep_w4_kernel<<<(HS * HS + 255) / 256, 256, 0, st>>>(w.block_feat[i], w.block_dir[i], HS, w4));
    ENC_LAUNCH(enc_gemm(in, ldi, w4, HS, uv, UVW, n_rows, UVW, HS, wimg, L.wimg_bytes, st));
    ENC_KERNEL(enc_edge_stats_kernel<<<stat_grid, HS, 0, st>>>(uv, idx, P, k, n_points, part));
    ENC_KERNEL(enc_bn_final_kernel<<<1, HS, 0, st>>>(part, stat_grid, HS, edge_count, 1, w.block_bn_w[i], w.block_bn_b[i],
                                                     w.block_bn_rm[i], w.block_bn_rv[i], bnp));
    ENC_KERNEL(enc_edge_apply_kernel<<<stat_grid, HS, 0, st>>>(uv, idx, P, k, n_points, bnp, hc + (size_t)i * HS, HC));
  }

  // ---- conv_c + mean over the points (:252-254) ----
  SMB_CUDA_OK(cudaMemsetAsync(wc, 0, (size_t)PCW * HC * 4, st));
  SMB_CUDA_OK(cudaMemcpyAsync(wc, w.conv_c_feat, (size_t)w.latent * HC * 4, cudaMemcpyDeviceToDevice, st));
  SMB_CUDA_OK(cudaMemcpyAsync(wc + (size_t)w.latent * HC, w.conv_c_dir, (size_t)HC * 4, cudaMemcpyDeviceToDevice, st));
  ENC_LAUNCH(enc_gemm(hc, HC, wc, HC, pc, PCW, n_rows, w.latent + 1, HC, wimg, L.wimg_bytes, st));
  const int c_grid = (int)((n_points + 7) / 8 < (size_t)STAT_CTAS ? (n_points + 7) / 8 : (size_t)STAT_CTAS);
  ENC_KERNEL(enc_c_stats_kernel<<<c_grid, 256, 0, st>>>(pc, w.latent, n_points, part));
  ENC_KERNEL(enc_bn_final_kernel<<<1, HS, 0, st>>>(part, c_grid, w.latent, (double)n_points, w.training, w.conv_c_bn_w, w.conv_c_bn_b,
                                                   w.conv_c_bn_rm, w.conv_c_bn_rv, bnp));
  ENC_KERNEL(enc_c_apply_kernel<<<B, 256, 0, st>>>(pc, w.latent, P, bnp, latent));
  return 0;
}

}  // namespace enc
}  // namespace smb

extern "C" {

size_t smb_encoder_workspace_bytes(const smb_encoder_weights* w, int32_t n_clouds, int32_t n_points) {
  if (smb::enc::check(w, n_clouds, n_points)) return 0;
  const size_t n = (size_t)(n_clouds > 0 ? n_clouds : 1) * (size_t)(n_points > 0 ? n_points : 1);
  return smb::enc::plan(w->n_blocks, w->latent, w->num_k, n, n_points > 0 ? n_points : 1).total;
}

int smb_vn_dgcnn_encode(const smb_encoder_weights* w, const float* clouds, int32_t n_clouds, int32_t n_points, float* latent,
                        void* workspace, size_t workspace_bytes, void* stream) {
  int rc = smb::enc::check(w, n_clouds, n_points);
  if (rc) return rc;
  if (n_clouds == 0) return 0;
  if (n_points == 0) { smb::set_error_msg("smb_vn_dgcnn_encode: empty point cloud"); return SMB_E_BADARG; }
  if (!clouds || !latent) { smb::set_error_msg("smb_vn_dgcnn_encode: null pointer"); return SMB_E_BADARG; }
  for (int i = 0; i < w->n_blocks; ++i)
    if (!w->block_feat[i] || !w->block_dir[i] || !w->block_bn_w[i] || !w->block_bn_b[i]) {
      smb::set_error_msg("smb_vn_dgcnn_encode: null block weight pointer");
      return SMB_E_BADARG;
    }
  if (!w->conv_pos_feat || !w->conv_pos_dir || !w->conv_pos_bn_w || !w->conv_pos_bn_b || !w->conv_c_feat || !w->conv_c_dir ||
      !w->conv_c_bn_w || !w->conv_c_bn_b || !w->conv_pos_bn_rm || !w->conv_pos_bn_rv || !w->conv_c_bn_rm || !w->conv_c_bn_rv) {
    smb::set_error_msg("smb_vn_dgcnn_encode: null weight pointer");
    return SMB_E_BADARG;
  }
  return smb::enc::encode(*w, clouds, n_clouds, n_points, latent, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
