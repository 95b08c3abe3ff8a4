// Generic-shape path (hidden_dim != 128, e.g. the 256-wide / k = 48 / 60-atom stress configuration of BASELINE
// configs[4]): the reference formulation of one network evaluation (models/molopt_score_model.py:286-320,
// models/uni_transformer.py:48-162,289-333,475-540) evaluated literally in fp32 with small dedicated kernels -- a gathered
// tcgen05 GEMM (split-bf16 products, smb_tc_gemm.cu) for every Linear -- the edge-MLP input kv = [r | h_dst | h_src | inv] is one
// concatenated gathered operand --, row-wise LayerNorm / activation, and per-destination attention kernels over the dense
// [N, k+1] neighbour table.  No fusion across the MLP: this path exists so that every supported configuration has a CUDA
// implementation whose results match the reference at the fp32-parity level; the fused kernels (smb_edge_ws.cu,
// smb_node_tc5.cu, smb_edge_attn.cu) cover hidden 128.  Weights are read from the raw fp32 copy at the end of the packed blob
// (ModelLayout::raw, in smb_param_name order).
#include <map>
#include <mutex>

#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

namespace {

// ---- row-wise LayerNorm(eps 1e-5, affine) + ReLU (models/common.py:50-64), or shifted softplus (:39-45): one warp per row --
__global__ void __launch_bounds__(128) row_act_kernel(float* __restrict__ X, int M, int Hd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int act) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float* x = X + (size_t)row * Hd;
  if (act == ACT_SSP) {
    for (int c = lane; c < Hd; c += 32) {
      const float v = x[c];
      x[c] = (v > 20.f ? v : log1pf(expf(v))) - 0.69314718055994531f;
    }
    return;
  }
  float s = 0.f;
  for (int c = lane; c < Hd; c += 32) s += x[c];
  const float mean = warp_sum(s) / (float)Hd;
  float q = 0.f;
  for (int c = lane; c < Hd; c += 32) { const float d = x[c] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)Hd + 1e-5f);
  for (int c = lane; c < Hd; c += 32) x[c] = fmaxf(fmaf((x[c] - mean) * rstd, gamma[c], beta[c]), 0.f);
}

// ---- edge geometry of the dense slot table: e = i * KS + slot ------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_geom_kernel(const float* __restrict__ x, const int* __restrict__ nbr, const int* __restrict__ deg,
                                                        const int* __restrict__ atom_mol, const int* __restrict__ mol_ptr, int n_atoms, int KS,
                                                        int* __restrict__ dst_idx, int* __restrict__ src_idx, int* __restrict__ mol_idx,
                                                        float* __restrict__ rbf, float* __restrict__ rel) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_atoms * KS) return;
  const int i = e / KS, s = e - i * KS;
  float r[kRbf];
#pragma unroll
  for (int g = 0; g < kRbf; ++g) r[g] = 0.f;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  int di = -1, si = -1, mi = -1;
  if (s < deg[i]) {
    const int m = atom_mol[i];
    const int j = mol_ptr[m] + nbr[(size_t)i * KS + s];
    di = i; si = j; mi = m;
    rx = x[i * 3] - x[j * 3]; ry = x[i * 3 + 1] - x[j * 3 + 1]; rz = x[i * 3 + 2] - x[j * 3 + 2];
    const float d = sqrtf(rx * rx + ry * ry + rz * rz);
#pragma unroll
    for (int g = 0; g < kRbf; ++g) { const float dd = d - rbf_centre(g); r[g] = expf(-0.5f * dd * dd); }
  }
  dst_idx[e] = di; src_idx[e] = si; mol_idx[e] = mi;
#pragma unroll
  for (int g = 0; g < kRbf; ++g) rbf[(size_t)e * kRbf + g] = r[g];
  rel[(size_t)e * 3] = rx; rel[(size_t)e * 3 + 1] = ry; rel[(size_t)e * 3 + 2] = rz;
}

// ---- e_w = sigmoid(w3 . hid + b3)  (uni_transformer.py:475-481): one warp per slot ----------------------------------------
__global__ void __launch_bounds__(128) gate_dot_kernel(const float* __restrict__ hid, int M, int Hd, const float* __restrict__ w3,
                                                       const float* __restrict__ b3, float* __restrict__ ew) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= M) return;
  float s = 0.f;
  for (int c = lane; c < Hd; c += 32) s = fmaf(hid[(size_t)e * Hd + c], w3[c], s);
  s = warp_sum(s);
  if (lane == 0) ew[e] = 1.f / (1.f + expf(-(s + b3[0])));
}

// ---- alpha = softmax over a destination's slots of <q_i, k_e> / sqrt(dh) per head (uni_transformer.py:77,147) -----------
// one warp per destination atom; lane handles slots lane and lane + 32
__global__ void __launch_bounds__(128) attn_alpha_kernel(const float* __restrict__ q, const float* __restrict__ kbuf, const int* __restrict__ deg,
                                                         int n_atoms, int KS, int Hd, int heads, float* __restrict__ alpha) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n_atoms) return;
  const int dg = deg[i], dh = Hd / heads;
  const float scale = 1.f / sqrtf((float)dh);
  for (int h = 0; h < heads; ++h) {
    float l[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int s = lane + 32 * u;
      l[u] = -INFINITY;
      if (s < dg) {
        const float* kr = kbuf + ((size_t)i * KS + s) * Hd + h * dh;
        const float* qr = q + (size_t)i * Hd + h * dh;
        float acc = 0.f;
        for (int c = 0; c < dh; ++c) acc = fmaf(qr[c], kr[c], acc);
        l[u] = acc * scale;
      }
    }
    float mx = fmaxf(l[0], l[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e0 = lane < dg ? expf(l[0] - mx) : 0.f, e1 = lane + 32 < dg ? expf(l[1] - mx) : 0.f;
    const float inv = 1.f / warp_sum(e0 + e1);
    if (lane < dg) alpha[((size_t)i * KS + lane) * heads + h] = e0 * inv;
    if (lane + 32 < dg) alpha[((size_t)i * KS + lane + 32) * heads + h] = e1 * inv;
  }
}

// Same, for hidden widths that are multiples of 128 with 4 | dh | 128: a warp reads every key row with coalesced 16-byte loads (lane =
// four consecutive channels of each 128-channel piece), the dh / 4 lanes of a head reduce with shuffles, logits go through shared
// memory [slot][head], and lane h does the softmax of head h.  (The kernel above reads a row 4 bytes at a time per (head, slot).)
template <int PIECES>
__global__ void __launch_bounds__(128) attn_alpha_rows_kernel(const float* __restrict__ q, const float* __restrict__ kbuf, const int* __restrict__ deg,
                                                              int n_atoms, int KS, int heads, float* __restrict__ alpha) {
  extern __shared__ float s_logit[];                      // [4 warps][KS][heads]
  constexpr int Hd = PIECES * 128;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + w;
  if (i >= n_atoms) return;
  const int dg = deg[i], dh = Hd / heads, lph = dh / 4;   // lanes per head
  const float scale = 1.f / sqrtf((float)dh);
  float* lg = s_logit + (size_t)w * KS * heads;
  float4 q4[PIECES];
#pragma unroll
  for (int p = 0; p < PIECES; ++p) q4[p] = __ldg(reinterpret_cast<const float4*>(q + (size_t)i * Hd + p * 128) + lane);
  const float4* krow = reinterpret_cast<const float4*>(kbuf + (size_t)i * KS * Hd) + lane;
  for (int s = 0; s < dg; ++s) {
    float part[PIECES];
#pragma unroll
    for (int p = 0; p < PIECES; ++p) {
      const float4 k4 = __ldg(krow + (size_t)s * (Hd / 4) + p * 32);
      part[p] = fmaf(q4[p].x, k4.x, fmaf(q4[p].y, k4.y, fmaf(q4[p].z, k4.z, q4[p].w * k4.w)));
    }
#pragma unroll
    for (int p = 0; p < PIECES; ++p) {
      float v = part[p];
      for (int o = 1; o < lph; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane % lph == 0) lg[s * heads + p * (128 / dh) + lane / lph] = v * scale;
    }
  }
  __syncwarp();
  for (int h = lane; h < heads; h += 32) {
    float mx = -INFINITY;
    for (int s = 0; s < dg; ++s) mx = fmaxf(mx, lg[s * heads + h]);
    float sum = 0.f;
    for (int s = 0; s < dg; ++s) sum += expf(lg[s * heads + h] - mx);
    const float inv = 1.f / sum;
    for (int s = 0; s < dg; ++s) alpha[((size_t)i * KS + s) * heads + h] = expf(lg[s * heads + h] - mx) * inv;
  }
}

int launch_attn_alpha(const float* q, const float* kbuf, const int* deg, int N, int KS, int Hd, int heads, float* alpha, cudaStream_t st) {
  if (N <= 0) return 0;
  const int dh = heads > 0 ? Hd / heads : 0;
  const size_t smem = (size_t)4 * KS * heads * sizeof(float);
  const bool rows = Hd % 128 == 0 && Hd <= 512 && dh >= 4 && dh <= 128 && (dh & (dh - 1)) == 0 && smem <= 48 * 1024 &&
                    (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(kbuf) & 15) == 0;
  const unsigned grid = (unsigned)((N + 3) / 4);
  if (rows && Hd == 128) attn_alpha_rows_kernel<1><<<grid, 128, smem, st>>>(q, kbuf, deg, N, KS, heads, alpha);
  else if (rows && Hd == 256) attn_alpha_rows_kernel<2><<<grid, 128, smem, st>>>(q, kbuf, deg, N, KS, heads, alpha);
  else if (rows && Hd == 384) attn_alpha_rows_kernel<3><<<grid, 128, smem, st>>>(q, kbuf, deg, N, KS, heads, alpha);
  else if (rows && Hd == 512) attn_alpha_rows_kernel<4><<<grid, 128, smem, st>>>(q, kbuf, deg, N, KS, heads, alpha);
  else attn_alpha_kernel<<<grid, 128, 0, st>>>(q, kbuf, deg, N, KS, Hd, heads, alpha);
  return (int)cudaGetLastError();
}

// ---- X2H aggregation: agg[i][c] = sum_s alpha[e][head(c)] e_w[e] v[e][c]  (uni_transformer.py:69-81) ---------------------
__global__ void __launch_bounds__(256) agg_v_kernel(const float* __restrict__ alpha, const float* __restrict__ ew, const float* __restrict__ vbuf,
                                                    const int* __restrict__ deg, int n_atoms, int KS, int Hd, int heads, float* __restrict__ agg) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_atoms * Hd) return;
  const int i = (int)(idx / Hd), c = (int)(idx % Hd), h = c / (Hd / heads);
  float acc = 0.f;
  for (int s = 0; s < deg[i]; ++s) {
    const size_t e = (size_t)i * KS + s;
    acc = fmaf(alpha[e * heads + h] * ew[e], vbuf[e * Hd + c], acc);
  }
  agg[idx] = acc;
}

// ---- H2X: o[i][h] = sum_s alpha e_w w (x_i - x_j); VN linear maps; BatchNorm partial sums (uni_transformer.py:139-156,
//      shape_vn_layers.py:95-110).  One warp per atom: lane = (feat | dir, channel). ------------------------------------------
__global__ void __launch_bounds__(128) xv_vn_kernel(const float* __restrict__ alpha, const float* __restrict__ ew, const float* __restrict__ wbuf,
                                                    const float* __restrict__ rel, const int* __restrict__ deg, const int* __restrict__ atom_mol,
                                                    const float* __restrict__ x, const float* __restrict__ shape, const float* __restrict__ vn_feat,
                                                    const float* __restrict__ vn_dir, int n_atoms, int KS, float* __restrict__ vn,
                                                    float* __restrict__ bn_partial) {
  __shared__ float s_o[4][kHeads][3];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + w;
  if (i >= n_atoms) return;
  if (lane < kHeads) {
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int s = 0; s < deg[i]; ++s) {
      const size_t e = (size_t)i * KS + s;
      const float c = alpha[e * kHeads + lane] * ew[e] * wbuf[e * kHeads + lane];
      ox = fmaf(c, rel[e * 3], ox); oy = fmaf(c, rel[e * 3 + 1], oy); oz = fmaf(c, rel[e * 3 + 2], oz);
    }
    s_o[w][lane][0] = ox; s_o[w][lane][1] = oy; s_o[w][lane][2] = oz;
  }
  __syncwarp();
  const int which = lane >> 4, ch = lane & 15;
  const float* wt = (which ? vn_dir : vn_feat) + ch * kVnStride;
  const float* sh = shape + (size_t)atom_mol[i] * kShape * 3;
  float vx = wt[0] * x[i * 3], vy = wt[0] * x[i * 3 + 1], vz = wt[0] * x[i * 3 + 2];
  for (int cc = 0; cc < kHeads; ++cc) {
    vx = fmaf(wt[1 + cc], s_o[w][cc][0], vx); vy = fmaf(wt[1 + cc], s_o[w][cc][1], vy); vz = fmaf(wt[1 + cc], s_o[w][cc][2], vz);
  }
  for (int cc = 0; cc < kShape; ++cc) {
    vx = fmaf(wt[1 + kHeads + cc], sh[cc * 3], vx); vy = fmaf(wt[1 + kHeads + cc], sh[cc * 3 + 1], vy); vz = fmaf(wt[1 + kHeads + cc], sh[cc * 3 + 2], vz);
  }
  float* row = vn + (size_t)i * kVnRow;
  row[3 + which * 48 + ch * 3] = vx; row[4 + which * 48 + ch * 3] = vy; row[5 + which * 48 + ch * 3] = vz;
  if (lane < 3) {
    float sm = 0.f;
    for (int cc = 0; cc < kHeads; ++cc) sm += s_o[w][cc][lane];
    row[lane] = sm * (1.f / kHeads);
  }
  if (which == 0) {   // one partial row per atom: [0..15] = nu, [16..31] = nu^2
    const float nu = sqrtf(vx * vx + vy * vy + vz * vz) + 1e-6f;
    bn_partial[(size_t)i * 32 + ch] = nu;
    bn_partial[(size_t)i * 32 + 16 + ch] = nu * nu;
  }
}

// C[M x N] (+)= A[rows gathered by a_idx][K] . W[N][K]^T (+ bias): the tcgen05 split-bf16 GEMM (smb_tc_gemm.cu)
struct Seg { const float* a; const int* idx; int lda; int k; };
// scratch for the pre-split weight image of the GEMM in flight: part of the calling forward's workspace (stream-ordered reuse)
thread_local void* t_wimg = nullptr;
thread_local long long t_wimg_bytes = 0;
int row_act(float* X, int M, int Hd, const float* g, const float* b, int act, cudaStream_t st);
// ln_g / ln_b: LayerNorm -> ReLU of the output rows, fused into the GEMM's epilogue when the whole row is in one column tile
int gemm_cat(const Seg* segs, int n_segs, const float* W, int ldw, int M, int N, const float* bias, bool accumulate, float* C, int ldc,
             cudaStream_t st, const float* ln_g = nullptr, const float* ln_b = nullptr) {
  TcGemmArgs g;
  memset(&g, 0, sizeof(g));
  int off = 0;
  for (int s = 0; s < n_segs; ++s) {
    g.seg[s].a = segs[s].a; g.seg[s].idx = segs[s].idx; g.seg[s].lda = segs[s].lda; g.seg[s].k = segs[s].k; g.seg[s].w_off = off;
    off += segs[s].k;
  }
  g.n_segs = n_segs; g.W = W; g.ldw = ldw; g.M = M; g.N = N; g.bias = bias; g.accumulate = accumulate ? 1 : 0; g.C = C; g.ldc = ldc;
  g.w_img = t_wimg; g.w_img_bytes = t_wimg_bytes;
  const bool fuse = ln_g && ln_b && ldc == N && tc_gemm_can_fuse_ln(N, accumulate);
  if (fuse) { g.ln_gamma = ln_g; g.ln_beta = ln_b; }
  if (int rc = launch_tc_gemm(g, 1, st)) return rc;
  return (ln_g && ln_b && !fuse) ? row_act(C, M, N, ln_g, ln_b, ACT_LN_RELU, st) : 0;
}
int gemm(const float* A, const int* a_idx, int lda, const float* W, int ldw, int M, int N, int K, const float* bias, bool accumulate,
         float* C, int ldc, cudaStream_t st, const float* ln_g = nullptr, const float* ln_b = nullptr) {
  const Seg sg = {A, a_idx, lda, K};
  return gemm_cat(&sg, 1, W, ldw, M, N, bias, accumulate, C, ldc, st, ln_g, ln_b);
}
int row_act(float* X, int M, int Hd, const float* g, const float* b, int act, cudaStream_t st) {
  if (M <= 0) return 0;
  row_act_kernel<<<(M + 3) / 4, 128, 0, st>>>(X, M, Hd, g, b, act);
  return (int)cudaGetLastError();
}

}  // namespace

// ---- raw parameter table ---------------------------------------------------------------------------------------------
size_t param_numel(const smb_model_dims& d, const std::string& n) {
  const size_t H = d.hidden, C = d.classes, T = d.time_dim, KV = 2 * H + kRbf + kShape, heads = d.heads;
  auto has = [&](const char* s) { return n.find(s) != std::string::npos; };
  auto ends = [&](const char* s) { const std::string t(s); return n.size() >= t.size() && n.compare(n.size() - t.size(), t.size(), t) == 0; };
  if (n == "time_emb.1.weight") return 2 * T * T;
  if (n == "time_emb.1.bias") return 2 * T;
  if (n == "time_emb.3.weight") return T * 2 * T;
  if (n == "time_emb.3.bias") return T;
  if (n == "ligand_atom_emb.weight") return H * (C + T);
  if (n == "ligand_atom_emb.bias") return H;
  if (n == "v_inference.0.weight") return H * H;
  if (n == "v_inference.0.bias") return H;
  if (n == "v_inference.2.weight") return C * H;
  if (n == "v_inference.2.bias") return C;
  if (has("shape_linear")) return (size_t)kHeads * kVnIn;
  if (has("invariant_shape_layer")) return ends("0.weight") || ends("3.weight") ? (size_t)kShape * kShape : (size_t)kShape;
  if (has("edge_pred_layer")) return ends("0.weight") ? H * kRbf : ends("3.weight") ? H : ends("3.bias") ? 1 : H;
  const bool edge = has("hk_func") || has("hv_func") || has("xk_func") || has("xv_func");
  if (ends("0.weight")) return edge ? H * KV : has("node_output") ? H * 2 * H : H * H;
  if (ends("3.weight")) return has("xv_func") ? heads * H : H * H;
  if (ends("3.bias")) return has("xv_func") ? heads : H;
  return H;   // first-Linear bias, LayerNorm weight / bias
}

namespace {
struct RawTable { std::map<std::string, size_t> off; size_t total = 0; };
const RawTable& raw_table(const smb_model_dims& d) {
  static std::mutex mu;
  static std::map<std::vector<int>, RawTable> cache;
  std::lock_guard<std::mutex> lock(mu);
  const std::vector<int> key = {d.hidden, d.heads, d.layers, d.classes, d.time_dim};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  RawTable t;
  for (const std::string& n : param_names(d)) { t.off[n] = t.total; t.total += align_up(param_numel(d, n) * 4, 16); }
  return cache.emplace(key, std::move(t)).first->second;
}
}  // namespace

size_t raw_weights_bytes(const smb_model_dims& d) { return raw_table(d).total; }
size_t raw_weight_offset(const smb_model_dims& d, const std::string& name) { return raw_table(d).off.at(name); }

#define SMB_G(expr)                                                                          \
  do {                                                                                       \
    int _rc = (expr);                                                                        \
    if (_rc != 0) { if (_rc > 0) set_error(#expr, (cudaError_t)_rc); return _rc; }           \
  } while (0)

// One network evaluation; same contract as forward_impl (smb_api.cu).  x, h0 (in h_a), tau / inv and the kNN table are already
// in the workspace (prep / embed / knn kernels are shape independent).
int forward_generic(const smb_model_dims& d, const void* blob, const ModelLayout& L, const Workspace& W, void* ws_base, const smb_batch& b,
                    const smb_forward_io& io, cudaStream_t st) {
  const int N = b.n_atoms, H = d.hidden, KS = d.k + 1, heads = d.heads, M = N * KS, KV = 2 * H + kRbf + kShape;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(blob) + L.raw;
  auto P = [&](const std::string& n) { return reinterpret_cast<const float*>(base + raw_weight_offset(d, n)); };
  auto wsf = [&](size_t off) { return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws_base) + off); };
  auto wsi = [&](size_t off) { return reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws_base) + off); };
  t_wimg = reinterpret_cast<unsigned char*>(ws_base) + W.g_wimg;
  t_wimg_bytes = (long long)W.g_wimg_bytes;
  float* x = wsf(W.x);
  float* inv = wsf(W.inv);
  const int* nbr = wsi(W.nbr);
  const int* deg = wsi(W.deg);
  float* ew = wsf(W.ew);
  float* vn = wsf(W.vn);
  float* bn_param = wsf(W.bn_param);
  float* hbuf[2] = {wsf(W.h_a), wsf(W.h_b)};
  float* q = wsf(W.q);
  float* agg = wsf(W.agg);
  float* ghid = wsf(W.g_hid);       // [M][H]
  float* gout = wsf(W.g_out);       // [M][H]
  float* galpha = wsf(W.g_alpha);   // [M][heads]
  float* grbf = wsf(W.g_rbf);       // [M][20]
  float* grel = wsf(W.g_rel);       // [M][3]
  float* gnode = wsf(W.g_node);     // [N][H] node-level hidden
  float* gbn = wsf(W.g_bn);         // [N][32]
  int* dst_idx = wsi(W.g_idx);
  int* src_idx = dst_idx + M;
  int* mol_idx = src_idx + M;

  auto geom = [&]() -> int {
    edge_geom_kernel<<<(M + 255) / 256, 256, 0, st>>>(x, nbr, deg, b.atom_mol, b.mol_ptr, N, KS, dst_idx, src_idx, mol_idx, grbf, grel);
    return (int)cudaGetLastError();
  };
  // Linear -> LayerNorm -> ReLU -> Linear on the edge rows  kv = [r | h_dst | h_src | inv_dst]  (MLP of models/common.py:47-67)
  auto edge_mlp = [&](const std::string& p, const float* h, int n2, float* out) -> int {
    const float* w0 = P(p + ".net.0.weight");
    // kv = [r | h_dst | h_src | inv_dst] (uni_transformer.py:61-63) as ONE gathered GEMM over the concatenated operand
    const Seg kv[4] = {{grbf, nullptr, kRbf, kRbf}, {h, dst_idx, H, H}, {h, src_idx, H, H}, {inv, mol_idx, kShape, kShape}};
    SMB_G(gemm_cat(kv, 4, w0, KV, M, H, P(p + ".net.0.bias"), false, ghid, H, st, P(p + ".net.1.weight"), P(p + ".net.1.bias")));
    SMB_G(gemm(ghid, nullptr, H, P(p + ".net.3.weight"), H, M, n2, H, P(p + ".net.3.bias"), false, out, n2, st));
    return 0;
  };
  auto node_mlp = [&](const std::string& p, const float* h, float* out) -> int {   // hq_func / xq_func
    SMB_G(gemm(h, nullptr, H, P(p + ".net.0.weight"), H, N, H, H, P(p + ".net.0.bias"), false, gnode, H, st, P(p + ".net.1.weight"),
               P(p + ".net.1.bias")));
    SMB_G(gemm(gnode, nullptr, H, P(p + ".net.3.weight"), H, N, H, H, P(p + ".net.3.bias"), false, out, H, st));
    return 0;
  };

  // ---- global edge gate from the input coordinates (uni_transformer.py:507) ----
  SMB_G(geom());
  {
    const std::string p = "refine_net.edge_pred_layer";
    SMB_G(gemm(grbf, nullptr, kRbf, P(p + ".net.0.weight"), kRbf, M, H, kRbf, P(p + ".net.0.bias"), false, ghid, H, st, P(p + ".net.1.weight"),
               P(p + ".net.1.bias")));
    gate_dot_kernel<<<(M + 3) / 4, 128, 0, st>>>(ghid, M, H, P(p + ".net.3.weight"), P(p + ".net.3.bias"), ew);
    SMB_G((int)cudaGetLastError());
  }
  int cur = 0;
  for (int l = 0; l < d.layers; ++l) {
    const bool last = l == d.layers - 1;
    const std::string blk = "refine_net.base_block." + std::to_string(l);
    const std::string x2h = blk + ".x2h_layers.0", h2x = blk + ".h2x_layers.0";
    float* h_in = hbuf[cur];
    float* h_out = last ? io.pred_h : hbuf[cur ^ 1];
    if (l > 0) SMB_G(geom());   // rel_x / r_feat from the current coordinates (uni_transformer.py:300-311)
    // ---- X2H (uni_transformer.py:48-90) ----
    SMB_G(edge_mlp(x2h + ".hk_func", h_in, H, gout));
    SMB_G(node_mlp(x2h + ".hq_func", h_in, q));
    SMB_G(launch_attn_alpha(q, gout, deg, N, KS, H, heads, galpha, st));
    SMB_G(edge_mlp(x2h + ".hv_func", h_in, H, gout));
    agg_v_kernel<<<(unsigned)(((size_t)N * H + 255) / 256), 256, 0, st>>>(galpha, ew, gout, deg, N, KS, H, heads, agg);
    SMB_G((int)cudaGetLastError());
    {
      const std::string p = x2h + ".node_output";
      const float* w0 = P(p + ".net.0.weight");
      const Seg cat[2] = {{agg, nullptr, H, H}, {h_in, nullptr, H, H}};     // [agg | h] (uni_transformer.py:82)
      SMB_G(gemm_cat(cat, 2, w0, 2 * H, N, H, P(p + ".net.0.bias"), false, gnode, H, st, P(p + ".net.1.weight"), P(p + ".net.1.bias")));
      SMB_CUDA_OK(cudaMemcpyAsync(h_out, h_in, (size_t)N * H * sizeof(float), cudaMemcpyDeviceToDevice, st));   // residual (:87-88)
      SMB_G(gemm(gnode, nullptr, H, P(p + ".net.3.weight"), H, N, H, H, P(p + ".net.3.bias"), true, h_out, H, st));
    }
    // ---- H2X on the updated h (uni_transformer.py:121-162) ----
    SMB_G(edge_mlp(h2x + ".xk_func", h_out, H, gout));
    SMB_G(node_mlp(h2x + ".xq_func", h_out, q));
    SMB_G(launch_attn_alpha(q, gout, deg, N, KS, H, heads, galpha, st));
    SMB_G(edge_mlp(h2x + ".xv_func", h_out, heads, gout));
    xv_vn_kernel<<<(N + 3) / 4, 128, 0, st>>>(galpha, ew, gout, grel, deg, b.atom_mol, x, io.shape,
                                              P(h2x + ".shape_linear.map_to_feat.weight"), P(h2x + ".shape_linear.map_to_dir.weight"), N, KS, vn, gbn);
    SMB_G((int)cudaGetLastError());
    BnArgs bn;
    bn.training = io.training; bn.n_atoms = N; bn.rows = N; bn.partial = gbn;
    bn.weight = io.bn_weight[l]; bn.bias = io.bn_bias[l];
    bn.running_mean = io.bn_running_mean[l]; bn.running_var = io.bn_running_var[l];
    bn.num_batches_tracked = io.bn_num_batches_tracked[l];
    bn.param = bn_param;
    SMB_G(launch_bn_final(bn, st));
    SMB_G(launch_vn_apply(vn, bn_param, x, last ? io.pred_pos : nullptr, N, st));
    cur ^= 1;
  }
  // type head v_inference: Linear -> shifted softplus -> Linear (molopt_score_model.py:262-266,305) as two tensor-core GEMMs
  // (type_head_generic below -- one CTA per atom, no scratch -- serves the scratch-free smb_type_head entry point)
  SMB_G(gemm(io.pred_h, nullptr, H, P("v_inference.0.weight"), H, N, H, H, P("v_inference.0.bias"), false, gnode, H, st));
  SMB_G(row_act(gnode, N, H, nullptr, nullptr, ACT_SSP, st));
  SMB_G(gemm(gnode, nullptr, H, P("v_inference.2.weight"), H, N, d.classes, H, P("v_inference.2.bias"), false, io.pred_v, d.classes, st));
  return 0;
}

// v_inference: Linear -> shifted softplus -> Linear (molopt_score_model.py:262-266,305); one CTA per atom, no scratch
namespace {
__global__ void __launch_bounds__(256) head_generic_kernel(const float* __restrict__ h, int Hd, int classes, const float* __restrict__ w0,
                                                           const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ b2,
                                                           float* __restrict__ logits) {
  extern __shared__ float s_y[];   // [Hd] input row, then [Hd] hidden
  float* s_h = s_y;
  float* s_z = s_y + Hd;
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < Hd; c += blockDim.x) s_h[c] = h[(size_t)i * Hd + c];
  __syncthreads();
  for (int c = threadIdx.x; c < Hd; c += blockDim.x) {
    float acc = b0[c];
    for (int k = 0; k < Hd; ++k) acc = fmaf(w0[(size_t)c * Hd + k], s_h[k], acc);
    s_z[c] = (acc > 20.f ? acc : log1pf(expf(acc))) - 0.69314718055994531f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < classes; c += blockDim.x) {
    float acc = b2[c];
    for (int k = 0; k < Hd; ++k) acc = fmaf(w2[(size_t)c * Hd + k], s_z[k], acc);
    logits[(size_t)i * classes + c] = acc;
  }
}
}  // namespace

int type_head_generic(const smb_model_dims& d, const void* blob, const ModelLayout& L, int N, const float* h, float* logits, cudaStream_t st) {
  if (N <= 0) return 0;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(blob) + L.raw;
  auto P = [&](const char* n) { return reinterpret_cast<const float*>(base + raw_weight_offset(d, n)); };
  head_generic_kernel<<<N, 256, 2 * d.hidden * sizeof(float), st>>>(h, d.hidden, d.classes, P("v_inference.0.weight"), P("v_inference.0.bias"),
                                                                     P("v_inference.2.weight"), P("v_inference.2.bias"), logits);
  int rc = (int)cudaGetLastError();
  if (rc > 0) set_error("head_generic_kernel launch", (cudaError_t)rc);
  return rc;
}

}  // namespace smb
