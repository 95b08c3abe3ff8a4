// Host side of the C ABI: error reporting, parameter enumeration, weight packing, workspace layout.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

#include "smb_common.cuh"
#include "smb_layout.h"

namespace smb { long long tc_gemm_w_img_bytes(int N, const int* seg_k, int n_segs, bool split3); }   // smb_tc_gemm.cu

namespace smb {

static thread_local char g_err[512] = "";

void set_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s failed: %s (%d)", what, cudaGetErrorString(e), (int)e);
}
void set_error_msg(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

int check_dims(const smb_model_dims& d) {
  // hidden 128: tuned tensor-core kernels; any other width (e.g. 256): generic fp32 path (smb_generic.cu)
  if (d.hidden < 32 || d.hidden > 512 || d.hidden % 32 != 0) { set_error_msg("hidden_dim must be a multiple of 32 in 32..512"); return SMB_E_UNSUPPORTED; }
  if (d.heads != kHeads) { set_error_msg("n_heads must be 16"); return SMB_E_UNSUPPORTED; }
  if (d.layers < 1 || d.layers > kMaxLayers) { set_error_msg("num_layers out of range"); return SMB_E_UNSUPPORTED; }
  if (d.k < 1 || d.k > SMB_MAX_K) { set_error_msg("knn out of range (1..63)"); return SMB_E_TOOBIG; }
  if (d.classes < 2 || d.classes > 16) { set_error_msg("num classes must be <= 16"); return SMB_E_UNSUPPORTED; }
  if (d.time_dim != 8) { set_error_msg("time_emb_dim must be 8"); return SMB_E_UNSUPPORTED; }
  if (d.precision != SMB_PREC_BF16X3 && d.precision != SMB_PREC_BF16) { set_error_msg("bad precision"); return SMB_E_BADARG; }
  if (d.timesteps < 1) { set_error_msg("bad timesteps"); return SMB_E_BADARG; }
  return 0;
}

// ---- layout ---------------------------------------------------------------------------------
namespace {
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; }
};

EdgeMlpOff carve_edge(Carver& c, const smb_model_dims& d, int n2, bool gate) {
  const int H = d.hidden;
  const size_t fb = frag_bytes(d);
  EdgeMlpOff e;
  e.w1r = c.take((size_t)(H / 8) * 2 * 32 * fb);
  e.b1 = c.take(H * 4);
  e.ln_g = c.take(H * 4);
  e.ln_b = c.take(H * 4);
  e.w2 = gate ? c.take(H * 4) : c.take((size_t)(n2 / 8) * (H / 16) * 32 * fb);
  e.b2 = c.take((gate ? 4 : n2) * 4);
  e.w1r_u = c.take((size_t)32 * H * 2);
  e.w2_u = gate ? e.w1r_u : c.take((size_t)n2 * H * 2);
  e.w1r_f = c.take((size_t)32 * H * 2);
  e.w2_f = gate ? c.take(H * 4) : c.take((size_t)n2 * H * 2);   // gate: fp32 vector w2 * |gamma|
  e.beta_f = c.take(H * 4);
  e.w2_q = (!gate && n2 == H) ? c.take((size_t)H * H * 2) : e.w2_f;
  return e;
}
NodeMlpOff carve_node(Carver& c, const smb_model_dims& d, int n1, int k1, int n2, bool folded = false, bool out_tc5 = false) {
  const int H = d.hidden;
  const size_t fb = frag_bytes(d);
  NodeMlpOff n;
  n.w1 = c.take((size_t)(n1 / 8) * (k1 / 16) * 32 * fb);
  n.b1 = c.take(n1 * 4);
  n.ln_g = c.take(H * 4);
  n.ln_b = c.take(H * 4);
  n.w2 = c.take((size_t)(n2 / 8) * (H / 16) * 32 * fb);
  n.b2 = c.take(n2 * 4);
  n.w1_f = folded ? c.take((size_t)(n1 / 8) * (k1 / 16) * 32 * fb) : n.w1;
  n.b1_f = folded ? c.take(n1 * 4) : n.b1;
  n.w1_t = folded ? c.take((size_t)5 * kNodeChunkBytes) : n.w1;
  n.w2_t = folded ? c.take((size_t)H * H * 2) : n.w2;
  n.beta_t = folded ? c.take(H * 4) : n.ln_b;
  if (out_tc5) {
    n.w1_t = c.take((size_t)128 * kNodeOutKx * 2);
    n.w2_t = c.take((size_t)H * H * 2);
    n.beta_t = c.take(H * 4);
  }
  return n;
}
}  // namespace

ModelLayout build_layout(const smb_model_dims& d) {
  const int H = d.hidden;
  Carver c;
  ModelLayout L;
  memset(&L, 0, sizeof(L));
  L.time_freq = c.take(d.time_dim / 2 * 4);
  L.time_w1 = c.take(d.time_dim * 2 * d.time_dim * 4);
  L.time_b1 = c.take(d.time_dim * 2 * 4);
  L.time_w2 = c.take(d.time_dim * d.time_dim * 2 * 4);
  L.time_b2 = c.take(d.time_dim * 4);
  L.emb_wT = c.take((size_t)(d.classes + d.time_dim) * H * 4);
  L.emb_b = c.take(H * 4);
  L.inv_w1 = c.take(kShape * kShape * 4);
  L.inv_b1 = c.take(kShape * 4);
  L.inv_g = c.take(kShape * 4);
  L.inv_bb = c.take(kShape * 4);
  L.inv_w2 = c.take(kShape * kShape * 4);
  L.inv_b2 = c.take(kShape * 4);
  L.gate = carve_edge(c, d, 0, true);
  L.head = carve_node(c, d, H, H, 16);
  for (int l = 0; l < d.layers; ++l) {
    LayerOff& y = L.layer[l];
    y.hk = carve_edge(c, d, H, false);
    y.hv = carve_edge(c, d, H, false);
    y.xk = carve_edge(c, d, H, false);
    y.xv = carve_edge(c, d, kHeads, false);
    y.x2h_pre = carve_node(c, d, 5 * H, H + kShape, H, true);
    y.node_out = carve_node(c, d, H, 2 * H, H, false, true);
    y.h2x_pre = carve_node(c, d, 5 * H, H + kShape, H, true);
    y.vn_feat = c.take(kHeads * kVnStride * 4);
    y.vn_dir = c.take(kHeads * kVnStride * 4);
  }
  L.raw = c.off;
  if (d.hidden != 128) c.take(raw_weights_bytes(d));
  L.total = c.off;
  return L;
}

Workspace build_workspace(const smb_model_dims& d, int N, int B) {
  const int H = d.hidden, KS = d.k + 1;
  Carver c;
  Workspace w;
  memset(&w, 0, sizeof(w));
  const size_t n = (size_t)(N > 0 ? N : 1), b = (size_t)(B > 0 ? B : 1);
  w.tau = c.take(b * 8 * 4);
  w.inv = c.take(b * kShape * 4);
  w.nbr = c.take(n * KS * 4);
  w.deg = c.take(n * 4);
  w.ew = c.take(n * KS * 4);
  w.max_tiles = (int)(n / 4 + b + 8);
  {  // alpha: [N][k+1][16] (atom-strided kernels) or [tile][kAlphaTileFloats] (warp-specialised pipeline)
    const size_t by_atom = n * KS * kHeads * 4, by_tile = (size_t)w.max_tiles * kAlphaTileFloats * 4;
    w.alpha = c.take(by_atom > by_tile ? by_atom : by_tile);
  }
  w.x = c.take(n * 3 * 4);
  w.h_a = c.take(align_up(n, 128) * H * 4);   // whole 128-row blocks: intermediate layers keep h as a tile image (bf16 mode)
  w.h_b = c.take(align_up(n, 128) * H * 4);
  w.ab = c.take(n * 4 * H * 4);
  w.q = c.take(align_up(n, 128) * H * 4);   // whole 128-row blocks (tile image of the tcgen05 node kernel)
  w.agg = c.take(n * H * 4);
  w.vn = c.take(n * kVnRow * 4);
  w.bn_part_rows = kEdgeMaxCtas * kEdgeWarps;
  w.bn_part = c.take((size_t)w.bn_part_rows * 32 * 4);
  w.bn_param = c.take(32 * 4);
  w.vn_shape = c.take((size_t)d.layers * b * 96 * 4);
  if (d.hidden != 128) {
    const size_t m = n * KS;
    w.g_hid = c.take(m * H * 4); w.g_out = c.take(m * H * 4); w.g_alpha = c.take(m * kHeads * 4);
    w.g_rbf = c.take(m * kRbf * 4); w.g_rel = c.take(m * 3 * 4); w.g_idx = c.take(3 * m * 4);
    w.g_node = c.take(n * H * 4); w.g_bn = c.take(n * 32 * 4);
    {
      const int seg_k[4] = {kRbf, H, H, kShape};
      w.g_wimg_bytes = (size_t)tc_gemm_w_img_bytes(H, seg_k, 4, false);
      w.g_wimg = c.take(w.g_wimg_bytes);
    }
  }
  w.tiles = c.take(16 + (size_t)w.max_tiles * 16);
  w.tiles_s = c.take(16 + (size_t)w.max_tiles * 16);
  w.total = c.off;
  return w;
}

// ---- parameter enumeration --------------------------------------------------------------------
const std::vector<std::string>& param_names(const smb_model_dims& d) {
  static std::mutex mu;
  static std::map<int, std::vector<std::string>> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(d.layers);
  if (it != cache.end()) return it->second;
  std::vector<std::string> v;
  const char* six[6] = {"0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias"};
  auto add_mlp = [&](const std::string& p) { for (auto s : six) v.push_back(p + ".net." + s); };
  v.push_back("time_emb.1.weight"); v.push_back("time_emb.1.bias");
  v.push_back("time_emb.3.weight"); v.push_back("time_emb.3.bias");
  v.push_back("ligand_atom_emb.weight"); v.push_back("ligand_atom_emb.bias");
  add_mlp("refine_net.edge_pred_layer");
  add_mlp("refine_net.invariant_shape_layer.hidden_layer");
  v.push_back("v_inference.0.weight"); v.push_back("v_inference.0.bias");
  v.push_back("v_inference.2.weight"); v.push_back("v_inference.2.bias");
  for (int l = 0; l < d.layers; ++l) {
    std::string b = "refine_net.base_block." + std::to_string(l);
    for (auto m : {"hk_func", "hv_func", "hq_func", "node_output"}) add_mlp(b + ".x2h_layers.0." + m);
    for (auto m : {"xk_func", "xv_func", "xq_func"}) add_mlp(b + ".h2x_layers.0." + m);
    v.push_back(b + ".h2x_layers.0.shape_linear.map_to_feat.weight");
    v.push_back(b + ".h2x_layers.0.shape_linear.map_to_dir.weight");
  }
  return cache.emplace(d.layers, std::move(v)).first->second;
}

// ---- packing ----------------------------------------------------------------------------------
namespace {
inline uint16_t f2bf(float f) {   // round-to-nearest-even
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

// B fragments of mma.m16n8k16 for B[k][n] = W(n,k): per (n-tile, k-step, lane) the four values
// (k=2t,2t+1 | k=2t+8,2t+9) at n = 8*nt+g, as bf16 hi (and lo) pairs.
template <class F>
void pack_frags(uint8_t* dst, int NT, int KS, int prec, F W) {
  const size_t fb = prec == SMB_PREC_BF16X3 ? 16 : 8;
  for (int nt = 0; nt < NT; ++nt)
    for (int ks = 0; ks < KS; ++ks)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3, n = nt * 8 + g, k0 = ks * 16 + 2 * t;
        const float v[4] = {W(n, k0), W(n, k0 + 1), W(n, k0 + 8), W(n, k0 + 9)};
        uint16_t hi[4], lo[4];
        for (int i = 0; i < 4; ++i) { hi[i] = f2bf(v[i]); lo[i] = f2bf(v[i] - bf2f(hi[i])); }
        uint32_t* o = reinterpret_cast<uint32_t*>(dst + ((size_t)(nt * KS + ks) * 32 + lane) * fb);
        o[0] = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16);
        o[1] = (uint32_t)hi[2] | ((uint32_t)hi[3] << 16);
        if (prec == SMB_PREC_BF16X3) {
          o[2] = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16);
          o[3] = (uint32_t)lo[2] | ((uint32_t)lo[3] << 16);
        }
      }
}
inline void put(uint8_t* blob, size_t off, const float* src, size_t n) { memcpy(blob + off, src, n * 4); }
}  // namespace

static int pack_impl(const smb_model_dims& d, const float* const* hp, uint8_t* blob) {
  const int H = d.hidden, C = d.classes, TD = d.time_dim, prec = d.precision;
  const int KV = 2 * H + kRbf + kShape;   // 308: [r | h_dst | h_src | inv]
  const ModelLayout L = build_layout(d);
  memset(blob, 0, L.total);
  const std::vector<std::string>& names = param_names(d);
  std::map<std::string, const float*> P;
  for (size_t i = 0; i < names.size(); ++i) P[names[i]] = hp[i];
  auto get = [&](const std::string& k) -> const float* { return P.at(k); };

  // time embedding: frequencies exp(-j ln(1e4)/(half-1))  (models/molopt_score_model.py:159-166)
  {
    float* f = reinterpret_cast<float*>(blob + L.time_freq);
    const int half = TD / 2;
    const float e = (float)(log(10000.0) / (half - 1));
    for (int j = 0; j < half; ++j) f[j] = expf((float)j * -e);
    put(blob, L.time_w1, get("time_emb.1.weight"), 2 * TD * TD);
    put(blob, L.time_b1, get("time_emb.1.bias"), 2 * TD);
    put(blob, L.time_w2, get("time_emb.3.weight"), TD * 2 * TD);
    put(blob, L.time_b2, get("time_emb.3.bias"), TD);
  }
  {  // atom embedding, transposed to [in][H]
    const float* w = get("ligand_atom_emb.weight");
    float* o = reinterpret_cast<float*>(blob + L.emb_wT);
    const int in = C + TD;
    for (int c = 0; c < H; ++c) for (int k = 0; k < in; ++k) o[(size_t)k * H + c] = w[(size_t)c * in + k];
    put(blob, L.emb_b, get("ligand_atom_emb.bias"), H);
  }
  {
    const std::string p = "refine_net.invariant_shape_layer.hidden_layer.net.";
    put(blob, L.inv_w1, get(p + "0.weight"), kShape * kShape); put(blob, L.inv_b1, get(p + "0.bias"), kShape);
    put(blob, L.inv_g, get(p + "1.weight"), kShape); put(blob, L.inv_bb, get(p + "1.bias"), kShape);
    put(blob, L.inv_w2, get(p + "3.weight"), kShape * kShape); put(blob, L.inv_b2, get(p + "3.bias"), kShape);
  }
  if (H != 128) {   // generic-shape path: the kernels read the parameters as they are
    const size_t nn = names.size();
    for (size_t i = 0; i < nn; ++i) memcpy(blob + L.raw + raw_weight_offset(d, names[i]), hp[i], param_numel(d, names[i]) * 4);
    return 0;
  }
  // LayerNorm folding (warp-specialised edge pipeline).  For an edge MLP with first Linear W1 / b1 and LayerNorm
  // (gamma, beta):  LN(v)_c = gamma_c (v_c - mean) rstd + beta_c.  Exact algebra:
  //   * subtracting the mean over the hidden channels from every column of W1 (and from b1) makes mean = 0;
  //   * relu(gamma_c n_c + beta_c) = |gamma_c| relu(sign(gamma_c) n_c + beta_c / |gamma_c|): the sign goes into
  //     channel c of the first Linear (the variance does not see it), |gamma_c| into column c of the second Linear.
  // fold.f[c] = sign, fold.mag[c] = |gamma_c| (floored so that gamma = 0 stays finite), fold.beta[c] = beta_c / |gamma_c|.
  struct Fold { std::vector<float> f, mag, beta; };
  auto make_fold = [&](const std::string& p) {
    const float* g = get(p + ".net.1.weight");
    const float* b = get(p + ".net.1.bias");
    Fold fo;
    fo.f.resize(H); fo.mag.resize(H); fo.beta.resize(H);
    for (int c = 0; c < H; ++c) {
      fo.f[c] = g[c] < 0.f ? -1.f : 1.f;
      fo.mag[c] = fmaxf(fabsf(g[c]), 1e-20f);
      fo.beta[c] = b[c] / fo.mag[c];
    }
    return fo;
  };
  // column k of a [H x ld] row-major weight, centred over the H rows and sign-folded
  auto folded_col = [&](const float* w, int ld, int k, const Fold& fo, std::vector<float>& out) {
    double m = 0.0;
    for (int n = 0; n < H; ++n) m += w[(size_t)n * ld + k];
    m /= H;
    out.resize(H);
    for (int n = 0; n < H; ++n) out[n] = fo.f[n] * (float)((double)w[(size_t)n * ld + k] - m);
  };
  auto pack_edge = [&](const EdgeMlpOff& e, const std::string& p, int n2, int ld1, bool gate) {
    const float* w1 = get(p + ".net.0.weight");
    pack_frags(blob + e.w1r, H / 8, 2, prec, [&](int n, int k) { return k < kRbf ? w1[(size_t)n * ld1 + k] : 0.f; });
    put(blob, e.b1, get(p + ".net.0.bias"), H);
    put(blob, e.ln_g, get(p + ".net.1.weight"), H);
    put(blob, e.ln_b, get(p + ".net.1.bias"), H);
    const float* w2 = get(p + ".net.3.weight");
    if (gate) {
      put(blob, e.w2, w2, H);
      put(blob, e.b2, get(p + ".net.3.bias"), 1);
      // warp-specialised pipeline: LayerNorm-folded first Linear with the (centred, signed) bias in k = 20 (hi) / 21 (lo)
      const Fold fo = make_fold(p);
      const float* b1 = get(p + ".net.0.bias");
      uint16_t* f1 = reinterpret_cast<uint16_t*>(blob + e.w1r_f);
      std::vector<float> col;
      for (int k = 0; k < kRbf; ++k) {
        folded_col(w1, ld1, k, fo, col);
        for (int n = 0; n < H; ++n) f1[((size_t)(n / 8) * 512 + (size_t)k * 16 + (n % 8) * 2) / 2] = f2bf(col[n]);
      }
      double mb = 0.0;
      for (int n = 0; n < H; ++n) mb += b1[n];
      mb /= H;
      for (int n = 0; n < H; ++n) {
        const float b = fo.f[n] * (float)((double)b1[n] - mb);
        const uint16_t hi = f2bf(b);
        f1[((size_t)(n / 8) * 512 + (size_t)kRbf * 16 + (n % 8) * 2) / 2] = hi;
        f1[((size_t)(n / 8) * 512 + (size_t)(kRbf + 1) * 16 + (n % 8) * 2) / 2] = f2bf(b - bf2f(hi));
      }
      float* wv = reinterpret_cast<float*>(blob + e.w2_f);
      for (int n = 0; n < H; ++n) wv[n] = w2[n] * fo.mag[n];
      put(blob, e.beta_f, fo.beta.data(), H);
    } else {
      pack_frags(blob + e.w2, n2 / 8, H / 16, prec, [&](int n, int k) { return w2[(size_t)n * H + k]; });
      put(blob, e.b2, get(p + ".net.3.bias"), n2);
      uint16_t* u1 = reinterpret_cast<uint16_t*>(blob + e.w1r_u);
      for (int k = 0; k < kRbf; ++k)
        for (int n = 0; n < H; ++n) u1[((size_t)(n / 8) * 512 + (size_t)k * 16 + (n % 8) * 2) / 2] = f2bf(w1[(size_t)n * ld1 + k]);
      uint16_t* u2 = reinterpret_cast<uint16_t*>(blob + e.w2_u);
      for (int n = 0; n < n2; ++n)
        for (int k = 0; k < H; ++k)
          u2[((size_t)(n / 8) * 2048 + (size_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2] = f2bf(w2[(size_t)n * H + k]);
      const Fold fo = make_fold(p);
      uint16_t* f1 = reinterpret_cast<uint16_t*>(blob + e.w1r_f);
      std::vector<float> col;
      for (int k = 0; k < kRbf; ++k) {
        folded_col(w1, ld1, k, fo, col);
        for (int n = 0; n < H; ++n) f1[((size_t)(n / 8) * 512 + (size_t)k * 16 + (n % 8) * 2) / 2] = f2bf(col[n]);
      }
      uint16_t* f2 = reinterpret_cast<uint16_t*>(blob + e.w2_f);
      for (int n = 0; n < n2; ++n)
        for (int k = 0; k < H; ++k)
          f2[((size_t)(n / 8) * 2048 + (size_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2] = f2bf(w2[(size_t)n * H + k] * fo.mag[k]);
      if (n2 == H) {   // transposed image for the query fold: row = input channel m, k = output channel c
        uint16_t* fq = reinterpret_cast<uint16_t*>(blob + e.w2_q);
        for (int m = 0; m < H; ++m)
          for (int c = 0; c < H; ++c)
            fq[((size_t)(m / 8) * 2048 + (size_t)(c / 8) * 128 + (m % 8) * 16 + (c % 8) * 2) / 2] = f2bf(w2[(size_t)c * H + m] * fo.mag[m]);
      }
      put(blob, e.beta_f, fo.beta.data(), H);
    }
  };
  pack_edge(L.gate, "refine_net.edge_pred_layer", 0, kRbf, true);
  {  // head: v_inference
    const float* w1 = get("v_inference.0.weight");
    const float* w2 = get("v_inference.2.weight");
    pack_frags(blob + L.head.w1, H / 8, H / 16, prec, [&](int n, int k) { return w1[(size_t)n * H + k]; });
    put(blob, L.head.b1, get("v_inference.0.bias"), H);
    pack_frags(blob + L.head.w2, 2, H / 16, prec, [&](int n, int k) { return n < C ? w2[(size_t)n * H + k] : 0.f; });
    put(blob, L.head.b2, get("v_inference.2.bias"), C);
  }
  for (int l = 0; l < d.layers; ++l) {
    const LayerOff& y = L.layer[l];
    const std::string b = "refine_net.base_block." + std::to_string(l);
    const std::string x2h = b + ".x2h_layers.0.", h2x = b + ".h2x_layers.0.";
    pack_edge(y.hk, x2h + "hk_func", H, KV, false);
    pack_edge(y.hv, x2h + "hv_func", H, KV, false);
    pack_edge(y.xk, h2x + "xk_func", H, KV, false);
    pack_edge(y.xv, h2x + "xv_func", kHeads, KV, false);
    // node-level projections of the two edge MLPs' first Linear + the query MLP
    auto pack_pre = [&](const NodeMlpOff& n, const std::string& kf, const std::string& vf, const std::string& qf) {
      const float* wk = get(kf + ".net.0.weight"); const float* bk = get(kf + ".net.0.bias");
      const float* wv = get(vf + ".net.0.weight"); const float* bv = get(vf + ".net.0.bias");
      const float* wq = get(qf + ".net.0.weight"); const float* bq = get(qf + ".net.0.bias");
      auto W = [&](int n_, int k) -> float {
        const int blk = n_ / H, r = n_ % H;
        if (blk == 4) return k < H ? wq[(size_t)r * H + k] : 0.f;
        const float* w = blk < 2 ? wk : wv;
        if ((blk & 1) == 0)   // A: dst part + shape part
          return k < H ? w[(size_t)r * KV + kRbf + k] : w[(size_t)r * KV + kRbf + 2 * H + (k - H)];
        return k < H ? w[(size_t)r * KV + kRbf + H + k] : 0.f;   // B: src part
      };
      pack_frags(blob + n.w1, 5 * H / 8, (H + kShape) / 16, prec, W);
      float* b1 = reinterpret_cast<float*>(blob + n.b1);
      for (int r = 0; r < H; ++r) { b1[r] = bk[r]; b1[2 * H + r] = bv[r]; b1[4 * H + r] = bq[r]; }
      {  // LayerNorm-folded pass-through blocks (A_k | B_k | A_v | B_v); the query block is unchanged
        const Fold fk = make_fold(kf), fv = make_fold(vf);
        const int K1 = H + kShape;
        std::vector<float> wf((size_t)5 * H * K1);
        for (int blk = 0; blk < 5; ++blk)
          for (int k = 0; k < K1; ++k) {
            if (blk == 4) { for (int r = 0; r < H; ++r) wf[(size_t)(blk * H + r) * K1 + k] = W(blk * H + r, k); continue; }
            const Fold& fo = blk < 2 ? fk : fv;
            double m = 0.0;
            for (int r = 0; r < H; ++r) m += W(blk * H + r, k);
            m /= H;
            for (int r = 0; r < H; ++r) wf[(size_t)(blk * H + r) * K1 + k] = fo.f[r] * (float)((double)W(blk * H + r, k) - m);
          }
        pack_frags(blob + n.w1_f, 5 * H / 8, K1 / 16, prec, [&](int n_, int k) { return wf[(size_t)n_ * K1 + k]; });
        float* bf = reinterpret_cast<float*>(blob + n.b1_f);
        double mk = 0.0, mv = 0.0;
        for (int r = 0; r < H; ++r) { mk += bk[r]; mv += bv[r]; }
        mk /= H; mv /= H;
        for (int r = 0; r < H; ++r) {
          bf[r] = fk.f[r] * (float)((double)bk[r] - mk); bf[H + r] = 0.f;
          bf[2 * H + r] = fv.f[r] * (float)((double)bv[r] - mv); bf[3 * H + r] = 0.f;
          bf[4 * H + r] = bq[r];
        }
        // tcgen05 images (node_pre_tc5_kernel): chunk 0 = the query MLP's hidden block, folded with its own LayerNorm
        const Fold fq = make_fold(qf);
        uint16_t* wt = reinterpret_cast<uint16_t*>(blob + n.w1_t);
        auto put_t = [&](int chunk, int nn, int k, float v) {
          wt[((size_t)chunk * kNodeChunkBytes + (size_t)(nn / 8) * (kNodeKx / 8) * 128 + (size_t)(k / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2) / 2] = f2bf(v);
        };
        auto put_bias = [&](int chunk, int nn, float b) {
          const uint16_t hi = f2bf(b);
          wt[((size_t)chunk * kNodeChunkBytes + (size_t)(nn / 8) * (kNodeKx / 8) * 128 + (size_t)(K1 / 8) * 128 + (nn % 8) * 16) / 2] = hi;
          wt[((size_t)chunk * kNodeChunkBytes + (size_t)(nn / 8) * (kNodeKx / 8) * 128 + (size_t)(K1 / 8) * 128 + (nn % 8) * 16 + 2) / 2] = f2bf(b - bf2f(hi));
        };
        for (int blk = 0; blk < 4; ++blk)
          for (int r = 0; r < H; ++r) {
            for (int k = 0; k < K1; ++k) put_t(1 + blk, r, k, wf[(size_t)(blk * H + r) * K1 + k]);
            put_bias(1 + blk, r, bf[blk * H + r]);
          }
        {
          double mb = 0.0;
          for (int r = 0; r < H; ++r) mb += bq[r];
          mb /= H;
          for (int k = 0; k < K1; ++k) {
            double m = 0.0;
            for (int r = 0; r < H; ++r) m += W(4 * H + r, k);
            m /= H;
            for (int r = 0; r < H; ++r) put_t(0, r, k, fq.f[r] * (float)((double)W(4 * H + r, k) - m));
          }
          for (int r = 0; r < H; ++r) put_bias(0, r, fq.f[r] * (float)((double)bq[r] - mb));
        }
        const float* w2q = get(qf + ".net.3.weight");
        uint16_t* w2t = reinterpret_cast<uint16_t*>(blob + n.w2_t);
        for (int nn = 0; nn < H; ++nn)
          for (int k = 0; k < H; ++k)
            w2t[((size_t)(nn / 8) * 2048 + (size_t)(k / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2) / 2] = f2bf(w2q[(size_t)nn * H + k] * fq.mag[k]);
        put(blob, n.beta_t, fq.beta.data(), H);
      }
      put(blob, n.ln_g, get(qf + ".net.1.weight"), H);
      put(blob, n.ln_b, get(qf + ".net.1.bias"), H);
      const float* w2 = get(qf + ".net.3.weight");
      pack_frags(blob + n.w2, H / 8, H / 16, prec, [&](int n_, int k) { return w2[(size_t)n_ * H + k]; });
      put(blob, n.b2, get(qf + ".net.3.bias"), H);
    };
    pack_pre(y.x2h_pre, x2h + "hk_func", x2h + "hv_func", x2h + "hq_func");
    pack_pre(y.h2x_pre, h2x + "xk_func", h2x + "xv_func", h2x + "xq_func");
    {
      const std::string p = x2h + "node_output";
      const float* w1 = get(p + ".net.0.weight");
      const float* w2 = get(p + ".net.3.weight");
      pack_frags(blob + y.node_out.w1, H / 8, 2 * H / 16, prec, [&](int n, int k) { return w1[(size_t)n * 2 * H + k]; });
      put(blob, y.node_out.b1, get(p + ".net.0.bias"), H);
      put(blob, y.node_out.ln_g, get(p + ".net.1.weight"), H);
      put(blob, y.node_out.ln_b, get(p + ".net.1.bias"), H);
      pack_frags(blob + y.node_out.w2, H / 8, H / 16, prec, [&](int n, int k) { return w2[(size_t)n * H + k]; });
      put(blob, y.node_out.b2, get(p + ".net.3.bias"), H);
      // tcgen05 images (node_tc5_kernel<1>): LayerNorm-folded first Linear [128 n][272 k] with the bias in k = 256 / 257
      const Fold fo = make_fold(p);
      const float* b1 = get(p + ".net.0.bias");
      uint16_t* wt = reinterpret_cast<uint16_t*>(blob + y.node_out.w1_t);
      auto at = [&](int nn, int k) -> uint16_t& {
        return wt[((size_t)(nn / 8) * (kNodeOutKx / 8) * 128 + (size_t)(k / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2) / 2];
      };
      for (int k = 0; k < 2 * H; ++k) {
        double m = 0.0;
        for (int r = 0; r < H; ++r) m += w1[(size_t)r * 2 * H + k];
        m /= H;
        for (int r = 0; r < H; ++r) at(r, k) = f2bf(fo.f[r] * (float)((double)w1[(size_t)r * 2 * H + k] - m));
      }
      double mb = 0.0;
      for (int r = 0; r < H; ++r) mb += b1[r];
      mb /= H;
      for (int r = 0; r < H; ++r) {
        const float b = fo.f[r] * (float)((double)b1[r] - mb);
        const uint16_t hi = f2bf(b);
        at(r, 2 * H) = hi;
        at(r, 2 * H + 1) = f2bf(b - bf2f(hi));
      }
      uint16_t* w2t = reinterpret_cast<uint16_t*>(blob + y.node_out.w2_t);
      for (int nn = 0; nn < H; ++nn)
        for (int k = 0; k < H; ++k)
          w2t[((size_t)(nn / 8) * 2048 + (size_t)(k / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2) / 2] = f2bf(w2[(size_t)nn * H + k] * fo.mag[k]);
      put(blob, y.node_out.beta_t, fo.beta.data(), H);
    }
    put(blob, y.vn_feat, get(h2x + "shape_linear.map_to_feat.weight"), kHeads * kVnIn);
    put(blob, y.vn_dir, get(h2x + "shape_linear.map_to_dir.weight"), kHeads * kVnIn);
  }
  return 0;
}

}  // namespace smb

extern "C" {

int smb_abi_version(void) { return SMB_ABI_VERSION; }
const char* smb_last_error_string(void) { return smb::g_err; }

int smb_param_count(const smb_model_dims* dims) {
  if (!dims) return SMB_E_BADARG;
  int rc = smb::check_dims(*dims);
  if (rc) return rc;
  return (int)smb::param_names(*dims).size();
}
const char* smb_param_name(const smb_model_dims* dims, int i) {
  if (!dims || smb::check_dims(*dims)) return nullptr;
  const auto& v = smb::param_names(*dims);
  if (i < 0 || i >= (int)v.size()) return nullptr;
  return v[i].c_str();
}
size_t smb_packed_weights_bytes(const smb_model_dims* dims) {
  if (!dims || smb::check_dims(*dims)) return 0;
  return smb::build_layout(*dims).total;
}
int smb_pack_weights(const smb_model_dims* dims, const float* const* host_params, int n_params, void* packed_host,
                     size_t packed_bytes) {
  if (!dims || !host_params || !packed_host) { smb::set_error_msg("smb_pack_weights: null argument"); return SMB_E_BADARG; }
  int rc = smb::check_dims(*dims);
  if (rc) return rc;
  if (n_params != (int)smb::param_names(*dims).size()) { smb::set_error_msg("smb_pack_weights: wrong parameter count"); return SMB_E_BADARG; }
  for (int i = 0; i < n_params; ++i)
    if (!host_params[i]) { smb::set_error_msg("smb_pack_weights: null parameter pointer"); return SMB_E_BADARG; }
  if (packed_bytes < smb::build_layout(*dims).total) { smb::set_error_msg("smb_pack_weights: output buffer too small"); return SMB_E_BADARG; }
  return smb::pack_impl(*dims, host_params, reinterpret_cast<uint8_t*>(packed_host));
}
size_t smb_workspace_bytes(const smb_model_dims* dims, int32_t n_atoms, int32_t n_mols) {
  if (!dims || smb::check_dims(*dims)) return 0;
  return smb::build_workspace(*dims, n_atoms, n_mols).total;
}

}  // extern "C"
