// extern "C" entry points that enqueue work: kNN graph, one network evaluation, one reverse step.
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

static int check_batch(const smb_batch* b) {
  if (!b) { set_error_msg("null batch"); return SMB_E_BADARG; }
  if (b->n_atoms < 0 || b->n_mols < 0) { set_error_msg("negative batch size"); return SMB_E_BADARG; }
  if (b->n_atoms > 0 && (!b->mol_ptr || !b->atom_mol)) { set_error_msg("null mol_ptr / atom_mol"); return SMB_E_BADARG; }
  if (b->max_atoms_per_mol > SMB_MAX_ATOMS_PER_MOL) { set_error_msg("molecule larger than SMB_MAX_ATOMS_PER_MOL"); return SMB_E_TOOBIG; }
  return 0;
}

static inline const float* fptr(const void* base, size_t off) {
  return reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(base) + off);
}
static inline const void* vptr(const void* base, size_t off) {
  return reinterpret_cast<const unsigned char*>(base) + off;
}
template <class T>
static inline T* wptr(void* base, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(base) + off); }

static void fill_edge_weights(EdgeArgs& e, const void* blob, const EdgeMlpOff& o) {
  e.w1r = vptr(blob, o.w1r); e.b1 = fptr(blob, o.b1); e.ln_g = fptr(blob, o.ln_g); e.ln_b = fptr(blob, o.ln_b);
  e.w2 = vptr(blob, o.w2); e.b2 = fptr(blob, o.b2);
  e.w1r_u = vptr(blob, o.w1r_u); e.w2_u = vptr(blob, o.w2_u);
  e.w1r_f = vptr(blob, o.w1r_f); e.w2_f = vptr(blob, o.w2_f); e.beta_f = fptr(blob, o.beta_f); e.w2_q = vptr(blob, o.w2_q);
}
static void fill_node_weights(NodeArgs& n, const void* blob, const NodeMlpOff& o) {
  n.w1 = vptr(blob, o.w1); n.b1 = fptr(blob, o.b1); n.ln_g = fptr(blob, o.ln_g); n.ln_b = fptr(blob, o.ln_b);
  n.w2 = vptr(blob, o.w2); n.b2 = fptr(blob, o.b2);
  n.w1_t = vptr(blob, o.w1_t); n.w2_t = vptr(blob, o.w2_t); n.beta_t = fptr(blob, o.beta_t);
}

#define SMB_LAUNCH(expr)                                                                     \
  do {                                                                                       \
    int _rc = (expr);                                                                        \
    if (_rc != 0) { if (_rc > 0) set_error(#expr, (cudaError_t)_rc); return _rc; }           \
  } while (0)

static int forward_impl(const smb_model_dims& d, const void* blob, const smb_batch& b, const smb_forward_io& io, void* ws_base,
                        size_t ws_bytes, cudaStream_t st) {
  const int N = b.n_atoms, B = b.n_mols, H = d.hidden;
  const Workspace W = build_workspace(d, N, B);
  if (ws_bytes < W.total || !ws_base) { set_error_msg("smb_forward: workspace too small"); return SMB_E_BADARG; }
  if (!io.pos || !io.v || !io.shape || !io.t || !io.pred_pos || !io.pred_h || !io.pred_v) {
    set_error_msg("smb_forward: null io pointer");
    return SMB_E_BADARG;
  }
  for (int l = 0; l < d.layers; ++l)
    if (!io.bn_weight[l] || !io.bn_bias[l] || !io.bn_running_mean[l] || !io.bn_running_var[l]) {
      set_error_msg("smb_forward: null BatchNorm pointer");
      return SMB_E_BADARG;
    }
  if (N == 0) return 0;
  const ModelLayout L = build_layout(d);
  const int n_max = b.max_atoms_per_mol > 0 ? b.max_atoms_per_mol : SMB_MAX_ATOMS_PER_MOL;

  float* x = wptr<float>(ws_base, W.x);
  float* tau = wptr<float>(ws_base, W.tau);
  float* inv = wptr<float>(ws_base, W.inv);
  int* nbr = wptr<int>(ws_base, W.nbr);
  int* deg = wptr<int>(ws_base, W.deg);
  float* ew = wptr<float>(ws_base, W.ew);
  float* alpha = wptr<float>(ws_base, W.alpha);
  float* hbuf[2] = {wptr<float>(ws_base, W.h_a), wptr<float>(ws_base, W.h_b)};
  float* ab = wptr<float>(ws_base, W.ab);
  float* q = wptr<float>(ws_base, W.q);
  float* agg = wptr<float>(ws_base, W.agg);
  float* vn = wptr<float>(ws_base, W.vn);
  float* bn_part = wptr<float>(ws_base, W.bn_part);
  float* bn_param = wptr<float>(ws_base, W.bn_param);

  SMB_CUDA_OK(cudaMemcpyAsync(x, io.pos, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));

  int prof_i = 0;
  auto prof = [&](int cls, bool begin) -> int {
    if (io.prof_kernel != cls || !io.prof_events || prof_i >= io.prof_capacity) return 0;
    cudaError_t e = cudaEventRecord((cudaEvent_t)io.prof_events[2 * prof_i + (begin ? 0 : 1)], st);
    if (!begin) ++prof_i;
    return (int)e;
  };
#define SMB_TIMED(cls, expr) do { SMB_LAUNCH(prof(cls, true)); SMB_LAUNCH(expr); SMB_LAUNCH(prof(cls, false)); } while (0)

  PrepArgs pa;
  pa.n_mols = B; pa.t = io.t; pa.shape = io.shape;
  pa.time_freq = fptr(blob, L.time_freq); pa.time_w1 = fptr(blob, L.time_w1); pa.time_b1 = fptr(blob, L.time_b1);
  pa.time_w2 = fptr(blob, L.time_w2); pa.time_b2 = fptr(blob, L.time_b2);
  pa.inv_w1 = fptr(blob, L.inv_w1); pa.inv_b1 = fptr(blob, L.inv_b1); pa.inv_g = fptr(blob, L.inv_g);
  pa.inv_bb = fptr(blob, L.inv_bb); pa.inv_w2 = fptr(blob, L.inv_w2); pa.inv_b2 = fptr(blob, L.inv_b2);
  pa.tau = tau; pa.inv = inv;
  float* vn_shape = wptr<float>(ws_base, W.vn_shape);
  pa.n_layers = d.layers; pa.vn_shape = H == 128 ? vn_shape : nullptr; pa.do_shape = io.reuse_static ? 0 : 1;
  for (int l = 0; l < d.layers; ++l) { pa.vn_w[l][0] = fptr(blob, L.layer[l].vn_feat); pa.vn_w[l][1] = fptr(blob, L.layer[l].vn_dir); }
  SMB_LAUNCH(launch_prep(pa, st));

  EmbedArgs ea;
  ea.n_atoms = N; ea.H = H; ea.classes = d.classes; ea.v = io.v; ea.atom_mol = b.atom_mol; ea.tau = tau;
  ea.emb_wT = fptr(blob, L.emb_wT); ea.emb_b = fptr(blob, L.emb_b); ea.h = hbuf[0]; ea.h0 = io.h0;
  SMB_LAUNCH(launch_embed(ea, st));

  SMB_TIMED(SMB_PROF_KNN, launch_knn(x, b.mol_ptr, B, d.k, nbr, deg, st));
  if (io.nbr) SMB_CUDA_OK(cudaMemcpyAsync(io.nbr, nbr, (size_t)N * (d.k + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));

  if (H != 128) return forward_generic(d, blob, L, W, ws_base, b, io, st);   // generic-shape fp32 path

  // warp-specialised tcgen05 pipeline for the three attention roles (plain-bf16 mode, <= 32 atoms per molecule)
  const bool ws = edge_ws_supported(d, n_max);
  const bool node_tc5 = ws && node_tc5_supported(d, n_max);
  int4* tiles = wptr<int4>(ws_base, W.tiles + 16);
  int* n_tiles = wptr<int>(ws_base, W.tiles);
  // two static tile lists: whole destinations per tile for the X2H block (ROLE_V's epilogue pays for every extra part of a tile),
  // runs that may begin / end inside a destination -- fewer, fuller tiles -- for the gate and the H2X block
  int4* tiles_s = wptr<int4>(ws_base, W.tiles_s + 16);
  int* n_tiles_s = wptr<int>(ws_base, W.tiles_s);
  if (ws && !io.reuse_static) {
    SMB_LAUNCH(launch_build_tiles(b.mol_ptr, B, d.k, false, tiles, n_tiles, st));
    SMB_LAUNCH(launch_build_tiles(b.mol_ptr, B, d.k, true, tiles_s, n_tiles_s, st));
  }
  auto edge = [&](int role, const EdgeArgs& e, int* bn_rows) -> int {
    return ws ? launch_edge_ws(role, e, bn_rows, st) : launch_edge(d, role, e, bn_rows, st);
  };

  EdgeArgs eb;
  memset(&eb, 0, sizeof(eb));
  eb.n_mols = B; eb.n_max = n_max; eb.k = d.k; eb.mol_ptr = b.mol_ptr; eb.x = x; eb.nbr = nbr; eb.deg = deg;
  eb.ab = ab; eb.q = q; eb.ew_in = ew; eb.ew_out = ew; eb.alpha = alpha; eb.agg = agg; eb.shape = io.shape; eb.vn = vn;
  eb.bn_partial = bn_part;
  eb.abh = ab; eb.tiles = tiles; eb.n_tiles = n_tiles; eb.alpha_t = alpha;

  {  // global edge gate, computed once from the input coordinates (uni_transformer.py:507)
    EdgeArgs e = eb;
    fill_edge_weights(e, blob, L.gate);
    e.tiles = tiles_s; e.n_tiles = n_tiles_s;
    SMB_TIMED(SMB_PROF_GATE, edge(ROLE_GATE, e, nullptr));
  }

  int cur = 0;
  for (int l = 0; l < d.layers; ++l) {
    const LayerOff& y = L.layer[l];
    const bool last = l == d.layers - 1;
    float* h_in = hbuf[cur];
    float* h_out = last ? io.pred_h : hbuf[cur ^ 1];
    NodeArgs na;
    memset(&na, 0, sizeof(na));
    na.n_atoms = N; na.atom_mol = b.atom_mol;
    // ---- X2H ----
    {
      NodeArgs n = na;
      n.x_mode = XMODE_H_INV; n.act = ACT_LN_RELU; n.n_pass = 4 * H; n.n2 = H; n.n2_valid = H;
      n.xa = h_in; n.xb = inv; n.out1 = ab; n.out2 = q;
      n.xa_image = node_tc5 && l > 0;
      fill_node_weights(n, blob, y.x2h_pre);
      if (ws) {   // bf16 operand images and LayerNorm-folded projections for the warp-specialised edge pipeline
        n.out1_h = reinterpret_cast<uint32_t*>(ab); n.mol_ptr = b.mol_ptr;
        n.w1 = vptr(blob, y.x2h_pre.w1_f); n.b1 = fptr(blob, y.x2h_pre.b1_f);
      }
      SMB_TIMED(SMB_PROF_NODE_PRE, node_tc5 ? launch_node_pre_tc5(n, st) : launch_node_mlp(d, n, st));
    }
    {
      EdgeArgs e = eb;
      e.col_a = 0; e.col_b = H;
      fill_edge_weights(e, blob, y.hk);
      SMB_TIMED(SMB_PROF_EDGE_K, edge(ROLE_K, e, nullptr));
    }
    {
      EdgeArgs e = eb;
      e.col_a = 2 * H; e.col_b = 3 * H;
      fill_edge_weights(e, blob, y.hv);
      SMB_TIMED(SMB_PROF_EDGE_V, edge(ROLE_V, e, nullptr));
    }
    {
      NodeArgs n = na;
      n.x_mode = XMODE_AGG_H; n.act = ACT_LN_RELU; n.n_pass = 0; n.n2 = H; n.n2_valid = H;
      n.xa = agg; n.xb = h_in; n.residual = h_in; n.out2 = h_out;
      n.xb_image = node_tc5 && l > 0; n.out2_image = node_tc5 && !last;
      fill_node_weights(n, blob, y.node_out);
      SMB_TIMED(SMB_PROF_NODE_OUT, node_tc5 ? launch_node_out_tc5(n, st) : launch_node_mlp(d, n, st));
    }
    // ---- H2X (uses the updated h) ----
    {
      NodeArgs n = na;
      n.x_mode = XMODE_H_INV; n.act = ACT_LN_RELU; n.n_pass = 4 * H; n.n2 = H; n.n2_valid = H;
      n.xa = h_out; n.xb = inv; n.out1 = ab; n.out2 = q;
      n.xa_image = node_tc5 && !last;
      fill_node_weights(n, blob, y.h2x_pre);
      if (ws) {   // bf16 operand images and LayerNorm-folded projections for the warp-specialised edge pipeline
        n.out1_h = reinterpret_cast<uint32_t*>(ab); n.mol_ptr = b.mol_ptr;
        n.w1 = vptr(blob, y.h2x_pre.w1_f); n.b1 = fptr(blob, y.h2x_pre.b1_f);
      }
      SMB_TIMED(SMB_PROF_NODE_PRE, node_tc5 ? launch_node_pre_tc5(n, st) : launch_node_mlp(d, n, st));
    }
    {
      EdgeArgs e = eb;
      e.col_a = 0; e.col_b = H;
      fill_edge_weights(e, blob, y.xk);
      e.tiles = tiles_s; e.n_tiles = n_tiles_s;
      e.zero_ptr = vn; e.zero_stride = kVnRow; e.zero_off = 3; e.zero_len = 3 * kHeads;   // ... ROLE_XV: the o sums in the vn row
      SMB_TIMED(SMB_PROF_EDGE_K, edge(ROLE_K, e, nullptr));
    }
    int bn_rows = 0;
    {
      EdgeArgs e = eb;
      e.col_a = 2 * H; e.col_b = 3 * H;
      e.vn_feat = fptr(blob, y.vn_feat); e.vn_dir = fptr(blob, y.vn_dir);
      e.vn_shape = vn_shape + (size_t)l * (B > 0 ? B : 1) * 96;
      fill_edge_weights(e, blob, y.xv);
      e.tiles = tiles_s; e.n_tiles = n_tiles_s;
      SMB_TIMED(SMB_PROF_EDGE_XV, edge(ROLE_XV, e, &bn_rows));
      if (ws) SMB_LAUNCH(launch_xv_split_finish(e, bn_rows, &bn_rows, st));
    }
    {
      BnArgs bn;
      bn.training = io.training; bn.n_atoms = N; bn.rows = bn_rows; bn.partial = bn_part;
      bn.weight = io.bn_weight[l]; bn.bias = io.bn_bias[l];
      bn.running_mean = io.bn_running_mean[l]; bn.running_var = io.bn_running_var[l];
      bn.num_batches_tracked = io.bn_num_batches_tracked[l];
      bn.param = bn_param;
      SMB_LAUNCH(launch_bn_final(bn, st));
    }
    SMB_LAUNCH(launch_vn_apply(vn, bn_param, x, last ? io.pred_pos : nullptr, N, st));
    cur ^= 1;
  }
  {  // type head (molopt_score_model.py:305)
    NodeArgs n;
    memset(&n, 0, sizeof(n));
    n.n_atoms = N; n.atom_mol = b.atom_mol;
    n.x_mode = XMODE_H; n.act = ACT_SSP; n.n_pass = 0; n.n2 = 16; n.n2_valid = d.classes;
    n.xa = io.pred_h; n.out2 = io.pred_v;
    fill_node_weights(n, blob, L.head);
    SMB_TIMED(SMB_PROF_HEAD, launch_node_mlp(d, n, st));
  }
  return 0;
}

}  // namespace smb

extern "C" {

int smb_knn_graph(const float* x, const smb_batch* batch, int32_t k, int32_t* nbr, int32_t* deg, void* stream) {
  int rc = smb::check_batch(batch);
  if (rc) return rc;
  if (k < 1 || k > SMB_MAX_K) { smb::set_error_msg("smb_knn_graph: k out of range"); return SMB_E_TOOBIG; }
  if (batch->n_atoms == 0) return 0;
  if (!x || !nbr || !deg) { smb::set_error_msg("smb_knn_graph: null pointer"); return SMB_E_BADARG; }
  rc = smb::launch_knn(x, batch->mol_ptr, batch->n_mols, k, nbr, deg, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("knn_kernel launch", (cudaError_t)rc);
  return rc;
}

int smb_forward(const smb_model_dims* dims, const void* packed_weights_dev, const smb_batch* batch, const smb_forward_io* io,
                void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims || !packed_weights_dev || !io) { smb::set_error_msg("smb_forward: null argument"); return SMB_E_BADARG; }
  int rc = smb::check_dims(*dims);
  if (rc) return rc;
  rc = smb::check_batch(batch);
  if (rc) return rc;
  return smb::forward_impl(*dims, packed_weights_dev, *batch, *io, workspace, workspace_bytes, (cudaStream_t)stream);
}

int smb_type_head(const smb_model_dims* dims, const void* packed_weights_dev, const smb_batch* batch, const float* h,
                  float* logits, void* stream) {
  if (!dims || !packed_weights_dev || !h || !logits) { smb::set_error_msg("smb_type_head: null argument"); return SMB_E_BADARG; }
  int rc = smb::check_dims(*dims);
  if (rc) return rc;
  rc = smb::check_batch(batch);
  if (rc) return rc;
  const smb::ModelLayout L = smb::build_layout(*dims);
  if (dims->hidden != 128) return smb::type_head_generic(*dims, packed_weights_dev, L, batch->n_atoms, h, logits, (cudaStream_t)stream);
  smb::NodeArgs n;
  memset(&n, 0, sizeof(n));
  n.n_atoms = batch->n_atoms; n.atom_mol = batch->atom_mol;
  n.x_mode = smb::XMODE_H; n.act = smb::ACT_SSP; n.n_pass = 0; n.n2 = 16; n.n2_valid = dims->classes;
  n.xa = h; n.out2 = logits;
  smb::fill_node_weights(n, packed_weights_dev, L.head);
  rc = smb::launch_node_mlp(*dims, n, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("node_mlp_kernel launch", (cudaError_t)rc);
  return rc;
}

int smb_posterior_step(const smb_model_dims* dims, const smb_batch* batch, const smb_posterior_io* io, void* stream) {
  if (!dims || !io) { smb::set_error_msg("smb_posterior_step: null argument"); return SMB_E_BADARG; }
  int rc = smb::check_dims(*dims);
  if (rc) return rc;
  rc = smb::check_batch(batch);
  if (rc) return rc;
  if (batch->n_atoms == 0) return 0;
  if (!io->pred_pos || !io->pred_v || !io->t || !io->pos || !io->v || !io->posterior_mean_c0_coef || !io->posterior_mean_ct_coef ||
      !io->posterior_logvar || !io->log_alphas_v || !io->log_one_minus_alphas_v || !io->log_alphas_cumprod_v ||
      !io->log_one_minus_alphas_cumprod_v || ((io->noise_pos == nullptr) != (io->noise_u == nullptr))) {
    smb::set_error_msg("smb_posterior_step: null io pointer (noise_pos and noise_u must both be given or both be NULL)");
    return SMB_E_BADARG;
  }
  smb::PosteriorArgs a;
  a.n_atoms = batch->n_atoms; a.classes = dims->classes; a.timesteps = dims->timesteps; a.atom_mol = batch->atom_mol; a.t = io->t;
  a.pred_pos = io->pred_pos; a.pred_v = io->pred_v; a.pos = io->pos; a.v = io->v;
  a.noise_pos = io->noise_pos; a.noise_u = io->noise_u; a.log_v0 = io->log_v0; a.log_post = io->log_post;
  a.seed = io->seed; a.atom_offset = io->atom_offset;
  a.c0 = io->posterior_mean_c0_coef; a.ct = io->posterior_mean_ct_coef; a.logvar = io->posterior_logvar;
  a.log_a = io->log_alphas_v; a.log_1m_a = io->log_one_minus_alphas_v;
  a.log_ac = io->log_alphas_cumprod_v; a.log_1m_ac = io->log_one_minus_alphas_cumprod_v;
  rc = smb::launch_posterior(a, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("posterior_kernel launch", (cudaError_t)rc);
  return rc;
}

int smb_pointcloud_guidance(const smb_batch* batch, const smb_guidance_io* io, void* stream) {
  if (!io) { smb::set_error_msg("smb_pointcloud_guidance: null argument"); return SMB_E_BADARG; }
  int rc = smb::check_batch(batch);
  if (rc) return rc;
  if (batch->n_atoms == 0) return 0;
  if (!io->pos || !io->cloud || io->n_cloud < 0) { smb::set_error_msg("smb_pointcloud_guidance: null pos / cloud"); return SMB_E_BADARG; }
  if (!(io->radius > 0.0) || !(io->ratio >= 0.0 && io->ratio < 0.8)) { smb::set_error_msg("smb_pointcloud_guidance: radius must be > 0 and 0 <= ratio < 0.8"); return SMB_E_BADARG; }
  rc = smb::launch_guidance(*io, batch->n_atoms, (io->cloud_ptr || io->t) ? batch->atom_mol : nullptr, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("guidance_kernel launch", (cudaError_t)rc);
  return rc;
}

int smb_check_stability(const smb_batch* batch, const float* pos, const int32_t* elem, const int32_t* thr, const int32_t* allowed,
                        int32_t n_elem, int32_t hs, int32_t* nr_bonds, int32_t* stable_atoms, void* stream) {
  int rc = smb::check_batch(batch);
  if (rc) return rc;
  if (batch->n_mols == 0) return 0;
  if (!pos || !elem || !thr || !allowed || !stable_atoms || n_elem <= 0) { smb::set_error_msg("smb_check_stability: null pointer / empty tables"); return SMB_E_BADARG; }
  rc = smb::launch_stability(pos, batch->mol_ptr, batch->n_mols, elem, thr, allowed, n_elem, hs, nr_bonds, stable_atoms, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("stability_kernel launch", (cudaError_t)rc);
  return rc;
}

int smb_shape_tanimoto(const smb_batch* batch, const float* pos, const double* ref, const int32_t* ref_ptr, int32_t n_ref, double k,
                       double coef, double den, double* out, void* stream) {
  int rc = smb::check_batch(batch);
  if (rc) return rc;
  if (batch->n_mols == 0) return 0;
  if (!pos || !ref || !out || (!ref_ptr && n_ref <= 0)) { smb::set_error_msg("smb_shape_tanimoto: null pointer / empty reference"); return SMB_E_BADARG; }
  rc = smb::launch_tanimoto(pos, batch->mol_ptr, batch->n_mols, ref, ref_ptr, n_ref, k, coef, den, out, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("tanimoto_kernel launch", (cudaError_t)rc);
  return rc;
}

#ifdef SMB_DEBUG
// -DSMB_DEBUG builds only (not declared in include/shapemol_b200.h): role trace of the warp-specialised edge pipeline
__attribute__((visibility("default"))) int smb_debug_ws_trace(int64_t* host_out) { return smb::debug_ws_trace(reinterpret_cast<long long*>(host_out)); }
#endif

int smb_decrement_t(int32_t* t, int32_t n_mols, void* stream) {
  if (n_mols > 0 && !t) { smb::set_error_msg("smb_decrement_t: null pointer"); return SMB_E_BADARG; }
  int rc = smb::launch_decrement_t(t, n_mols, (cudaStream_t)stream);
  if (rc > 0) smb::set_error("decrement_t launch", (cudaError_t)rc);
  return rc;
}

}  // extern "C"
