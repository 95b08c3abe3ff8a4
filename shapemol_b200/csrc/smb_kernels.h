// Internal kernel argument blocks and launchers (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smb_layout.h"

namespace smb {

struct PrepArgs {
  int n_mols;
  const int* t;            // [B]
  const float* shape;      // [B,32,3]
  const float *time_freq, *time_w1, *time_b1, *time_w2, *time_b2;
  const float *inv_w1, *inv_b1, *inv_g, *inv_bb, *inv_w2, *inv_b2;
  float* tau;              // [B,8]
  float* inv;              // [B,32]
  // shape-embedding part of every layer's VN linear maps (BaseH2XAttLayer.shape_linear, uni_transformer.py:153-156):
  // vn_shape[l][m][feat | dir][16][3] = sum_c W_l[ch][17 + c] * shape[m][c][:]   (constant over the layer's atoms)
  int n_layers;
  const float* vn_w[kMaxLayers][2];   // map_to_feat / map_to_dir weights [16][49]
  float* vn_shape;         // [L][B][96] or NULL
  int do_shape;            // 0: only the time embedding (the shape-dependent outputs are still valid)
};

struct EmbedArgs {
  int n_atoms, H, classes;
  const int* v;
  const int* atom_mol;
  const float* tau;
  const float* emb_wT;
  const float* emb_b;
  float* h;
  float* h0;   // optional copy
};

struct BnArgs {
  int training, n_atoms, rows;
  const float* partial;   // [rows][32]
  const float* weight;
  const float* bias;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float* param;           // [32] scale | shift
};

struct PosteriorArgs {
  int n_atoms, classes, timesteps;
  const int* atom_mol;
  const int* t;
  const float* pred_pos;
  const float* pred_v;
  float* pos;
  int* v;
  const float* noise_pos;
  const float* noise_u;
  float* log_v0;
  float* log_post;
  uint64_t seed;
  int64_t atom_offset;
  const float *c0, *ct, *logvar, *log_a, *log_1m_a, *log_ac, *log_1m_ac;
};

// Node-level chain GEMM (smb_node_mlp.cu)
enum NodeXMode { XMODE_H_INV = 0, XMODE_AGG_H = 1, XMODE_H = 2 };
enum NodeAct { ACT_LN_RELU = 0, ACT_SSP = 1 };
struct NodeArgs {
  int n_atoms;
  int x_mode, act;
  int n_pass;              // number of leading stage-1 columns written to out1 (multiple of 64)
  int n2;                  // stage-2 outputs (multiple of 8), n2_valid of them stored
  int n2_valid;
  const float* xa;         // first input  [N][H]   (h, or agg for XMODE_AGG_H)
  const float* xb;         // second input: inv [B][32] (H_INV) or h [N][H] (AGG_H)
  const int* atom_mol;
  const void* w1; const float* b1;
  const float *ln_g, *ln_b;
  const void* w2; const float* b2;
  const float* residual;   // optional [N][n2]
  float* out1;             // [N][n_pass]
  uint32_t* out1_h;        // optional: the pass-through columns as bf16, in the per-molecule operand layout of the
                           // warp-specialised edge pipeline instead of out1: molecule block at byte a0 * 1024,
                           // [part 0..3][column group of 8][atom][8 columns]  (needs mol_ptr)
  const int* mol_ptr;
  float* out2;             // [N][n2_valid]
  const void *w1_t, *w2_t; const float* beta_t;   // tcgen05 operand images (NodeMlpOff::w1_t / w2_t / beta_t)
  // tcgen05 node kernels: activations in the tile image [128-row block][32 column groups][128 rows][4 floats] instead of
  // row-major [N][128] (warp-coalesced for thread-per-row accesses).  xa_image: XMODE_H_INV input h; xb_image: XMODE_AGG_H
  // input h (= residual); out2_image: XMODE_AGG_H output h' (XMODE_H_INV always writes q as an image)
  int xa_image, xb_image, out2_image;
  int dbg;                 // SMB_NODE_DBG bit mask (timing experiments only: results are wrong when set)
};
int launch_node_mlp(const smb_model_dims& d, const NodeArgs& a, cudaStream_t st);
// tcgen05 implementation of XMODE_H_INV for the warp-specialised edge pipeline (smb_node_tc5.cu): needs out1_h,
// mol_ptr, w1_t, w2_t, beta_t, b2
bool node_tc5_supported(const smb_model_dims& d, int n_max);
int launch_node_pre_tc5(const NodeArgs& a, cudaStream_t st);
// XMODE_AGG_H (node_output MLP + residual): needs w1_t, w2_t, beta_t, b2, residual
int launch_node_out_tc5(const NodeArgs& a, cudaStream_t st);

// Edge kernels (smb_edge_attn.cu)
enum EdgeRole { ROLE_GATE = 0, ROLE_K = 1, ROLE_V = 2, ROLE_XV = 3 };
struct EdgeArgs {
  int n_mols, n_max, k;
  const int* mol_ptr;
  const float* x;          // [N][3] current coordinates
  const int* nbr;          // [N][k+1]
  const int* deg;          // [N]
  const float* ab;         // [N][4H]; columns [col_a, col_a+H) = dst part, [col_b, col_b+H) = src part
  int col_a, col_b;
  const float* q;          // [N][H] fp32 (mma.sync kernels) / bf16 chunk image, pre-scaled (warp-specialised pipeline)   (ROLE_K)
  const float* ew_in;      // [N][k+1]         (ROLE_K: folded into alpha)
  float* ew_out;           // [N][k+1]         (ROLE_GATE)
  float* alpha;            // [N][k+1][16]     (ROLE_K out; ROLE_V / ROLE_XV in)
  float* agg;              // [N][H]           (ROLE_V out)
  // ROLE_XV
  const float* shape;      // [B][32][3]
  const float *vn_feat, *vn_dir;   // [16][49]
  const float* vn_shape;   // [B][96] this layer's slice of PrepArgs::vn_shape (warp-specialised pipeline)
  float* vn;               // [N][kVnRow]
  float* bn_partial;       // [gridDim.x*warps][32]
  // weights
  const void* w1r; const float* b1; const float *ln_g, *ln_b; const void* w2; const float* b2;
  const void* w1r_u; const void* w2_u;   // tcgen05 operand images (EdgeMlpOff::w1r_u / w2_u)
  const void* w1r_f; const void* w2_f; const float* beta_f;   // LayerNorm-folded images (EdgeMlpOff::w1r_f / w2_f / beta_f)
  const void* w2_q;        // EdgeMlpOff::w2_q (ROLE_K of the warp-specialised pipeline)
  // warp-specialised pipeline (smb_edge_ws.cu)
  const void* abh;         // bf16 node projections in per-molecule operand layout (node_mlp_kernel out1_h)
  const int4* tiles;       // static tile list (build_tiles_kernel)
  const int* n_tiles;
  float* alpha_t;          // [tile][kAlphaTileFloats] (smb_layout.h)  (ROLE_K out; ROLE_V / ROLE_XV in)
  // ROLE_K: rows that ROLE_V / ROLE_XV accumulate into from two tiles (a destination split between tiles) are cleared here:
  // floats [zero_off, zero_off + zero_len) of row `atom` of zero_ptr (row stride zero_stride)
  float* zero_ptr; int zero_stride, zero_off, zero_len;
  int dbg;                 // SMB_WS_DBG bit mask (timing experiments only: results are wrong when set)
};
// bn_rows_out (ROLE_XV): number of [32]-float rows of a.bn_partial the launch writes
int launch_edge(const smb_model_dims& d, int role, const EdgeArgs& a, int* bn_rows_out, cudaStream_t st);
// warp-specialised tcgen05 pipeline (smb_edge_ws.cu): plain-bf16 mode, molecules of <= 32 atoms, all three roles together
bool edge_ws_supported(const smb_model_dims& d, int n_max);
int launch_build_tiles(const int* mol_ptr, int n_mols, int k, bool split, int4* tiles, int* n_tiles, cudaStream_t st);
// ROLE_XV follow-up: VN maps + BatchNorm partial sums of the destinations whose rows were split between two tiles (their o sums
// were accumulated into the vn rows); writes *bn_rows_out further rows of a.bn_partial behind the bn_rows_in rows of the edge launch
int launch_xv_split_finish(const EdgeArgs& a, int bn_rows_in, int* bn_rows_out, cudaStream_t st);
int launch_edge_ws(int role, const EdgeArgs& a, int* bn_rows_out, cudaStream_t st);
int debug_ws_trace(long long* host_out);   // -DSMB_DEBUG builds: [15 events][128 tiles] clock64 stamps of CTA 0 (SMB_WS_DBG & 16)

// Gathered fp32 GEMM on tcgen05 with split-bf16 products (smb_tc_gemm.cu):
//   C[M x N] (+)= [A_0[idx_0[m]] | ... | A_{n-1}[idx_{n-1}[m]]] . W[N][K]^T + bias,   K = sum of the segments' k
struct TcGemmSeg {
  const float* a;      // [rows][lda] fp32
  const int* idx;      // row of `a` for output row m (NULL: m; negative: a zero row)
  long long lda;
  int k;               // columns of this segment
  int w_off;           // first column of W this segment multiplies
};
struct TcGemmArgs {
  TcGemmSeg seg[4];
  int n_segs;
  const float* W;      // [N][ldw] fp32
  long long ldw;
  int M, N;
  const float* bias;   // [N] or NULL
  int accumulate;      // C += ...
  float* C;
  long long ldc;
  long long a_batch, w_batch, c_batch, idx_batch;   // element strides between the grid.z batches (0: shared)
  int split3;          // three bf16 pieces per operand (fp32-level products, six MMAs per step) instead of two
  int vec, vec_c;      // set by the launcher: 16-byte aligned operand / output rows
  // optional scratch for the pre-split image of W (shared weights only, w_batch == 0): a small kernel splits W once per launch
  // and the GEMM's CTAs bulk-copy its chunks instead of each re-splitting the same tile; NULL / too small: in-kernel staging
  void* w_img;
  long long w_img_bytes;
  int use_img;         // set by the launcher
  // optional fused epilogue: y = relu(LayerNorm(C row) * ln_gamma + ln_beta) (eps 1e-5) instead of the plain row -- the Linear ->
  // LayerNorm -> ReLU head of the reference's MLP (models/common.py:47-67).  Needs N <= 256, N % 16 == 0, no accumulate.
  const float* ln_gamma; const float* ln_beta;
};
bool tc_gemm_can_fuse_ln(int N, bool accumulate);
long long tc_gemm_w_img_bytes(int N, const int* seg_k, int n_segs, bool split3);   // scratch needed for TcGemmArgs::w_img
int launch_tc_gemm(const TcGemmArgs& g, int n_batch, cudaStream_t st);

// generic-shape fp32 path (smb_generic.cu): hidden_dim != 128
int forward_generic(const smb_model_dims& d, const void* blob, const ModelLayout& L, const Workspace& W, void* ws_base, const smb_batch& b,
                    const smb_forward_io& io, cudaStream_t st);
int type_head_generic(const smb_model_dims& d, const void* blob, const ModelLayout& L, int N, const float* h, float* logits, cudaStream_t st);

int launch_prep(const PrepArgs& a, cudaStream_t st);
int launch_knn(const float* x, const int* mol_ptr, int n_mols, int k, int* nbr, int* deg, cudaStream_t st);
int launch_embed(const EmbedArgs& a, cudaStream_t st);
int launch_bn_final(const BnArgs& a, cudaStream_t st);
int launch_vn_apply(const float* vn, const float* bn_param, float* x, float* x_out, int n_atoms, cudaStream_t st);
int launch_posterior(const PosteriorArgs& a, cudaStream_t st);
int launch_decrement_t(int* t, int n, cudaStream_t st);
int launch_tanimoto(const float* pos, const int* mol_ptr, int n_mols, const double* ref, const int* ref_ptr, int n_ref, double k,
                    double coef, double den, double* out, cudaStream_t st);
int launch_stability(const float* pos, const int* mol_ptr, int n_mols, const int* elem, const int* thr, const int* allowed, int n_elem,
                     int hs, int* nr_bonds, int* stable_atoms, cudaStream_t st);
int launch_guidance(const smb_guidance_io& io, int n_atoms, const int* atom_mol, cudaStream_t st);

}  // namespace smb
