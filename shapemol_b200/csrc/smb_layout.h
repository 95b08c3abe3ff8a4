// Packed-weight and workspace layout shared by the host packer and the kernel launchers.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/shapemol_b200.h"

namespace smb {

constexpr int kHeads = 16;
constexpr int kShape = SMB_SHAPE_DIM;       // shape_dim = shape_latent_dim = 32
constexpr int kRbf = SMB_N_RBF;             // 20 (padded to K=32 for the tensor-core first Linear)
constexpr int kVnIn = 1 + kHeads + kShape;  // 49 channels into shape_linear
constexpr int kVnStride = 49;               // row stride of the VN weights in the blob
constexpr int kMaxLayers = 16;

// Edge MLP (Linear(2H+20+32 -> H) -> LN -> ReLU -> Linear(H -> N2)) with its first Linear split:
//   W1 [r | h_dst | h_src | inv_dst]: the r part lives here (tensor-core B fragments, K padded to
//   32); the h_dst/inv part and the h_src part are node-level projections (NodeMlp below).
struct EdgeMlpOff {
  size_t w1r;    // B fragments [H/8][2][32] (uint4 hi/lo  or  uint2 hi)
  size_t b1;     // [H] fp32 (only used by the gate MLP: the other MLPs fold b1 into the A projection)
  size_t ln_g;   // [H]
  size_t ln_b;   // [H]
  size_t w2;     // B fragments [N2/8][H/16][32]   (gate: fp32 [H] vector)
  size_t b2;     // [N2] fp32
  // tcgen05 (UMMA) operand images, bf16, no swizzle (smb_tc.cuh): used by the plain-bf16 edge kernels
  size_t w1r_u;  // [32 k][H n] MN-major: byte(k, n) = (n/8)*512 + k*16 + (n%8)*2      (k >= 20 zero)
  size_t w2_u;   // [N2 n][H k] K-major:  byte(n, k) = (n/8)*2048 + (k/8)*128 + (n%8)*16 + (k%8)*2
  // LayerNorm-folded images for the warp-specialised pipeline (smb_edge_ws.cu; see fold_ln in smb_host.cu):
  // first-Linear columns centred over the 128 hidden channels and multiplied by sign(gamma), second-Linear
  // k-columns multiplied by |gamma|, so that the kernel's LayerNorm is  z = relu(v * rstd + beta / |gamma|)
  size_t w1r_f;  // as w1r_u
  size_t w2_f;   // as w2_u
  size_t beta_f; // [H] fp32
  // hk / xk only: the folded second Linear TRANSPOSED, [H m][H c] K-major (byte(m, c) as w2_u with n = m, k = c): A operand of
  // the per-tile query fold  M[m, (d, h)] = sum_{c in head h} W2[c, m] Q_d[c]  (ROLE_K of smb_edge_ws.cu)
  size_t w2_q;
};

// Node-level chain: Y1 = X W1^T + b1 ; first n_pass columns are written out as they are, the last H
// columns go through LN/ReLU (or shifted softplus) and a second Linear.
struct NodeMlpOff {
  size_t w1;     // B fragments [N1/8][K1/16][32]
  size_t b1;     // [N1]
  size_t ln_g;   // [H]
  size_t ln_b;   // [H]
  size_t w2;     // B fragments [N2/8][H/16][32]
  size_t b2;     // [N2]
  // x2h_pre / h2x_pre only: W1 / b1 with the four pass-through blocks LayerNorm-folded like EdgeMlpOff::w1r_f
  size_t w1_f;
  size_t b1_f;
  // x2h_pre / h2x_pre / node_out: tcgen05 operand images for node_tc5_kernel (smb_node_tc5.cu); node_out has the
  // hidden block only, [128 n][kNodeOutKx k].  Every block is
  // LayerNorm-folded (the four pass-through blocks with their edge MLP's LayerNorm, the hidden block with the query
  // MLP's); the bias rides in two extra K columns (bf16 hi | lo) against constant-one activations.
  size_t w1_t;   // 5 chunks [hidden | A_k | B_k | A_v | B_v] of [128 n][kNodeKx k] bf16, K-major:
                 //   byte(n, k) = (n/8)*(kNodeKx/8)*128 + (k/8)*128 + (n%8)*16 + (k%8)*2
  size_t w2_t;   // [128 n][128 k] K-major (as EdgeMlpOff::w2_u), k-columns multiplied by |gamma|
  size_t beta_t; // [H] beta / |gamma|
};
// h of the tcgen05 node kernels between layers (bf16 mode): tile image, float offset of (row r, column c) =
//   (r / 128) * 128 * 128 + (c / 4) * 128 * 4 + (r % 128) * 4 + c % 4     (the workspace slots are sized for whole 128-row blocks)
// q of the tcgen05 node kernel: bf16, pre-multiplied by log2(e) / sqrt(head_dim), as a chunk image -- one 16-byte chunk per
// (row, head):  byte offset of (row r, head h) = (r / 128) * 32768 + h * 2048 + (r % 128) * 16   (8 channels of the head)
constexpr int kQChunkBlockBytes = 128 * 128 * 2;
// alpha (softmax x gate) of the warp-specialised edge pipeline, per tile of <= 128 consecutive edge rows of one molecule:
//   [16 heads][parts] fp32, part pd = the tile's rows of its pd-th destination (cnt(pd) rows: a tile may begin and end inside a
//   destination), rup4(cnt) slots each (slots >= cnt are zero); head stride = sum of the part lengths <= 128 + 3 * 8.
//   float kAlphaSumOff   + h * 8 + pd        sum_j alpha of the part
//   float kAlphaSplitOff + w * 32 + h * 2    (max logit, sum of exponentials) of the tile's first (w = 0) / last (w = 1) part when
//                                            that destination continues in the neighbouring tile: such parts hold
//                                            exp2(logit - max) x gate, NOT normalised, and the consumers combine the two parts
constexpr int kAlphaHeadMax = 152;
constexpr int kAlphaSumOff = 16 * kAlphaHeadMax;
constexpr int kAlphaSplitOff = kAlphaSumOff + 16 * 8;
constexpr int kAlphaTileFloats = kAlphaSplitOff + 64;   // 2624 floats = 10496 bytes per tile (a 27-atom tile uses 2240)
constexpr int kNodeKx = 176;                            // 128 (h) + 32 (inv) + 16 (bias hi | bias lo | zeros)
constexpr int kNodeChunkBytes = 128 * kNodeKx * 2;      // 45056
constexpr int kNodeOutKx = 272;                         // node_output: 128 (agg) + 128 (h) + 16 (bias hi | bias lo | zeros)

struct LayerOff {
  EdgeMlpOff hk, hv, xk, xv;
  NodeMlpOff x2h_pre;   // X=[h|inv] (K1=H+32) -> [A_hk|B_hk|A_hv|B_hv | hq hidden] -> Q_h
  NodeMlpOff node_out;  // X=[agg|h] (K1=2H)  -> hidden -> h' (+h residual)
  NodeMlpOff h2x_pre;   // X=[h'|inv]         -> [A_xk|B_xk|A_xv|B_xv | xq hidden] -> Q_x
  size_t vn_feat;       // [16][49] fp32
  size_t vn_dir;        // [16][49] fp32
};

struct ModelLayout {
  size_t total;
  size_t time_freq;     // [time_dim/2]
  size_t time_w1, time_b1, time_w2, time_b2;
  size_t emb_wT;        // [classes+time_dim][H] (transposed ligand_atom_emb.weight)
  size_t emb_b;         // [H]
  size_t inv_w1, inv_b1, inv_g, inv_bb, inv_w2, inv_b2;   // invariant_shape_layer MLP 32->32->32
  EdgeMlpOff gate;      // edge_pred_layer
  NodeMlpOff head;      // v_inference: X=h (K1=H) -> H -> ssp -> classes (padded to 16)
  LayerOff layer[kMaxLayers];
  size_t raw;           // hidden != 128 only: raw fp32 copy of every parameter (smb_param_name order) for the generic path
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t frag_bytes(const smb_model_dims& d) { return d.precision == SMB_PREC_BF16X3 ? 16 : 8; }

ModelLayout build_layout(const smb_model_dims& d);
// generic-shape path (smb_generic.cu)
size_t param_numel(const smb_model_dims& d, const std::string& name);
size_t raw_weights_bytes(const smb_model_dims& d);
size_t raw_weight_offset(const smb_model_dims& d, const std::string& name);
const std::vector<std::string>& param_names(const smb_model_dims& d);
int check_dims(const smb_model_dims& d);

struct Workspace {
  size_t total;
  size_t tau;       // [B][8]
  size_t inv;       // [B][32]
  size_t nbr;       // [N][k+1] int32
  size_t deg;       // [N] int32
  size_t ew;        // [N][k+1]
  size_t alpha;     // [N][k+1][16]
  size_t x;         // [N][3]
  size_t h_a;       // [N][H]
  size_t h_b;       // [N][H]
  size_t ab;        // [N][4H]
  size_t q;         // [N][H]
  size_t agg;       // [N][H]
  size_t vn;        // [N][100]  (o_mean 3 | p 48 | d 48 | pad)
  size_t bn_part;   // [bn_part_rows][32]
  size_t bn_param;  // [32] scale | shift
  size_t vn_shape;  // [L][B][96] shape-embedding part of the VN linear maps
  // generic-shape path (hidden != 128): dense slot table e = atom * (k+1) + slot, M = N (k+1)
  size_t g_hid, g_out;     // [M][H]
  size_t g_alpha;          // [M][heads]
  size_t g_rbf, g_rel;     // [M][20], [M][3]
  size_t g_idx;            // int32 [3][M]: destination / source atom, molecule (-1: empty slot)
  size_t g_node, g_bn;     // [N][H], [N][32]
  size_t g_wimg, g_wimg_bytes;   // pre-split weight image of the GEMM in flight (smb_tc_gemm.cu)
  size_t tiles;     // [max_tiles] int4 tile descriptors, preceded by the tile count (16 bytes): whole destinations per tile (X2H block)
  size_t tiles_s;   // same, tiles that may begin / end inside a destination (gate, H2X block)
  int max_tiles;    // N/4 + B + 8: a tile holds >= 4 destination atoms unless it is the last of its molecule
  int bn_part_rows;
};
constexpr int kVnRow = 100;
constexpr int kEdgeMaxCtas = 148 * 2;
constexpr int kEdgeWarps = 16;
Workspace build_workspace(const smb_model_dims& d, int n_atoms, int n_mols);

}  // namespace smb
