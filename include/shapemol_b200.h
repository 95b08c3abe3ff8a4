/*
 * shapemol_b200 -- C ABI of the B200-native (sm_100a) ShapeMol denoising-step library.
 *
 * The reference (Amelie-Schreiber/ShapeMol) is pure Python/PyTorch and has no FFI of its own; each
 * entry point below names the reference Python interface it replaces (file:line relative to the
 * reference root).  The Python host side (shapemol_b200/dropin/models/*.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a cudaError_t (>0) or a negative SMB_E_* code;
 *     smb_last_error_string() describes the most recent failure of the calling thread;
 *   - the CALLER owns every buffer: device pointers come from the caller's allocator (PyTorch), the
 *     library never allocates or frees device memory; scratch space is a caller-provided workspace
 *     sized by smb_workspace_bytes();
 *   - all work is enqueued asynchronously on the caller's stream (a cudaStream_t passed as void*),
 *     no host synchronisation: safe to capture into a CUDA graph.  The only process-wide state is a per-device cache of
 *     idempotent launch configuration (maximum dynamic shared memory per kernel, SM count) and the per-thread error string;
 *   - the caller selects the device (cudaSetDevice) before calling: kernels are configured and launched on the CURRENT
 *     device, which must own every pointer passed in;
 *   - no CPU fallback, no multi-backend dispatch: unsupported configuration => SMB_E_UNSUPPORTED.
 */
#ifndef SHAPEMOL_B200_H
#define SHAPEMOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB_ABI_VERSION 2

#if defined(__GNUC__)
#define SMB_API __attribute__((visibility("default")))
#else
#define SMB_API
#endif

enum {
  SMB_E_UNSUPPORTED = -1, /* configuration outside the fast path (SURVEY 5 "Config / flags") */
  SMB_E_BADARG = -2,      /* null / misaligned / too-small buffer                            */
  SMB_E_TOOBIG = -3       /* molecule larger than SMB_MAX_ATOMS_PER_MOL or k too large       */
};

#define SMB_MAX_ATOMS_PER_MOL 64
#define SMB_MAX_K 63
#define SMB_MAX_LAYERS 16
#define SMB_N_RBF 20
#define SMB_SHAPE_DIM 32

/* arithmetic of the edge/node MLP contractions (everything else is fp32) */
enum {
  SMB_PREC_BF16X3 = 0, /* split-bf16, 3 tensor-core products per fp32 product: fp32-parity mode (<=1e-3 rel) */
  SMB_PREC_BF16 = 1    /* plain bf16 operands, fp32 accumulate: throughput mode (parity stated separately)   */
};

/* Hyper-parameters of ScorePosNet3D / UniTransformerO2TwoUpdateGeneral
 * (models/molopt_score_model.py:171-283, models/uni_transformer.py:337-393). */
typedef struct smb_model_dims {
  int32_t hidden;     /* config.hidden_dim: 128 (256 for the stress config)      */
  int32_t heads;      /* config.n_heads: 16                                       */
  int32_t layers;     /* config.num_layers (num_blocks must be 1)                 */
  int32_t k;          /* config.knn                                               */
  int32_t classes;    /* ligand_atom_feature_dim: 15                              */
  int32_t time_dim;   /* config.time_emb_dim: 8                                   */
  int32_t timesteps;  /* config.num_diffusion_timesteps: 1000                     */
  int32_t precision;  /* SMB_PREC_*                                               */
} smb_model_dims;

/* One ragged batch of molecules: atoms of a molecule are contiguous (batch_ligand is sorted,
 * scripts/sample_diffusion.py:72).  All pointers are DEVICE pointers. */
typedef struct smb_batch {
  int32_t n_atoms;            /* N                                                            */
  int32_t n_mols;             /* B                                                            */
  int32_t max_atoms_per_mol;  /* upper bound on any molecule's atom count (<= SMB_MAX_ATOMS_PER_MOL) */
  const int32_t* mol_ptr;     /* [B+1] first atom of each molecule                            */
  const int32_t* atom_mol;    /* [N]   molecule of each atom (== batch_ligand as int32)       */
} smb_batch;

/* ---- introspection ------------------------------------------------------------------------ */
SMB_API int smb_abi_version(void);
SMB_API const char* smb_last_error_string(void);

/* ---- weights -------------------------------------------------------------------------------
 * Replaces: nn.Module parameter access inside ScorePosNet3D.forward (the reference multiplies by
 * the fp32 nn.Linear weights directly).  The host passes the fp32 tensors of the reference
 * state_dict in the order smb_param_name(i) enumerates; the library re-lays them out for the
 * tensor-core kernels (first-Linear split into r / dst / src / shape parts, bf16 hi+lo fragments).
 * BatchNorm running statistics are NOT packed: they are live device tensors passed per call. */
SMB_API int smb_param_count(const smb_model_dims* dims);
/* name of parameter i using the reference's own state_dict keys, e.g.
 * "refine_net.base_block.0.x2h_layers.0.hk_func.net.0.weight" */
SMB_API const char* smb_param_name(const smb_model_dims* dims, int i);
SMB_API size_t smb_packed_weights_bytes(const smb_model_dims* dims);
/* host_params[i]: HOST pointer to contiguous fp32 data of parameter i.  packed_host: HOST buffer of
 * smb_packed_weights_bytes(); the caller then copies it to the device. */
SMB_API int smb_pack_weights(const smb_model_dims* dims, const float* const* host_params, int n_params,
                     void* packed_host, size_t packed_bytes);

/* ---- workspace ------------------------------------------------------------------------------ */
SMB_API size_t smb_workspace_bytes(const smb_model_dims* dims, int32_t n_atoms, int32_t n_mols);

/* ---- kNN graph -------------------------------------------------------------------------------
 * Replaces: UniTransformerO2TwoUpdateGeneral._connect_edge -> torch_geometric.nn.knn_graph
 * (models/uni_transformer.py:466-468).  Dense output, no atomics: nbr[i*(k+1)+s] = molecule-local
 * index of the s-th nearest neighbour of atom i (ascending (d2, index), self removed), -1 padded;
 * deg[i] = number of valid slots (== min(k, n-1) except for >k coincident duplicates). */
SMB_API int smb_knn_graph(const float* x, const smb_batch* batch, int32_t k, int32_t* nbr, int32_t* deg, void* stream);

/* ---- one network evaluation -------------------------------------------------------------------
 * Replaces: ScorePosNet3D.forward (models/molopt_score_model.py:286-320) including
 * UniTransformerO2TwoUpdateGeneral.forward (models/uni_transformer.py:483-540). */
typedef struct smb_forward_io {
  const float* pos;        /* [N,3]   ligand_pos_perturbed                                   */
  const int32_t* v;        /* [N]     ligand_v_perturbed (class index)                       */
  const float* shape;      /* [B,32,3] ligand_shape                                          */
  const int32_t* t;        /* [B]     time_step                                              */
  float* pred_pos;         /* [N,3]   out: pred_ligand_pos (x0 prediction)                   */
  float* pred_h;           /* [N,H]   out: pred_ligand_h                                     */
  float* pred_v;           /* [N,C]   out: pred_ligand_v (type logits)                       */
  float* h0;               /* [N,H]   out, optional (NULL): embedding output (return_all)    */
  int32_t* nbr;            /* [N,k+1] out, optional (NULL): kNN table of this evaluation      */
  /* BatchNorm of BaseH2XAttLayer.shape_linear (models/uni_transformer.py:119), the live tensors of
   * layer l: [heads] each.  training != 0: batch statistics over all N atoms, running stats updated
   * in place (momentum 0.1, unbiased variance) and num_batches_tracked += 1;  training == 0: running
   * statistics are used. */
  const float* bn_weight[SMB_MAX_LAYERS];
  const float* bn_bias[SMB_MAX_LAYERS];
  float* bn_running_mean[SMB_MAX_LAYERS];
  float* bn_running_var[SMB_MAX_LAYERS];
  int64_t* bn_num_batches_tracked[SMB_MAX_LAYERS]; /* may be NULL */
  int32_t training;
  /* Optional per-kernel timing (bench.py's roofline line): when prof_kernel != 0 the library records
   * the cudaEvent_t pair prof_events[2i], prof_events[2i+1] on `stream` around its i-th launch of that
   * kernel class inside this call (i < prof_capacity).  Classes: SMB_PROF_*. */
  int32_t prof_kernel;
  int32_t prof_capacity;
  void* const* prof_events;
  /* != 0: the caller promises that `workspace` still holds the step-independent quantities of an earlier smb_forward call
   * with the SAME batch, shape tensor and weights (invariant shape embedding, shape part of the VN maps, edge tile list);
   * they are not recomputed.  The sampling loop sets it from its second step on (models/molopt_score_model.py:558-681 calls
   * the network 1000 times with the same condition shape). */
  int32_t reuse_static;
} smb_forward_io;

enum {
  SMB_PROF_NONE = 0, SMB_PROF_EDGE_K = 1, SMB_PROF_EDGE_V = 2, SMB_PROF_EDGE_XV = 3, SMB_PROF_NODE_PRE = 4,
  SMB_PROF_NODE_OUT = 5, SMB_PROF_GATE = 6, SMB_PROF_KNN = 7, SMB_PROF_HEAD = 8
};

SMB_API int smb_forward(const smb_model_dims* dims, const void* packed_weights_dev, const smb_batch* batch,
                const smb_forward_io* io, void* workspace, size_t workspace_bytes, void* stream);

/* Type head only: logits = v_inference(h).  Replaces: `self.v_inference(h)` applied to intermediate h
 * when return_all=True (models/molopt_score_model.py:315). */
SMB_API int smb_type_head(const smb_model_dims* dims, const void* packed_weights_dev, const smb_batch* batch,
                          const float* h, float* logits, void* stream);

/* ---- one reverse-diffusion update ----------------------------------------------------------------
 * Replaces: the posterior block of ScorePosNet3D.sample_diffusion (models/molopt_score_model.py:
 * 655-673): q_pos_posterior :400-404, log-categorical posterior :377-385, log_sample_categorical
 * :98-104.  Noise is either injected (parity mode: noise_pos = randn_like(pos), noise_u =
 * rand_like(logits), the reference's draw order) or generated in-kernel with Philox4x32-10 keyed by
 * (seed, global atom index + atom_offset, step) when both are NULL. */
typedef struct smb_posterior_io {
  const float* pred_pos;    /* [N,3]  x0 prediction                                           */
  const float* pred_v;      /* [N,C]  logits                                                  */
  const int32_t* t;         /* [B]    current time step per molecule                          */
  float* pos;               /* [N,3]  in: x_t, out: x_{t-1}                                   */
  int32_t* v;               /* [N]    in: v_t, out: v_{t-1}                                   */
  const float* noise_pos;   /* [N,3]  or NULL                                                 */
  const float* noise_u;     /* [N,C]  or NULL                                                 */
  float* log_v0;            /* [N,C]  out, optional: log_softmax(logits)   (v0_traj entry)    */
  float* log_post;          /* [N,C]  out, optional: log posterior         (vt_traj entry)    */
  uint64_t seed;            /* Philox key (used when noise_* are NULL)                        */
  int64_t atom_offset;      /* global index of local atom 0 (multi-GPU shards)                */
  /* the seven fp32 [timesteps] schedule tables of the reference state_dict */
  const float* posterior_mean_c0_coef;
  const float* posterior_mean_ct_coef;
  const float* posterior_logvar;
  const float* log_alphas_v;
  const float* log_one_minus_alphas_v;
  const float* log_alphas_cumprod_v;
  const float* log_one_minus_alphas_cumprod_v;
} smb_posterior_io;

SMB_API int smb_posterior_step(const smb_model_dims* dims, const smb_batch* batch, const smb_posterior_io* io,
                       void* stream);

/* t[b] -= 1 on the device (keeps the sampling loop free of host syncs / graph-capturable). */
SMB_API int smb_decrement_t(int32_t* t, int32_t n_mols, void* stream);

/* ---- point-cloud shape guidance ----------------------------------------------------
 * Replaces pointcloud_shape_guidance (models/molopt_score_model.py:699-740; called at :582-591 on the predicted x0
 * while t > grad_step): every atom whose mean distance to its 3 nearest cloud points exceeds `radius` is pulled
 * towards their centroid by a random fraction in [ratio, 0.8), up to 5 times, until that mean distance is below
 * `radius`.  The reference does this on the host (sklearn KDTree, numpy RNG, D2H + H2D per step); here it is one
 * kernel, brute-force 3-NN in float64 in the reference's evaluation order (bit-exact against it for the same scalars).
 *   pos        [N,3] fp32, in/out
 *   cloud      [M,3] fp64 (utils/shape.py:164-173 produces float64), cloud_ptr NULL: every atom uses all M points;
 *              else int32 [n_mols+1]: molecule m uses points cloud_ptr[m] .. cloud_ptr[m+1]-1 (needs batch->atom_mol)
 *   t          NULL, or int32 [n_mols]: atoms of molecule m are only touched when t[m] > grad_step (device-side test,
 *              so the call can sit inside a captured CUDA graph)
 *   u          NULL: Philox4x32-10 keyed by (seed, atom_offset + atom, t or step, iteration); else fp64 [5][N] scalars
 *              in [0,1) (parity tests; entry [j][i] is used iff atom i is still far in iteration j) */
typedef struct smb_guidance_io {
  float* pos;
  const double* cloud;
  const int32_t* cloud_ptr;
  int32_t n_cloud;
  const int32_t* t;
  int32_t grad_step;
  int32_t step;            /* Philox counter word when t == NULL */
  double radius;
  double ratio;            /* reference default 0.2 */
  const double* u;
  uint64_t seed;
  int64_t atom_offset;
} smb_guidance_io;
SMB_API int smb_pointcloud_guidance(const smb_batch* batch, const smb_guidance_io* io, void* stream);

/* ---- alignment-free shape Tanimoto ---------------------------------------------------
 * Replaces get_ROCS (utils/evaluation/shaep_utils.py:59-83) for a whole batch: molecule m's generated centres
 * pos[mol_ptr[m] .. mol_ptr[m+1]) against the reference centres ref[ref_ptr[m] .. ref_ptr[m+1]) (fp64 [R,3]);
 * ref_ptr NULL: every molecule is compared with all n_ref reference centres.
 *   out[m] = V_AB / (V_AA + V_BB - V_AB),  V_XY = sum_ij coef * exp(-k |x_i - y_j|^2) / den
 * (k, coef, den): the reference's float32 per-atom constants for prefactor 0.8 / alpha 0.81, see
 * oracle rocs_constants(); float64 accumulation in a fixed order (deterministic). */
/* ---- stability check -----------------------------------------------------------------
 * Replaces check_stability / get_bond_order (utils/evaluation/analyze.py:249-297) for a whole batch: bond order of every
 * atom pair of a molecule from its distance (float32, x100 = pm) against thr[k][e_i][e_j] = bond length + margin, k = single /
 * double / triple; an atom is stable iff allowed[e] >= sum of its bond orders > 0 (hs != 0: ==).
 *   elem   int32 [N] element index 0..n_elem-1;  thr int32 [3][n_elem][n_elem];  allowed int32 [n_elem]   (device pointers)
 *   nr_bonds int32 [N] out (may be NULL);  stable_atoms int32 [n_mols] out;  a molecule is stable iff stable_atoms[m] == n_m */
SMB_API int smb_check_stability(const smb_batch* batch, const float* pos, const int32_t* elem, const int32_t* thr, const int32_t* allowed,
                        int32_t n_elem, int32_t hs, int32_t* nr_bonds, int32_t* stable_atoms, void* stream);

SMB_API int smb_shape_tanimoto(const smb_batch* batch, const float* pos, const double* ref, const int32_t* ref_ptr, int32_t n_ref,
                       double k, double coef, double den, double* out, void* stream);

/* ---- VN-DGCNN shape encoder ---------------------------------------------------------------------
 * Replaces: VN_DGCNN_Encoder.forward (models/shape_pointcloud_modelAE.py:231-255).
 * clouds [B,P,3] fp32 -> latent [B,latent,3].  Weight pointers are DEVICE fp32 tensors taken from
 * the live module (the 4 DGCNN blocks are unregistered in the reference and always use batch
 * statistics). */
typedef struct smb_encoder_weights {
  int32_t hidden;      /* 128 */
  int32_t latent;      /* 32  */
  int32_t n_blocks;    /* 4   */
  int32_t num_k;       /* 20  */
  const float* conv_pos_feat;   /* [hidden,2]  */
  const float* conv_pos_dir;    /* [hidden,2]  */
  const float* conv_pos_bn_w;   /* [hidden]    */
  const float* conv_pos_bn_b;
  float* conv_pos_bn_rm;        /* running mean / var, updated in training mode */
  float* conv_pos_bn_rv;
  const float* block_feat[8];   /* [hidden,2*hidden] */
  const float* block_dir[8];
  const float* block_bn_w[8];
  const float* block_bn_b[8];
  float* block_bn_rm[8];
  float* block_bn_rv[8];
  const float* conv_c_feat;     /* [latent, n_blocks*hidden] */
  const float* conv_c_dir;      /* [1, n_blocks*hidden]      */
  const float* conv_c_bn_w;
  const float* conv_c_bn_b;
  float* conv_c_bn_rm;
  float* conv_c_bn_rv;
  int32_t training;             /* BN mode of conv_pos / conv_c (blocks always use batch stats) */
} smb_encoder_weights;

SMB_API size_t smb_encoder_workspace_bytes(const smb_encoder_weights* w, int32_t n_clouds, int32_t n_points);
SMB_API int smb_vn_dgcnn_encode(const smb_encoder_weights* w, const float* clouds, int32_t n_clouds, int32_t n_points,
                        float* latent, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHAPEMOL_B200_H */
