// Warp-specialised tcgen05 / TMEM pipeline for the fused edge kernels (plain-bf16 contraction mode,
// molecules of <= 32 atoms: the MOSES workload).  Same math as smb_edge_attn.cu
// (BaseX2HAttLayer / BaseH2XAttLayer, models/uni_transformer.py:48-162); this file changes the
// SCHEDULE: one persistent CTA per SM, a static list of 128-row tiles, and five roles that hand tiles
// to each other through mbarriers so that global-load latency, the two MMAs and the SIMT epilogues of
// different tiles overlap.
//
// Tile = up to 128 consecutive edge slots of ONE molecule (slot = destination * deg + neighbour; the kNN degree
// deg = min(k, n-1) is the same for every atom of a molecule) touching nd <= 8 destinations.  Two static lists, built once
// per batch by build_tiles_kernel: whole destinations per tile (X2H block: ROLE_K + ROLE_V), and equal runs that may begin
// and end inside a destination (gate, H2X block: ROLE_K + ROLE_XV) -- see struct Tile / TileWalk.
//
// Roles (warp ids per role: struct Ring; 28 warps, register budgets re-balanced per role with setmaxnreg):
//   P       4 warps (ROLE_XV: 2, two rows per thread): neighbour index, |x_i - x_j|, 20 RBFs, one-hot(i), one-hot(j)
//           -> A1[128 x 96] bf16 in a 2-slot smem ring
//   loader  1 warp (ROLE_XV: inside the GEMM1 warp): bulk copies of the molecule's bf16 projection tiles (3-slot ring)
//   GEMM1   1 warp: (SS)  D[128 x 128] = A1 . [W1r ; A-tile ; B-tile]         (K = 96)
//   LN      8 warps: D -> LayerNorm -> ReLU -> z (bf16; TMEM columns, or smem for ROLE_V); thread = (row, column half)
//   fold    ROLE_K: 1 issuer warp + the 4-warp conversion group: M = W2^T blockdiag_h(Q_d) per tile (a small extra MMA, then
//           fp32 accumulator -> bf16 operand), so that  logit[e, (d, h)] = z_e . M[:, (d, h)]  comes straight out of GEMM2
//   GEMM2   1 warp: ROLE_K (TS): D = z . M;  ROLE_XV (TS): D = z . W2^T;  ROLE_V (SS): D^T = W2 . z^T
//   E2      groups of 4 warps, tiles round-robin (ROLE_K / ROLE_V: 2 groups, ROLE_XV: 4):
//                      ROLE_K  a row reads the 16 logit columns of its destination; softmax over the destination's rows in the
//                              tile, x gate -> alpha (+ the part's max / sum when the destination continues in the next tile)
//                      ROLE_V  sum_j alpha (W2 z + b2), lane = channel, in-thread over the rows
//                      ROLE_XV alpha w (x_i - x_j) -> VN linear maps -> BatchNorm partial sums (split destinations: o sums added
//                              into the vn row, finished by xv_split_finish_kernel)
//
// TMEM (512 columns): ROLE_K D[2] at 0/128 (GEMM2 overwrites GEMM1's accumulator in place), z[2] at 256/320, the query
// fold's accumulator at 384.  alpha travels between the kernels tile-strided (smb_layout.h kAlphaTileFloats).
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
constexpr int K1 = 32 + 2 * G;         // 96
constexpr int A1_SBO = (K1 / 8) * 128; // 1536
constexpr int A1_BYTES = TM * K1 * 2;  // 24576
constexpr int AB_BYTES = 2 * G * H * 2;  // 16384: A-tile | B-tile, each [32 k][128 n] MN-major
constexpr int LS = 17;                 // padded row stride of per-row head scratch
constexpr int NDMAX = 8;               // destinations per tile: min(128 / deg, NDMAX) (the epilogues' staging areas and the
                                       // per-tile query operand of ROLE_K are sized for it; small-k graphs get 8 x deg-row tiles)
// 28 warps = 7 warpgroups, launched at 72 registers per thread and re-balanced per role with setmaxnreg (Regs<ROLE>):
// in ascending warp id = ascending issue priority:
//   ROLE_XV: WG0-3 E2 (four groups, 56) | WG4-5 LN (96) | WG6 P (2 warps, two rows per thread), GEMM1 issuer + loader, GEMM2
//            issuer (72, unchanged)
//   others : WG0 P (four warps, one row per thread, 64) | WG1-2 E2 (two groups; ROLE_K 64, ROLE_V 88) | WG3-4 LN (96) |
//            WG5 ROLE_K's conversion group (72; else idle, 24) | WG6 idle warp, ROLE_K's fold issuer, GEMM1 issuer + loader,
//            GEMM2 issuer (48)
// The producer is latency bound (one dependent chain per row): four warps halve its time per tile where warps are spare, and
// it runs a tile or more ahead, so it gets the lowest priority.
// (measured: ROLE_XV 208 us with LN 88 / E2 64, 193 us with 96 / 56; ROLE_K 261 us with LN 88 / E2 104, 242 us with 96 / 96)
constexpr int GRP_WARPS = 8, NG_MAX = 4, WARPS = 28, THREADS = WARPS * 32;
// The SM's warp arbiter favours the highest warp id among eligible warps (B300_MICROARCH.md "hi-wid-first"), so the two MMA
// issuers -- one thread each, on every tile's critical path -- get the top ids, then the producer, then the epilogues.
constexpr int TOP_WARP0 = 24, F_WARP = 25, MMA_WARP = 26, G2_WARP = 27;
constexpr int REGS_IDLE = 24;
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }
constexpr int BAR_LN = 1, BAR_E2 = 2, BAR_LNQ = 6;   // named barriers (0 is __syncthreads); E2 group g uses BAR_E2 + g, LN quadrant q BAR_LNQ + q
constexpr int E2_GRP_THREADS = 128;
constexpr int GRP_THREADS = GRP_WARPS * 32;
constexpr uint32_t TMEM_COLS = 512;

// mbarriers: A1 ring (2 slots), z buffers (2), D buffers (3)
// GEMM1's completion (one tcgen05.commit per tile) is waited on by the LayerNorm, by the producer (operand slot reuse) and by
// the projection loader (slot reuse): its barrier ring has NB1 = 12 slots -- a common multiple of every ring depth involved,
// and deep enough that no waiter can fall a whole ring behind (a parity wait two phases late would block for ever).
constexpr int NB1 = 12;
enum { B_A1_FULL = 0, B_Z_FULL = 4, B_D1_FULL = 6, B_D2_FULL = 18, B_E2_DONE = 30, B_AB_FULL = 42, B_D1_FREE = 45,
       // ROLE_K query fold: Q operand staged (2 slots), fold accumulator full / read, folded operand M ready
       B_QB_FULL = 47, B_DM_FULL = 49, B_DM_FREE = 50, B_BM_FULL = 51, N_BARS = 52 };
// TMEM accumulator rings (512 columns in all):
//   ROLE_K : 3 x 128 (GEMM2 overwrites GEMM1's accumulator in place) + z[2] x 64
//   ROLE_V : 3 x 128 in place (+ 128 spare columns: the epilogue reads 32 columns per destination whatever its degree);
//            z lives in shared memory
//   ROLE_XV: GEMM2 has 16 output columns, so it gets its own ring: D1 2 x 128 | z[2] x 64 | D2 8 x 16.  D1 is free again as
//            soon as the LayerNorm has read it, and a late epilogue no longer stalls GEMM1.
template <int ROLE> struct Ring {
  // epilogue groups: ROLE_V's shared memory (z^T operands) has room for two staging areas only
  static constexpr int NG = ROLE == ROLE_GATE ? 0 : ROLE == ROLE_XV ? 4 : 2;
  // register budgets (pool: 7 warpgroups x 72 = 504):  72 + 2 LN + NG E2 + (4 - NG) 24 <= 504
#ifndef SMB_K_LN
#define SMB_K_LN 96
#define SMB_K_E2 64
#define SMB_K_CV 72
#endif
#ifndef SMB_V_LN
#define SMB_V_LN 96
#define SMB_V_E2 88
#endif
#ifndef SMB_XV_LN
#define SMB_XV_LN 96
#define SMB_XV_E2 56
#endif
  static constexpr int REGS_LN = ROLE == ROLE_XV ? SMB_XV_LN : ROLE == ROLE_V ? SMB_V_LN : ROLE == ROLE_K ? SMB_K_LN : 96;
  static constexpr int REGS_E2 = ROLE == ROLE_XV ? SMB_XV_E2 : ROLE == ROLE_V ? SMB_V_E2 : ROLE == ROLE_K ? SMB_K_E2 : 96;
  static constexpr int REGS_CV = SMB_K_CV;   // ROLE_K: the conversion group (third E2 warpgroup)
  // producer: ROLE_XV has no spare warpgroup (2 warps beside the MMA issuers); the other roles use WG5
  static constexpr int P_WARPS = ROLE == ROLE_XV ? 2 : 4, P_WARP0 = ROLE == ROLE_XV ? TOP_WARP0 : 0;
  static constexpr int E2_WARP0 = ROLE == ROLE_XV ? 0 : 4, E2_WARPS = ROLE == ROLE_XV ? 16 : 8;
  static constexpr int LN_WARP0 = ROLE == ROLE_XV ? 16 : 12, CV_WARP0 = 20;
  static constexpr int P_ROWS = TM / (P_WARPS * 32);   // rows of a tile per producer thread
  static constexpr int REGS_P = 64, REGS_TOP = 48;
  static_assert(ROLE == ROLE_XV ? 72 + 2 * REGS_LN + 4 * REGS_E2 <= 504
                                : REGS_TOP + REGS_P + 2 * REGS_LN + NG * REGS_E2 + (ROLE == ROLE_K ? REGS_CV : REGS_IDLE) + (2 - NG) * REGS_IDLE <= 504,
                "register pool");
  // D2_FULL / E2_DONE mbarriers are indexed by tile % NB2 (the TMEM buffer by tile % ND2).  A parity wait is only
  // unambiguous if the waiter visits every phase of its barrier: an epilogue group sees tiles g, g + NG, ..., so NB2 must
  // be a multiple of both NG and ND2 (ROLE_K: 3 buffers, 4 groups -> 12 barriers)
  static constexpr int NB2 = ROLE == ROLE_K ? 2 : (ROLE == ROLE_XV || ROLE == ROLE_GATE) ? 8 : 6;
  static constexpr bool SEP = ROLE == ROLE_XV || ROLE == ROLE_GATE;   // ROLE_GATE has no GEMM2 at all
  static constexpr int ND1 = SEP ? 2 : ROLE == ROLE_V ? 3 : 2;   // ROLE_V: 3 x 128 + 128 spare columns that its 32-column reads may run into
  static constexpr int ND2 = SEP ? 8 : ND1;
  static constexpr uint32_t Z_COL = 256;
  static constexpr uint32_t DM_COL = 384;   // ROLE_K: accumulator of the query fold, [128 m][8 h + d]
  static constexpr uint32_t D2_COL = SEP ? 384 : 0, D2_STRIDE = SEP ? 16 : 128;
};

template <int ROLE>
struct Plan {
  static constexpr int o_bar = 0;                 // 48 mbarriers
  static constexpr int o_tmem = 448;
  static constexpr int o_vec = 512;               // ln_g | ln_b | b2   (3 x 128 floats)
  static constexpr int o_w1r = o_vec + 1536;      // 8192
  static constexpr int o_w2 = o_w1r + 8192;       // second Linear (ROLE_K: its transposed image, EdgeMlpOff::w2_q)
  static constexpr int w2_bytes = ROLE == ROLE_GATE ? 0 : ROLE == ROLE_XV ? kHeads * H * 2 : H * H * 2;
  static constexpr int o_a1 = o_w2 + w2_bytes;    // 2 slots
  static constexpr int o_ab = o_a1 + 2 * A1_BYTES;   // 3 slots
  static constexpr int o_stat = o_ab + 3 * AB_BYTES; // LN: float[2 buffers][2 halves][128]
  static constexpr int o_z = o_stat + 2048;          // ROLE_V: z^T operand, 2 x 32768
  static constexpr int o_e2 = o_z + (ROLE == ROLE_V ? 2 * TM * H * 2 : 0);
  // E2 scratch, one per group.  ROLE_V / ROLE_XV: one staging slot (the tile's alpha block; ROLE_XV: + shape float[96]), then
  // role scratch.  ROLE_K: logits float[128][17] | gate float[128]
  static constexpr int alpha_bytes = kAlphaTileFloats * 4;
  static constexpr int o_nbst = ROLE == ROLE_XV ? alpha_bytes + 384 : alpha_bytes;   // neighbour tiles' split statistics: 2 x 32 floats
  static constexpr int stage_bytes = ROLE == ROLE_K ? 0 : o_nbst + 256;
  static constexpr int e_stage = 0;
  static constexpr int e_ptab = stage_bytes;          // int4[NDMAX]: first row, row count, alpha offset of the tile's parts
  static constexpr int e_log = e_ptab + 128;          // ROLE_K logits / ROLE_XV w : float[128][17]
  static constexpr int e_ew = e_log + 8704;           // ROLE_K float[128]
  static constexpr int e_rel = e_log + 8704;          // ROLE_XV float4[128]
  static constexpr int e_o = e_rel + 2048;            // ROLE_XV float[8][16][4]
  static constexpr int e2_bytes = ROLE == ROLE_K ? e_ew + 512 : ROLE == ROLE_V ? e_log : e_o + 2048;
  static constexpr int NGP = ROLE == ROLE_XV ? 4 : 2;
  static constexpr int o_vnw = o_e2 + NGP * e2_bytes;     // ROLE_XV: vn_feat | vn_dir
  static constexpr int o_bn = o_vnw + (ROLE == ROLE_XV ? 2 * kHeads * kVnStride * 4 : 0);   // ROLE_XV: float[16 E2 warps][32] BatchNorm partial sums
  // ROLE_K: Q operand of the query fold (2 slots x 8 k-steps x [16 (hh, d) rows][16 channels] bf16) and the folded operand M
  // ([128 m][16 nd (d, h)] bf16, MN-major, column-group stride 2048)
  static constexpr int QB_BYTES = 8 * 512;
  static constexpr int o_qb = o_bn + (ROLE == ROLE_XV ? 4 * NG_MAX * 32 * 4 : 0);
  static constexpr int o_bm = o_qb + (ROLE == ROLE_K ? 2 * QB_BYTES : 0);
  static constexpr int total = o_bm + (ROLE == ROLE_K ? 2 * NDMAX * 2048 : 0);
  static_assert(total <= 227 * 1024, "shared memory budget");
  static_assert(N_BARS * 8 <= o_tmem, "mbarrier area");
  static_assert(o_e2 % 128 == 0 && stage_bytes % 16 == 0 && o_z % 128 == 0 && e2_bytes % 16 == 0 && o_qb % 128 == 0 && o_bm % 128 == 0, "alignment");
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The 20 Gaussian RBFs exp(-(d - mu_g)^2 / 2) of GaussianSmearing (models/common.py:11-28).  The centres are three
// uniform runs (1..3 step .25, 3..6 step .5, 6..10 step 1) plus mu = 0, so inside a run
//   E(mu + D) = E(mu) R(mu),  R(mu) = exp(D (d - mu) - D^2 / 2),  R(mu + D) = R(mu) exp(-D^2):
// 7 MUFU + 2 FMUL per further centre instead of 20 MUFU.  Each run restarts from a directly evaluated anchor (relative
// error <= ~2^-19, the operand is rounded to bf16 afterwards); d is clamped at 32, beyond which every RBF is 0 in fp32.
__device__ __forceinline__ void rbf20(float dist, float (&e)[20]) {
  constexpr float c = 0.72134752044448170f;   // log2(e) / 2
  const float d = fminf(dist, 32.f);
  e[0] = fast_ex2(-c * d * d);
  {
    const float t = d - 1.f;
    float E = fast_ex2(-c * t * t), R = fast_ex2(fmaf(0.5f * c, t, -0.0625f * c));
    e[1] = E;
#pragma unroll
    for (int g = 2; g <= 8; ++g) { E *= R; R *= 0.93941306281347578f; e[g] = E; }   // exp(-1/16)
  }
  {
    const float t = d - 3.f;
    float E = fast_ex2(-c * t * t), R = fast_ex2(fmaf(c, t, -0.25f * c));
    e[9] = E;
#pragma unroll
    for (int g = 10; g <= 14; ++g) { E *= R; R *= 0.77880078307140487f; e[g] = E; }  // exp(-1/4)
  }
  {
    const float t = d - 6.f;
    float E = fast_ex2(-c * t * t), R = fast_ex2(fmaf(2.f * c, t, -c));
    e[15] = E;
#pragma unroll
    for (int g = 16; g <= 19; ++g) { E *= R; R *= 0.36787944117144233f; e[g] = E; }  // exp(-1)
  }
}

// Tile = `nrows` <= 128 consecutive edge slots of ONE molecule (slot e = destination * deg + neighbour): it starts `s0` slots into
// destination d0 and touches nd <= NDMAX destinations ("parts"); the first and the last part may be incomplete -- the rest of
// such a destination is the last / first part of the neighbouring tile.
//   int4 = { first atom of the molecule, molecule | s0 << 24, n | d0 << 8 | nd << 16 | deg << 24, (65536 / deg + 1) | nrows << 20 }
struct Tile {
  int a0, mol, n, d0, nd, deg, s0, nrows;
  uint32_t recip;
  __device__ __forceinline__ explicit Tile(const int4& t)
      : a0(t.x), mol(t.y & 0xffffff), n(t.z & 0xff), d0((t.z >> 8) & 0xff), nd((t.z >> 16) & 0xff), deg((t.z >> 24) & 0xff),
        s0((t.y >> 24) & 0xff), nrows((t.w >> 20) & 0xff), recip((uint32_t)t.w & 0xfffffu) {}
  __device__ __forceinline__ int rows() const { return nrows; }
  __device__ __forceinline__ int dst_of(int r) const { return (int)(((uint32_t)(r + s0) * recip) >> 16); }   // part of row r < 128
  __device__ __forceinline__ int slot_of(int r, int pd) const { return r + s0 - pd * deg; }                   // its neighbour slot
  __device__ __forceinline__ int first(int pd) const { return max(pd * deg - s0, 0); }                        // first row of part pd
  __device__ __forceinline__ int count(int pd) const { return min((pd + 1) * deg - s0, nrows) - first(pd); }
  __device__ __forceinline__ bool head_split() const { return s0 > 0; }
  __device__ __forceinline__ bool tail_split() const { return s0 + nrows < nd * deg; }
  // alpha block: part pd starts at float aoff(pd) of a head's hstride() floats (smb_layout.h)
  __device__ __forceinline__ int aoff(int pd) const { return pd == 0 ? 0 : ((count(0) + 3) & ~3) + (pd - 1) * ((deg + 3) & ~3); }
  __device__ __forceinline__ int hstride() const { return aoff(nd - 1) + ((count(nd - 1) + 3) & ~3); }
};

// timing experiments (SMB_WS_DBG & 16): per-tile clock64 stamps of CTA 0's roles, read back with smb_debug_ws_trace
constexpr int TRACE_EVENTS = 20, TRACE_TILES = 128;
__device__ long long g_trace[TRACE_EVENTS][TRACE_TILES];
#define SMB_TRACE(ev, t, cond) do { if (SMB_DBG(a, 16) && (a.dbg >> 8) == ROLE && blockIdx.x == 0 && (t) < TRACE_TILES && (cond)) g_trace[ev][t] = clock64(); } while (0)

template <int ROLE>
__global__ void __launch_bounds__(THREADS, 1) edge_ws_kernel(EdgeArgs a) {
  using P = Plan<ROLE>;
  using R = Ring<ROLE>;
  constexpr int ND1 = R::ND1, ND2 = R::ND2, NB2 = R::NB2;
  static_assert(NB2 % ND2 == 0 && (R::NG == 0 || NB2 % R::NG == 0) && NB2 <= 12, "barrier ring");
  constexpr uint32_t Z_COL = R::Z_COL;
  constexpr int P_WARPS = R::P_WARPS, P_WARP0 = R::P_WARP0, P_ROWS = R::P_ROWS;
  constexpr int E2_WARP0 = R::E2_WARP0, E2_WARPS = R::E2_WARPS, LN_WARP0 = R::LN_WARP0, CV_WARP0 = R::CV_WARP0;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P::o_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::o_tmem);
  float* s_g = reinterpret_cast<float*>(smem + P::o_vec);
  float* s_be = s_g + H;
  float* s_b2 = s_be + H;
  unsigned char* s_w1r = smem + P::o_w1r;
  unsigned char* s_w2 = smem + P::o_w2;
  unsigned char* s_a1 = smem + P::o_a1;
  unsigned char* s_ab = smem + P::o_ab;
  float* s_vnw = reinterpret_cast<float*>(smem + P::o_vnw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KSTR = a.k + 1;
  const int n_tiles = *a.n_tiles;
  const int t_begin = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
  const int nt = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x) - t_begin;
  const int4* tiles = a.tiles + t_begin;

  // ---- once per CTA: weights -> smem, rings zeroed, TMEM allocation, mbarriers ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.w1r_f);
    uint4* dst = reinterpret_cast<uint4*>(s_w1r);
    for (int p = tid; p < 8192 / 16; p += THREADS) dst[p] = src[p];
    const uint4* s2 = reinterpret_cast<const uint4*>(ROLE == ROLE_K ? a.w2_q : a.w2_f);
    uint4* d2 = reinterpret_cast<uint4*>(s_w2);
    for (int p = tid; p < P::w2_bytes / 16; p += THREADS) d2[p] = s2[p];
    if (tid < H) {
      s_g[tid] = ROLE == ROLE_GATE ? reinterpret_cast<const float*>(a.w2_f)[tid] : 0.f;   // ROLE_GATE: second Linear (a vector) x |gamma|
      s_be[tid] = a.beta_f[tid];   // beta / |gamma| (LayerNorm folded into the weight images)
      s_b2[tid] = ROLE == ROLE_GATE ? a.b2[0] : ROLE == ROLE_XV ? (tid < kHeads ? a.b2[tid] : 0.f) : a.b2[tid];
    }
    if (ROLE == ROLE_XV)
      for (int p = tid; p < kHeads * kVnStride; p += THREADS) {
        s_vnw[p] = a.vn_feat[p];
        s_vnw[kHeads * kVnStride + p] = a.vn_dir[p];
      }
    // A1 and projection rings start at zero: rows / atoms that a tile does not use keep finite stale data
    // that no one-hot column of a valid row selects
    uint4* z0 = reinterpret_cast<uint4*>(s_a1);
    for (int p = tid; p < (2 * A1_BYTES + 3 * AB_BYTES) / 16; p += THREADS) z0[p] = make_uint4(0u, 0u, 0u, 0u);
    if (ROLE == ROLE_K) {   // Q operand: block diagonal, its zero chunks are never written again; M operand: finite everywhere
      uint4* q0 = reinterpret_cast<uint4*>(smem + P::o_qb);
      for (int p = tid; p < (2 * P::QB_BYTES + 2 * NDMAX * 2048) / 16; p += THREADS) q0[p] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (tid == 32) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar + B_A1_FULL + b, P_WARPS * 32);
      mbar_init(bar + B_Z_FULL + b, GRP_THREADS);
    }
    for (int b = 0; b < NB1; ++b) mbar_init(bar + B_D1_FULL + b, 1);
    for (int b = 0; b < 12; ++b) {
      mbar_init(bar + B_D2_FULL + b, 1);
      mbar_init(bar + B_E2_DONE + b, E2_GRP_THREADS);
    }
    for (int b = 0; b < 3; ++b) mbar_init(bar + B_AB_FULL + b, 1);
    for (int b = 0; b < 2; ++b) mbar_init(bar + B_D1_FREE + b, GRP_THREADS);
    for (int b = 0; b < 2; ++b) mbar_init(bar + B_QB_FULL + b, E2_GRP_THREADS);
    mbar_init(bar + B_DM_FULL, 1);
    mbar_init(bar + B_DM_FREE, E2_GRP_THREADS);
    mbar_init(bar + B_BM_FULL, E2_GRP_THREADS);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (ROLE == ROLE_V && warp < 4) {
    // the ROLE_V epilogue reads up to 3 columns past a tile's last row (zero alpha weights): they must hold finite values from
    // the start, in the accumulators not yet written and in the spare columns behind the ring
    uint32_t zr[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) zr[q] = 0u;
    for (int c0 = 0; c0 < (int)TMEM_COLS; c0 += 32) tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + c0, zr);
    wait_st();
    fence_before_sync();
  }
  if (ROLE == ROLE_V) { __syncthreads(); fence_after_sync(); }

  // projection tiles of tile t's molecule: two bulk copies (dst part | src part, n * 256 bytes each) into slot t % 3, in the
  // MN-major operand layout with an n * 16-byte column-group stride (written like that by node_tc5_kernel); one thread
  auto load_ab = [&](int t, const int4& td) {
    if (t < nt) {
      const Tile T(td);
      unsigned char* dst = s_ab + (t % 3) * AB_BYTES;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.abh) + (size_t)T.a0 * (4 * H * 2);
      const uint32_t bytes = (uint32_t)T.n * (H * 2);
      uint64_t* fb = bar + B_AB_FULL + t % 3;
      mbar_arrive_expect_tx(fb, 2 * bytes);
      bulk_g2s(dst, src + (size_t)(a.col_a / H) * bytes, bytes, fb);
      bulk_g2s(dst + G * H * 2, src + (size_t)(a.col_b / H) * bytes, bytes, fb);
    }
  };

  // every role branch starts with its warpgroups' register re-balancing (setmaxnreg is warpgroup-aligned)
  if (ROLE != ROLE_XV && warp >= TOP_WARP0) reg_dec<R::REGS_TOP>();   // WG6: two idle warps + the two MMA issuers
  if (warp >= P_WARP0 && warp < P_WARP0 + P_WARPS) {
    if (ROLE != ROLE_XV) reg_dec<R::REGS_P>();
    const int ptid = tid - P_WARP0 * 32;
    // =====================================================================================
    // P: A1 operand.  Row r of the tile is edge (i <- j): i = d0 + r / deg, slot s = r % deg.
    // =====================================================================================
    // thread `ptid` builds row ptid (ROLE_XV: rows ptid and ptid + 64).  Three-deep software pipeline: while tile t is written, the
    // coordinates of tile t + 1, the neighbour indices of tile t + 2 and the descriptor of tile t + 3 are in flight.
    // Straight-line code: rows beyond the tile's last one repeat that row (clamped indices, finite operand rows that nothing
    // reads), so the two rows of a thread form one basic block and their dependency chains interleave.
    uint32_t oi_cur[P_ROWS], oj_cur[P_ROWS], oi_oth[P_ROWS], oj_oth[P_ROWS];   // byte offsets of the one-hot ones: this slot / the other
#pragma unroll
    for (int u = 0; u < P_ROWS; ++u) { oi_cur[u] = oi_oth[u] = 4 * 128; oj_cur[u] = oj_oth[u] = 8 * 128; }
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 td_a = nt > 0 ? __ldg(tiles) : zero4, td_b = nt > 1 ? __ldg(tiles + 1) : zero4, td_c = nt > 2 ? __ldg(tiles + 2) : zero4;
    int j_a[P_ROWS], j_b[P_ROWS];
    float xr[P_ROWS][6];
    auto fetch_j = [&](const Tile& N, int (&j)[P_ROWS]) {
      const int last = max(N.rows() - 1, 0);
#pragma unroll
      for (int u = 0; u < P_ROWS; ++u) {
        const int r = min(ptid + u * (P_WARPS * 32), last);
        const int il = N.dst_of(r);
        j[u] = __ldg(a.nbr + (size_t)(N.a0 + N.d0 + il) * KSTR + N.slot_of(r, il));   // consumed an iteration later (-1 only in a single-atom molecule)
      }
    };
    auto fetch_x = [&](const Tile& N, const int (&j)[P_ROWS]) {
      const int last = max(N.rows() - 1, 0);
#pragma unroll
      for (int u = 0; u < P_ROWS; ++u) {
        const int r = min(ptid + u * (P_WARPS * 32), last);
        const float* xi = a.x + (size_t)(N.a0 + N.d0 + N.dst_of(r)) * 3;
        const float* xj = a.x + (size_t)(N.a0 + max(j[u], 0)) * 3;
        xr[u][0] = __ldg(xi); xr[u][1] = __ldg(xi + 1); xr[u][2] = __ldg(xi + 2);
        xr[u][3] = __ldg(xj); xr[u][4] = __ldg(xj + 1); xr[u][5] = __ldg(xj + 2);
      }
    };
    if (nt > 0) { fetch_j(Tile(td_a), j_a); fetch_x(Tile(td_a), j_a); }
    if (nt > 1) fetch_j(Tile(td_b), j_b);
#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      const Tile T(td_a);
      const int slot = t & 1;
      const int last = max(T.rows() - 1, 0);
      int i[P_ROWS], j[P_ROWS];
      float dist[P_ROWS];
#pragma unroll
      for (int u = 0; u < P_ROWS; ++u) {
        j[u] = max(j_a[u], 0);
        i[u] = T.d0 + T.dst_of(min(ptid + u * (P_WARPS * 32), last));
        const float rx = xr[u][0] - xr[u][3], ry = xr[u][1] - xr[u][4], rz = xr[u][2] - xr[u][5];
        dist[u] = sqrtf(rx * rx + ry * ry + rz * rz);
      }
      // next loads (consumed one / two / three iterations from now)
      td_a = td_b; td_b = td_c;
#pragma unroll
      for (int u = 0; u < P_ROWS; ++u) j_a[u] = j_b[u];
      if (t + 1 < nt) fetch_x(Tile(td_a), j_a);
      if (t + 2 < nt) fetch_j(Tile(td_b), j_b);
      if (t + 3 < nt) td_c = __ldg(tiles + t + 3);
      SMB_TRACE(13, t, ptid == 0);
      if (t >= 2) mbar_wait(bar + B_D1_FULL + (t - 2) % NB1, ((t - 2) / NB1) & 1);   // GEMM1 of tile t - 2 has read this slot
      SMB_TRACE(14, t, ptid == 0);
      if (!(SMB_DBG(a, 4) && t >= 2)) {
        float e[P_ROWS][20];
#pragma unroll
        for (int u = 0; u < P_ROWS; ++u) rbf20(dist[u], e[u]);
#pragma unroll
        for (int u = 0; u < P_ROWS; ++u) {
          const int r = ptid + u * (P_WARPS * 32);
          unsigned char* arow = s_a1 + slot * A1_BYTES + (r >> 3) * A1_SBO + (r & 7) * 16;
          *reinterpret_cast<uint4*>(arow) = make_uint4(pack_bf16(e[u][0], e[u][1]), pack_bf16(e[u][2], e[u][3]), pack_bf16(e[u][4], e[u][5]), pack_bf16(e[u][6], e[u][7]));
          *reinterpret_cast<uint4*>(arow + 128) =
              make_uint4(pack_bf16(e[u][8], e[u][9]), pack_bf16(e[u][10], e[u][11]), pack_bf16(e[u][12], e[u][13]), pack_bf16(e[u][14], e[u][15]));
          // ROLE_GATE: k = 20, 21 are constant ones against the bias rows (hi | lo) of the folded first Linear
          *reinterpret_cast<uint4*>(arow + 256) = make_uint4(pack_bf16(e[u][16], e[u][17]), pack_bf16(e[u][18], e[u][19]), ROLE == ROLE_GATE ? 0x3F803F80u : 0u, 0u);
          // one-hot(dst) in k = 32..63, one-hot(src) in k = 64..95: clear this row's previous one, set the new one
          if (ROLE != ROLE_GATE) {
            const uint32_t oi = (uint32_t)((4 + (i[u] >> 3)) * 128 + (i[u] & 7) * 2);
            const uint32_t oj = (uint32_t)((8 + (j[u] >> 3)) * 128 + (j[u] & 7) * 2);
            *reinterpret_cast<uint16_t*>(arow + oi_cur[u]) = 0;
            *reinterpret_cast<uint16_t*>(arow + oj_cur[u]) = 0;
            *reinterpret_cast<uint16_t*>(arow + oi) = 0x3F80;
            *reinterpret_cast<uint16_t*>(arow + oj) = 0x3F80;
            // the slots alternate: what was just written becomes "the other slot" of the next tile
            oi_cur[u] = oi_oth[u]; oj_cur[u] = oj_oth[u];
            oi_oth[u] = oi; oj_oth[u] = oj;
          }
        }
      }
      fence_async_smem();
      SMB_TRACE(0, t, ptid == 0);
      mbar_arrive(bar + B_A1_FULL + slot);
    }
  } else if (warp >= LN_WARP0 && warp < LN_WARP0 + GRP_WARPS) {
    reg_inc<R::REGS_LN>();
    // =====================================================================================
    // LN: D -> LayerNorm -> ReLU -> z (bf16).  thread = (row, column half); every tile.
    // =====================================================================================
    const int gw = warp - LN_WARP0;
    const int half = gw >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
    float* s_stat = reinterpret_cast<float*>(smem + P::o_stat);
#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      const int b3 = t % ND1, zb = t & 1;
      mbar_wait(bar + B_D1_FULL + t % NB1, (t / NB1) & 1);
      fence_after_sync();
      SMB_TRACE(2, t, gw == 0 && lane == 0);
      uint32_t v[64];
      tmem_ld32(lane_addr + b3 * 128 + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(lane_addr + b3 * 128 + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      wait_ld();
      if (R::SEP) { fence_before_sync(); mbar_arrive(bar + B_D1_FREE + b3); }   // D1 may be overwritten by GEMM1(t + 2)
      // LayerNorm with the affine part folded into the operands (smb_host.cu fold_ln): the accumulator is already
      // centred over the 128 channels and carries sign(gamma), so  z = relu(v * rstd + beta / |gamma|)
      const float sq = ln_sumsq64(v);
      float* st = s_stat + zb * (2 * TM);          // double-buffered: a fast thread may already be one tile ahead
      st[half * TM + r] = sq;
      // the two column halves of a row live in warps (qd) and (qd + 4): a 64-thread barrier per TMEM lane quadrant
      named_sync(BAR_LNQ + qd, 64);
      const float rstd = rsqrtf((sq + st[(half ^ 1) * TM + r]) * (1.f / H) + 1e-5f);
      if (ROLE == ROLE_GATE) {
        // e_w = sigmoid(w2 . relu(LN(.)) + b2)   (uni_transformer.py:475-481): each half-row thread dots its 64 columns
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_be + half * 64 + e);
          const float4 ww = *reinterpret_cast<const float4*>(s_g + half * 64 + e);
          dot = fmaf(fmaxf(fmaf(__uint_as_float(v[e]), rstd, bb.x), 0.f), ww.x, dot);
          dot = fmaf(fmaxf(fmaf(__uint_as_float(v[e + 1]), rstd, bb.y), 0.f), ww.y, dot);
          dot = fmaf(fmaxf(fmaf(__uint_as_float(v[e + 2]), rstd, bb.z), 0.f), ww.z, dot);
          dot = fmaf(fmaxf(fmaf(__uint_as_float(v[e + 3]), rstd, bb.w), 0.f), ww.w, dot);
        }
        float* sd = reinterpret_cast<float*>(smem + P::o_e2) + zb * (2 * TM);
        sd[half * TM + r] = dot;
        named_sync(BAR_LNQ + qd, 64);
        if (half == 0) {
          const Tile T(__ldg(tiles + t));
          if (r < T.rows()) {
            const int dl = T.dst_of(r);
            a.ew_out[(size_t)(T.a0 + T.d0 + dl) * KSTR + T.slot_of(r, dl)] = 1.f / (1.f + __expf(-(dot + sd[TM + r] + s_b2[0])));
          }
        }
        continue;
      }
      // z[zb] (TMEM columns / smem operand) was last read by GEMM2(t - 2)
      if (t >= 2) mbar_wait(bar + B_D2_FULL + (t - 2) % NB2, ((t - 2) / NB2) & 1);
      // two passes of 32 columns keep the packed output at 16 registers
      if (!SMB_DBG(a, 2))
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        uint32_t zp[16];
        ln_apply32(v, hp * 32, rstd, s_be + half * 64 + hp * 32, zp);
        if (ROLE == ROLE_V) {
          // z^T operand: K-major [row][k]; this thread owns k = 64 half .. 64 half + 63 of row r
          unsigned char* zrow = smem + P::o_z + zb * (TM * H * 2) + (r >> 3) * 2048 + (r & 7) * 16 + (half * 8 + hp * 4) * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(zrow + q * 128) = make_uint4(zp[4 * q], zp[4 * q + 1], zp[4 * q + 2], zp[4 * q + 3]);
        } else {
          tmem_st16(lane_addr + Z_COL + zb * 64 + half * 32 + hp * 16, zp);
        }
      }
      if (ROLE == ROLE_V) fence_async_smem();
      else wait_st();
      fence_before_sync();
      SMB_TRACE(3, t, gw == 0 && lane == 0);
      mbar_arrive(bar + B_Z_FULL + zb);
    }
  } else if (ROLE == ROLE_K && warp >= CV_WARP0 && warp < CV_WARP0 + 4) {
    reg_inc<R::REGS_CV>();
    // =====================================================================================
    // ROLE_K conversion group (every tile): stages the tile's queries as the block-diagonal B operand of the query fold, then
    // turns the fold's accumulator  M[m, 8 h + d]  (fp32, TMEM)  into the bf16 B operand of GEMM2,  [K = m][N = 16 d + h].
    // =====================================================================================
    const int tg = (warp - CV_WARP0) * 32 + lane;      // = TMEM lane = input channel m
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned char* s_qb = smem + P::o_qb;
    unsigned char* s_bm = smem + P::o_bm;
    const unsigned char* qimg = reinterpret_cast<const unsigned char*>(a.q);
    // chunk (h, d) = Q_d[8 h .. 8 h + 7] (bf16, 16 bytes): k-step j = h / 2, row n = (h & 1) * 8 + d, k-group h & 1
    auto stage_q = [&](int t, int row0, int nd) {      // row0 = first destination atom of the tile, nd destinations
      unsigned char* slot = s_qb + (t & 1) * P::QB_BYTES;
      for (int p = tg; p < nd * kHeads; p += E2_GRP_THREADS) {
        const int h = p / nd, d = p - h * nd, row = row0 + d;
        cp_async16(slot + (h >> 1) * 512 + (h & 1) * 384 + d * 16,
                   qimg + (size_t)(row >> 7) * kQChunkBlockBytes + (size_t)h * 2048 + (row & 127) * 16);
      }
    };
    // Q of tile t + 1 is staged and published while tile t is converted, so the fold of tile t + 1 only waits for this
    // group to have READ the fold of tile t
    int nd_cur = 0, nd_n1 = 0;
    if (nt > 0) { const int4 td = __ldg(tiles); nd_cur = (td.z >> 16) & 0xff; stage_q(0, td.x + ((td.z >> 8) & 0xff), nd_cur); }
    cp_async_commit();
    if (nt > 1) { const int4 td = __ldg(tiles + 1); nd_n1 = (td.z >> 16) & 0xff; stage_q(1, td.x + ((td.z >> 8) & 0xff), nd_n1); }
    cp_async_commit();
    cp_async_wait<1>();
    fence_async_smem();
    mbar_arrive(bar + B_QB_FULL + 0);
#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      const int nd = nd_cur;
      nd_cur = nd_n1;
      int4 td2 = make_int4(0, 0, 0, 0);
      if (t + 2 < nt) td2 = __ldg(tiles + t + 2);   // needed after the waits below
      cp_async_wait<0>();                       // Q of tile t + 1 (issued one iteration ago) has landed
      fence_async_smem();
      SMB_TRACE(15, t, tg == 0);
      if (t + 1 < nt) mbar_arrive(bar + B_QB_FULL + ((t + 1) & 1));
      nd_n1 = (td2.z >> 16) & 0xff;
      mbar_wait(bar + B_DM_FULL, t & 1);
      fence_after_sync();
      SMB_TRACE(16, t, tg == 0);
      // slot t & 1 was read by the fold of tile t, which has completed: stage tile t + 2 there
      if (t + 2 < nt) stage_q(t + 2, td2.x + ((td2.z >> 8) & 0xff), nd_n1);
      cp_async_commit();
      // four passes of 32 accumulator columns (heads 4 q .. 4 q + 3): half a 16-byte operand chunk per destination each
      // (not unrolled: one pass of 32 registers live at a time; hf / ps instead of q >> 1 / q & 1 -- nvcc 12.9 mis-derives
      //  the address of a barrier indexed by q >> 1 inside this loop)
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll 1
        for (int ps = 0; ps < 2; ++ps) {
          uint32_t v[32];
          tmem_ld32(lane_addr + R::DM_COL + hf * 64 + ps * 32, v);
          wait_ld();
          if (hf == 1 && ps == 1) { fence_before_sync(); mbar_arrive(bar + B_DM_FREE); }   // the accumulator may be overwritten by the next fold
          // the single M operand was last read by GEMM2 of tile t - 1
          if (hf == 0 && ps == 0 && t >= 1) mbar_wait(bar + B_D2_FULL + (t - 1) % NB2, ((t - 1) / NB2) & 1);
#pragma unroll
          for (int d = 0; d < NDMAX; ++d) {
            if (d < nd)   // operand column 16 d + 8 hf + 4 ps + i  <-  accumulator column 8 (8 hf + 4 ps + i) + d
              *reinterpret_cast<uint2*>(s_bm + (2 * d + hf) * 2048 + tg * 16 + ps * 8) =
                  make_uint2(pack_bf16(__uint_as_float(v[d]), __uint_as_float(v[8 + d])), pack_bf16(__uint_as_float(v[16 + d]), __uint_as_float(v[24 + d])));
          }
        }
      }
      fence_async_smem();
      SMB_TRACE(17, t, tg == 0);
      mbar_arrive(bar + B_BM_FULL);
    }
    cp_async_wait<0>();
  } else if (warp < TOP_WARP0 && (warp < E2_WARP0 || warp >= E2_WARP0 + E2_WARPS || ROLE == ROLE_GATE || ((warp - E2_WARP0) >> 2) >= R::NG)) {
    // ROLE_GATE ends in the LayerNorm role: no GEMM2, no role epilogue; every warpgroup below the top one that has no role
    // gives its registers back.
    reg_dec<REGS_IDLE>();
  } else if (warp < TOP_WARP0) {
    if (R::REGS_E2 > 72) reg_inc<R::REGS_E2>(); else reg_dec<R::REGS_E2>();
    const int e2w = warp - E2_WARP0;
    const int g = e2w >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;            // TMEM lane = tile row (ROLE_V: output channel)
    const int tg = (e2w & 3) * 32 + lane;    // thread index inside the group
    const int bar_id = BAR_E2 + g;
    constexpr int NG = R::NG > 0 ? R::NG : 1;
    const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
    unsigned char* es = smem + P::o_e2 + g * P::e2_bytes;
    float* s_log = reinterpret_cast<float*>(es + P::e_log);             // ROLE_K logits, ROLE_XV w
    float* s_ew = reinterpret_cast<float*>(es + P::e_ew);               // ROLE_K gate
    float4* s_rel = reinterpret_cast<float4*>(es + P::e_rel);           // ROLE_XV
    float* s_o = reinterpret_cast<float*>(es + P::e_o);                 // ROLE_XV
    float bn_s = 0.f, bn_q = 0.f;   // ROLE_XV: per-warp BatchNorm partial sums (lane & 15 = channel, lanes < 16)

    // what tile `t` needs from global memory, fetched after the group's previous tile: the tile's alpha block / shape
    // (cp.async) and this thread's gate value / relative position (registers)
    float pre_ew = 0.f, pre_x = 0.f, pre_y = 0.f, pre_z = 0.f;
    auto stage = [&](int t, const Tile& T) {
      unsigned char* slot = es + P::e_stage;
      const bool valid = r < T.rows();
      const int dl = valid ? T.dst_of(r) : 0, sl = T.slot_of(r, dl);
      if (ROLE == ROLE_K) {
        pre_ew = 0.f;
        if (valid) pre_ew = __ldg(a.ew_in + (size_t)(T.a0 + T.d0 + dl) * KSTR + sl);
      } else {
        const float* src = a.alpha_t + (size_t)(t_begin + t) * kAlphaTileFloats;
        float* dst = reinterpret_cast<float*>(slot);
#pragma unroll
        if (ROLE == ROLE_V) {   // whole-destination tiles: nd x SL floats per head; fixed trip count (the block's unused tail included)
#pragma unroll
          for (int p = tg; p < kAlphaTileFloats / 4; p += E2_GRP_THREADS) cp_async16(dst + p * 4, src + p * 4);
        } else {
          for (int p = tg; p < 4 * T.hstride(); p += E2_GRP_THREADS) cp_async16(dst + p * 4, src + p * 4);   // 16 heads x hstride floats
          if (tg < (kAlphaTileFloats - kAlphaSumOff) / 4) cp_async16(dst + kAlphaSumOff + tg * 4, src + kAlphaSumOff + tg * 4);   // sums | split statistics
        }
        // statistics of the other part of a split destination: the previous tile's last part | the next tile's first part
        // (read only when this tile's first / last part is incomplete, which implies that the neighbour exists)
        if (ROLE != ROLE_V && tg >= E2_GRP_THREADS - 16) {
          const int q = tg - (E2_GRP_THREADS - 16), w = q >> 3;
          if (w ? T.tail_split() : T.head_split())
            cp_async16(slot + P::o_nbst + q * 16, a.alpha_t + (size_t)(t_begin + t + (w ? 1 : -1)) * kAlphaTileFloats + kAlphaSplitOff + (w ? 0 : 32) + (q & 7) * 4);
        }
      }
      if (ROLE == ROLE_XV) {
        if (tg < 2 * kHeads * 3 / 4) cp_async16(slot + P::alpha_bytes + tg * 16, a.vn_shape + (size_t)T.mol * (2 * kHeads * 3) + tg * 4);
        pre_x = pre_y = pre_z = 0.f;
        if (valid) {
          const int i = T.d0 + dl;
          const int j = __ldg(a.nbr + (size_t)(T.a0 + i) * KSTR + sl);
          const float* xi = a.x + (size_t)(T.a0 + i) * 3;
          const float* xj = a.x + (size_t)(T.a0 + j) * 3;
          pre_x = __ldg(xi) - __ldg(xj); pre_y = __ldg(xi + 1) - __ldg(xj + 1); pre_z = __ldg(xi + 2) - __ldg(xj + 2);
        }
      }
    };

    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 td_cur = g < nt ? __ldg(tiles + g) : zero4;
    int4 td_nx = g + NG < nt ? __ldg(tiles + g + NG) : zero4;
    if (g < nt) stage(g, Tile(td_cur));
    cp_async_commit();
#pragma unroll 1
    for (int t = g; t < nt; t += NG) {
      const Tile T(td_cur);
      const float ew_r = pre_ew, relx = pre_x, rely = pre_y, relz = pre_z;
      td_cur = td_nx;
      if (t + 2 * NG < nt) td_nx = __ldg(tiles + t + 2 * NG);
      if (ROLE == ROLE_K && t + NG < nt) stage(t + NG, Tile(td_cur));   // registers only: the next tile's gate value
      if (ROLE != ROLE_K && t + NG < nt && tg < (kAlphaTileFloats * 4 + 127) / 128) {
        // the staging slot is busy until this tile is done: pull the group's next alpha block into L2 meanwhile, so that
        // the cp.async issued after the trailing barrier does not pay the DRAM latency
        // (only the lines that tile uses: 16 heads x hstride floats, then the sums | split statistics behind kAlphaSumOff)
        const float* nxt = a.alpha_t + (size_t)(t_begin + t + NG) * kAlphaTileFloats + tg * 32;
        if (ROLE == ROLE_V || tg * 32 < 16 * Tile(td_cur).hstride() || tg * 32 + 31 >= kAlphaSumOff) asm volatile("prefetch.L2 [%0];" :: "l"(nxt));
        // ... and the split statistics of that tile's neighbours (the 128-byte line of the previous tile's last part / the next
        // tile's first part), when it begins / ends inside a destination
        if (ROLE != ROLE_V && tg < 2) {
          const Tile N(td_cur);
          if (tg ? N.tail_split() : N.head_split())
            asm volatile("prefetch.L2 [%0];" :: "l"(a.alpha_t + (size_t)(t_begin + t + NG + (tg ? 1 : -1)) * kAlphaTileFloats + kAlphaSplitOff + (tg ? 0 : 32)));
        }
      }
      const int b3 = t % ND2, bb = t % NB2;   // TMEM buffer / barrier slot
      const uint32_t dcol = R::D2_COL + (uint32_t)b3 * R::D2_STRIDE;
      const int rows = T.rows();
      const bool valid = r < rows;
      const int dl = valid ? T.dst_of(r) : 0;
      const int hs = T.hstride();             // alpha floats per head
      const unsigned char* slot = es + P::e_stage;
      const float* s_al = reinterpret_cast<const float*>(slot);           // ROLE_V / ROLE_XV: the tile's alpha block
      const float* s_nb = reinterpret_cast<const float*>(slot + P::o_nbst);   // neighbours' split statistics [prev last | next first][16][2]
      // factor that turns the unnormalised alpha of an incomplete part into the destination's softmax: this part's (max, sum)
      // and the other part's, from the neighbouring tile.  w = 0: the tile's first part, 1: its last part.
      auto split_scale = [&](int w, int hd) {
        const float m_own = s_al[kAlphaSplitOff + w * 32 + hd * 2], s_own = s_al[kAlphaSplitOff + w * 32 + hd * 2 + 1];
        const float m_oth = s_nb[w * 32 + hd * 2], s_oth = s_nb[w * 32 + hd * 2 + 1];
        const float mg = fmaxf(m_own, m_oth);
        const float f_own = fast_ex2(m_own - mg);
        return f_own / fmaf(s_own, f_own, s_oth * fast_ex2(m_oth - mg));
      };
      // part table (one broadcast 16-byte read per use: keeps the per-part index arithmetic out of the unrolled loops)
      int4* s_ptab = reinterpret_cast<int4*>(es + P::e_ptab);
      if (ROLE != ROLE_V && tg < NDMAX) s_ptab[tg] = make_int4(T.first(tg), tg < T.nd ? T.count(tg) : 0, T.aoff(tg), 0);
      auto part_split = [&](int pd) { return s_ptab[pd].y < T.deg; };   // (deg = 0: never)
      const float* s_vs = reinterpret_cast<const float*>(slot + P::alpha_bytes);   // ROLE_XV: shape part of the VN maps [feat | dir][16][3]

      if (ROLE != ROLE_K) cp_async_wait<0>();   // this thread's share of tile t's staged data has landed
      if (ROLE == ROLE_XV) s_rel[r] = make_float4(relx, rely, relz, 0.f);
      mbar_wait(bar + B_D2_FULL + bb, (t / NB2) & 1);
      fence_after_sync();
      SMB_TRACE(5, t, tg == 0);
      if (ROLE != ROLE_K) named_sync(bar_id, E2_GRP_THREADS);      // staged alpha / shape / rel visible to the group

      if (ROLE == ROLE_K) {
        // logits of this row: the 16 accumulator columns of its destination (already scaled by log2(e) / sqrt(dh) through q;
        // the second Linear's bias shifts every logit of (i, head) by the same <Q_i, b2>: softmax-invariant, dropped)
        {
          float l[16];
#pragma unroll
          for (int hh = 0; hh < 16; ++hh) l[hh] = 0.f;
          const int r0 = qd * 32;
          if (r0 < rows) {
            const int d_lo = T.dst_of(r0), d_hi = T.dst_of(min(r0 + 31, rows - 1));
            for (int d = d_lo; d <= d_hi; ++d) {     // warp-uniform: the destinations this warp's 32 rows belong to
              uint32_t v[16];
              tmem_ld16(lane_addr + dcol + d * 16, v);
              wait_ld();
              if (d == dl) {
#pragma unroll
                for (int hh = 0; hh < 16; ++hh) l[hh] = __uint_as_float(v[hh]);
              }
            }
          }
          fence_before_sync();
          mbar_arrive(bar + B_E2_DONE + bb);      // the accumulator may be overwritten by GEMM1 of tile t + 2
#pragma unroll
          for (int hh = 0; hh < 16; ++hh) s_log[r * LS + hh] = l[hh];
          s_ew[r] = ew_r;
        }
        named_sync(bar_id, E2_GRP_THREADS);
        // per (destination, head): softmax over the destination's rows x gate.  Two threads share a pair; each owns 16
        // consecutive slots, holds them in registers and writes its part of alpha[head][destination][slot] with 16-byte stores
        {
          const int part = tg & 1;
          float* at = a.alpha_t + (size_t)(t_begin + t) * kAlphaTileFloats;
          for (int pair = tg >> 1; pair < T.nd * kHeads; pair += E2_GRP_THREADS / 2) {
            const int pd = pair >> 4, hd = pair & 15;
            const int4 pt = s_ptab[pd];
            const int first = pt.x, cnt = pt.y;
            const bool split = cnt < T.deg;        // the destination continues in a neighbouring tile: alpha stays unnormalised
            const float* col = s_log + (first + 16 * part) * LS + hd;
            const float* gw = s_ew + first + 16 * part;
            const int nq = cnt - 16 * part;        // valid slots of this half (may be <= 0)
            float lv[16];
            float mx = -INFINITY;
#pragma unroll
            for (int qq = 0; qq < 16; ++qq) {
              lv[qq] = qq < nq ? col[qq * LS] : -INFINITY;
              mx = fmaxf(mx, lv[qq]);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            float se = 0.f;
#pragma unroll
            for (int qq = 0; qq < 16; ++qq) { lv[qq] = fast_ex2(lv[qq] - mx); se += lv[qq]; }
            se += __shfl_xor_sync(0xffffffffu, se, 1);
            const float inv = split ? 1.f : 1.f / se;
            float asum = 0.f;
#pragma unroll
            for (int qq = 0; qq < 16; ++qq) { lv[qq] = qq < nq ? lv[qq] * inv * gw[qq] : 0.f; asum += lv[qq]; }
            asum += __shfl_xor_sync(0xffffffffu, asum, 1);
            float4* dst = reinterpret_cast<float4*>(at + hd * hs + pt.z + 16 * part);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              if (16 * part + 4 * q4 < cnt) dst[q4] = make_float4(lv[4 * q4], lv[4 * q4 + 1], lv[4 * q4 + 2], lv[4 * q4 + 3]);
            if (part == 0) {
              at[kAlphaSumOff + hd * NDMAX + pd] = asum;
              if (split) {
                float* sp = at + kAlphaSplitOff + ((pd == 0 && T.head_split()) ? 0 : 32) + hd * 2;
                sp[0] = mx; sp[1] = se;
              }
            }
          }
          // the rows a split destination's two parts accumulate into (ROLE_V: agg, ROLE_XV: the o sums in the vn row) start at zero;
          // the tile that holds the FIRST part clears them
          if (T.tail_split() && a.zero_ptr && tg < a.zero_len)
            a.zero_ptr[(size_t)(T.a0 + T.d0 + T.nd - 1) * a.zero_stride + a.zero_off + tg] = 0.f;
        }
      } else if (ROLE == ROLE_V) {
        // thread = output channel c (TMEM lane), columns = edge rows
        const int c = r, hq = c >> 3;
        const float b2c = s_b2[c];
        if (T.deg == 0) {   // single-atom molecule: empty neighbour sum
          a.agg[(size_t)(T.a0 + T.d0) * H + c] = 0.f;
        } else {
          // ROLE_V runs on the whole-destination tile list (smb_api.cu): part pd = destination pd, deg rows, no split parts.
          // A destination's deg <= 31 columns are read as ONE 32-column load (16 when deg <= 16) -- whatever follows them, the
          // next destination's columns or the spare TMEM columns behind the ring, is multiplied by the zero padding of its
          // alpha slots (SL = deg rounded up to 4) or not used at all -- and two destinations share one tcgen05.wait::ld:
          // the load latency is paid nd / 2 times per tile.  Groups of four slots: one 16-byte alpha load, two packed FMAs.
          const int SL = (T.deg + 3) & ~3;
          const int n4 = SL >> 2;
          auto reduce32 = [&](int pd, const uint32_t (&v)[32]) {
            const float* al = s_al + hq * hs + pd * SL;
            uint64_t acc2 = 0ull;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < n4) {
                const float4 w = *reinterpret_cast<const float4*>(al + 4 * q);
                acc2 = fma2(pk2u(v[4 * q], v[4 * q + 1]), pk2(w.x, w.y), acc2);
                acc2 = fma2(pk2u(v[4 * q + 2], v[4 * q + 3]), pk2(w.z, w.w), acc2);
              }
            a.agg[(size_t)(T.a0 + T.d0 + pd) * H + c] = fmaf(b2c, s_al[kAlphaSumOff + hq * NDMAX + pd], sum2(acc2));
          };
          for (int pd = 0; pd < T.nd; pd += 2) {
            uint32_t va[32], vb[32];
            const bool two = pd + 1 < T.nd;
            tmem_ld32(lane_addr + dcol + pd * T.deg, va);
            if (two) tmem_ld32(lane_addr + dcol + (pd + 1) * T.deg, vb);
            wait_ld();
            reduce32(pd, va);
            if (two) reduce32(pd + 1, vb);
          }
        }
        fence_before_sync();
        mbar_arrive(bar + B_E2_DONE + bb);
      } else {   // ROLE_XV
        {
          uint32_t v[16];
          tmem_ld16(lane_addr + dcol, v);
          wait_ld();
          fence_before_sync();
          mbar_arrive(bar + B_E2_DONE + bb);
          if (valid) {
            const int4 pt = s_ptab[dl];
            const float* al = s_al + pt.z + (r - pt.x);     // + head * hs
#pragma unroll
            for (int hh = 0; hh < 16; ++hh) s_log[r * LS + hh] = al[hh * hs] * (__uint_as_float(v[hh]) + s_b2[hh]);
          }
        }
        named_sync(bar_id, E2_GRP_THREADS);
        // o_i^a = sum_j alpha e_w w (x_i - x_j)
        for (int p = tg; p < T.nd * kHeads; p += E2_GRP_THREADS) {
          const int pd = p >> 4, hd = p & 15;
          const int r0 = s_ptab[pd].x, cnt = s_ptab[pd].y;
          float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll 4
          for (int q = 0; q < cnt; ++q) {
            const float w = s_log[(r0 + q) * LS + hd];
            const float4 rl = s_rel[r0 + q];
            ox = fmaf(w, rl.x, ox); oy = fmaf(w, rl.y, oy); oz = fmaf(w, rl.z, oz);
          }
          if (cnt < T.deg) {   // incomplete part: to the destination's softmax (the other part is in a neighbouring tile)
            const float sc = split_scale((pd == 0 && T.head_split()) ? 0 : 1, hd);
            ox *= sc; oy *= sc; oz *= sc;
          }
          float* o = s_o + (pd * kHeads + hd) * 4;
          o[0] = ox; o[1] = oy; o[2] = oz;
        }
        named_sync(bar_id, E2_GRP_THREADS);
        // VN linear maps (shape_vn_layers.py:100,105): lanes 0..15 map_to_feat channel, 16..31 map_to_dir channel
        for (int pd = e2w & 3; pd < T.nd; pd += 4) {
          const int ch = lane & 15, which = lane >> 4;
          const float* w = s_vnw + (which * kHeads + ch) * kVnStride;
          const float* so = s_o + pd * kHeads * 4;
          if (part_split(pd)) {
            // a destination split between two tiles: both parts add their o sums into the vn row (cleared by ROLE_K; two
            // addends: deterministic); its VN maps and BatchNorm terms follow in xv_split_finish_kernel
            float* acc = a.vn + (size_t)(T.a0 + T.d0 + pd) * kVnRow + 3;
            if (lane < kHeads) { atomicAdd(acc + lane * 3, so[lane * 4]); atomicAdd(acc + lane * 3 + 1, so[lane * 4 + 1]); atomicAdd(acc + lane * 3 + 2, so[lane * 4 + 2]); }
            continue;
          }
          const float* xp = a.x + (size_t)(T.a0 + T.d0 + pd) * 3;
          const float xi = __ldg(xp), yi = __ldg(xp + 1), zi = __ldg(xp + 2);
          float vx = w[0] * xi, vy = w[0] * yi, vz = w[0] * zi;
#pragma unroll
          for (int cc = 0; cc < kHeads; ++cc) {
            const float wc = w[1 + cc];
            vx = fmaf(wc, so[cc * 4], vx); vy = fmaf(wc, so[cc * 4 + 1], vy); vz = fmaf(wc, so[cc * 4 + 2], vz);
          }
          {   // the shape-embedding channels of z = [x | o | shape_emb] do not depend on the step: prep_kernel contracted them
            const float* vs = s_vs + (which * kHeads + ch) * 3;
            vx += vs[0]; vy += vs[1]; vz += vs[2];
          }
          float* row = a.vn + (size_t)(T.a0 + T.d0 + pd) * kVnRow;
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
      SMB_TRACE(6, t, tg == 0);
      named_sync(bar_id, E2_GRP_THREADS);   // scratch and the staging slot are reused by the group's next tile
      if (ROLE != ROLE_K) {
        if (t + NG < nt) stage(t + NG, Tile(td_cur));
        cp_async_commit();
      }
    }   // tiles
    cp_async_wait<0>();
    if (ROLE == ROLE_XV && lane < 16) {
      float* part = reinterpret_cast<float*>(smem + P::o_bn) + e2w * 32;   // summed per CTA after the final barrier
      part[lane] = bn_s;
      part[16 + lane] = bn_q;
    }
  } else if (warp == MMA_WARP) {
    // =====================================================================================
    // GEMM1 issuer (lane 0 issues the MMAs): three barrier waits, six MMAs, ONE commit per tile -- every one of these costs
    // the issuing thread 60..180 cycles, and this serial loop bounds the pipeline's period
    // =====================================================================================
    constexpr uint32_t IDESC1 = idesc_bf16(H, true);
    const uint32_t a1_base = smem_u32(s_a1), ab_base = smem_u32(s_ab), w1r_base = smem_u32(s_w1r);
    int n_cur = nt > 0 ? (__ldg(tiles).z & 0xff) : 0, n_nx = nt > 1 ? (__ldg(tiles + 1).z & 0xff) : 0;
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 td_ld = zero4;   // ROLE_XV (no spare warp for the loader): descriptor of tile t + 2, loaded one iteration early
    if (ROLE == ROLE_XV) {
      if (lane == 0) { load_ab(0, nt > 0 ? __ldg(tiles) : zero4); load_ab(1, nt > 1 ? __ldg(tiles + 1) : zero4); }
      if (nt > 2) td_ld = __ldg(tiles + 2);
    }
#pragma unroll 1
    for (int t = 0; t < nt; ++t) {
      const int n_mol = n_cur;
      n_cur = n_nx;
      if (t + 2 < nt) n_nx = __ldg(tiles + t + 2).z & 0xff;
      const int4 td_l = td_ld;
      if (ROLE == ROLE_XV && t + 3 < nt) td_ld = __ldg(tiles + t + 3);
      mbar_wait<32>(bar + B_A1_FULL + (t & 1), (t >> 1) & 1);
      SMB_TRACE(7, t, lane == 0);
      if (R::SEP) { if (t >= ND1) mbar_wait<32>(bar + B_D1_FREE + t % ND1, (t / ND1 - 1) & 1); }   // LayerNorm(t - ND1) has read D1[t % ND1]
      else if (t >= ND1) mbar_wait<32>(bar + B_E2_DONE + (t - ND1) % NB2, ((t - ND1) / NB2) & 1);   // tile t - ND1 left D[t % ND1]
      SMB_TRACE(12, t, lane == 0);
      if (ROLE != ROLE_GATE) mbar_wait<32>(bar + B_AB_FULL + t % 3, (t / 3) & 1);
      fence_after_sync();
      SMB_TRACE(1, t, lane == 0);
      if (lane == 0) {
        const uint32_t d = tmem + (uint32_t)(t % ND1) * 128u;
        const uint32_t a1 = a1_base + (t & 1) * A1_BYTES;
        const uint32_t ab = ab_base + (t % 3) * AB_BYTES;
        const uint32_t sbo_ab = (uint32_t)n_mol * 16u;   // n * 16: column-group stride of the projection tiles
        mma_ss(d, smem_desc(a1, 128, A1_SBO), smem_desc(w1r_base, 128, 512), IDESC1, 0);
        mma_ss(d, smem_desc(a1 + 256, 128, A1_SBO), smem_desc(w1r_base + 256, 128, 512), IDESC1, 1);
        if (ROLE != ROLE_GATE) {
#pragma unroll
          for (int ks = 2; ks < K1 / 16; ++ks)
            mma_ss(d, smem_desc(a1 + ks * 256, 128, A1_SBO), smem_desc(ab + (ks >> 2) * (G * H * 2) + (ks & 1) * 256, 128, sbo_ab), IDESC1, 1);
        }
        SMB_TRACE(8, t, true);
        mma_commit(bar + B_D1_FULL + t % NB1);
        SMB_TRACE(9, t, true);
      }
      __syncwarp();
      if (ROLE == ROLE_XV) {   // projection slot (t + 2) % 3 was read by GEMM1(t - 1)
        if (t >= 1) mbar_wait<32>(bar + B_D1_FULL + (t - 1) % NB1, ((t - 1) / NB1) & 1);
        if (lane == 0) load_ab(t + 2, td_l);
      }
    }
  } else if (ROLE != ROLE_XV && ROLE != ROLE_GATE && warp == TOP_WARP0) {
    // =====================================================================================
    // projection-tile loader (ROLE_XV has no spare warp: its GEMM2 issuer runs these steps)
    // =====================================================================================
    if (lane == 0) {
      const int4 zero4 = make_int4(0, 0, 0, 0);
      load_ab(0, nt > 0 ? __ldg(tiles) : zero4);
      load_ab(1, nt > 1 ? __ldg(tiles + 1) : zero4);
      int4 td_nx = nt > 2 ? __ldg(tiles + 2) : zero4;
#pragma unroll 1
      for (int t = 0; t + 2 < nt; ++t) {
        const int4 td_ld = td_nx;
        if (t + 3 < nt) td_nx = __ldg(tiles + t + 3);
        // projection slot (t + 2) % 3 was read by GEMM1(t - 1)
        if (t >= 1) mbar_wait<32>(bar + B_D1_FULL + (t - 1) % NB1, ((t - 1) / NB1) & 1);
        SMB_TRACE(10, t, true);
        load_ab(t + 2, td_ld);
        SMB_TRACE(11, t, true);
      }
    }
  } else if (ROLE == ROLE_K && warp == F_WARP) {
    // =====================================================================================
    // ROLE_K query-fold issuer (its own warp: a tcgen05.mma costs the issuing thread ~50-60 cycles whatever its N, and the
    // fold of tile u + 1 must not queue behind -- or ahead of -- GEMM2 of tile u in one instruction stream)
    //   M[m, 8 h + d] = sum_{c in head h} W2[c, m] Q_d[c]:  k-step j covers the channels of heads 2 j, 2 j + 1, whose 16
    //   output columns (hh, d) start at 16 j;  A = W2^T [m][c] (K-major), B = block-diagonal Q chunk rows [16][16] (K-major)
    // =====================================================================================
    constexpr uint32_t IDESC_F = idesc_bf16(16, false);
    const uint32_t w2_base = smem_u32(s_w2), qb_base = smem_u32(smem + P::o_qb);
#pragma unroll 1
    for (int u = 0; u < nt; ++u) {
      mbar_wait<32>(bar + B_QB_FULL + (u & 1), (u >> 1) & 1);
      if (u >= 1) mbar_wait<32>(bar + B_DM_FREE, (u - 1) & 1);       // the conversion group has read the fold of tile u - 1
      fence_after_sync();
      SMB_TRACE(18, u, lane == 0);
      if (lane == 0) {
        const uint32_t qb = qb_base + (u & 1) * P::QB_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          mma_ss(tmem + R::DM_COL + 16 * j, smem_desc(w2_base + j * 256, 128, 2048), smem_desc(qb + j * 512, 128, 256), IDESC_F, 0);
        mma_commit(bar + B_DM_FULL);
      }
      __syncwarp();
    }
  } else if (warp == G2_WARP) {
    // =====================================================================================
    // GEMM2 issuer
    // =====================================================================================
    constexpr uint32_t IDESC2 = idesc_bf16(ROLE == ROLE_XV ? kHeads : H, false);
    const uint32_t w2_base = smem_u32(s_w2);
    const uint32_t bm_base = smem_u32(smem + P::o_bm);
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 td_cur = nt > 0 ? __ldg(tiles) : zero4, td_nx = nt > 1 ? __ldg(tiles + 1) : zero4;
#pragma unroll 1
    for (int u = 0; u < (ROLE == ROLE_GATE ? 0 : nt); ++u) {
      const int zb = u & 1;
      const int nd = (td_cur.z >> 16) & 0xff;
      td_cur = td_nx;
      if (u + 2 < nt) td_nx = __ldg(tiles + u + 2);
      mbar_wait<32>(bar + B_Z_FULL + zb, (u >> 1) & 1);
      if (R::SEP && u >= ND2) mbar_wait<32>(bar + B_E2_DONE + (u - ND2) % NB2, ((u - ND2) / NB2) & 1);   // epilogue(u - ND2) has read D2[u % ND2]
      if (ROLE == ROLE_K) mbar_wait<32>(bar + B_BM_FULL, u & 1);
      fence_after_sync();
      SMB_TRACE(4, u, lane == 0);
      if (lane == 0) {
        const uint32_t d = tmem + R::D2_COL + (uint32_t)(u % ND2) * R::D2_STRIDE;
        if (ROLE == ROLE_V) {
          const uint32_t zt = smem_u32(smem + P::o_z + zb * (TM * H * 2));
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks)
            mma_ss(d, smem_desc(w2_base + ks * 256, 128, 2048), smem_desc(zt + ks * 256, 128, 2048), IDESC2, ks > 0);
        } else if (ROLE == ROLE_K) {
          // logits[e, 16 d + h] = z[e, :] . M[:, 16 d + h]:  N = 16 nd, B = M operand [K = m][N] MN-major
          const uint32_t idk = idesc_bf16(16 * nd, true);
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks)
            mma_ts(d, tmem + Z_COL + zb * 64 + ks * 8, smem_desc(bm_base + ks * 256, 128, 2048), idk, ks > 0);
        } else {
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks)
            mma_ts(d, tmem + Z_COL + zb * 64 + ks * 8, smem_desc(w2_base + ks * 256, 128, 2048), IDESC2, ks > 0);
        }
        mma_commit(bar + B_D2_FULL + u % NB2);
      }
      __syncwarp();
    }
  }

  fence_before_sync();
  __syncthreads();
  if (ROLE == ROLE_XV && warp == 1) {   // one BatchNorm partial row per CTA, fixed summation order (deterministic)
    const float* part = reinterpret_cast<const float*>(smem + P::o_bn);
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 4 * R::NG; ++w) acc += part[w * 32 + lane];
    a.bn_partial[(size_t)blockIdx.x * 32 + lane] = acc;
  }
  if (warp == 0) tmem_free<TMEM_COLS>(tmem);
}

// ---- static tile list -------------------------------------------------------------------------------
// Tiles of a molecule of n atoms (deg = min(k, n - 1) neighbour slots per destination, E = n deg edge slots).
//   split = false: whole destinations, min(128 / deg, NDMAX) per tile (27 atoms: 4 x 26 = 104 of 128 rows, 7 tiles)
//   split = true : consecutive runs of the E slots, all about E / ceil(E / 128) rows long, that may begin and end inside a
//                  destination (27 atoms: 6 x 117 rows), never more than NDMAX parts
struct TileWalk {
  int deg, E, target, e;
  bool split;
  __device__ __forceinline__ TileWalk(int n, int k, bool split_) : deg(min(k, n - 1)), e(0), split(split_) {
    E = n * deg;
    const int per = deg > 0 ? min(TM / deg, NDMAX) * deg : 0;
    target = split ? (E + (E + TM - 1) / TM - 1) / max((E + TM - 1) / TM, 1) : per;
  }
  __device__ __forceinline__ bool done() const { return e >= E; }
  // next tile: first slot e0, rows, parts
  __device__ __forceinline__ void next(int& e0, int& rows, int& d0, int& nd) {
    e0 = e;
    d0 = e / deg;
    rows = min(target, E - e);
    if ((e + rows - 1) / deg - d0 + 1 > NDMAX) rows = (d0 + NDMAX) * deg - e;
    nd = (e + rows - 1) / deg - d0 + 1;
    e += rows;
  }
};
__device__ __forceinline__ int tiles_of_walk(int n, int k, bool split) {
  TileWalk w(n, k, split);
  int cnt = 0, e0, rows, d0, nd;
  while (!w.done()) { w.next(e0, rows, d0, nd); ++cnt; }
  return cnt;
}
// a molecule's tiles split destinations only where that saves a tile (with the 8-part cap an equal run can even need one more)
__device__ __forceinline__ bool split_pays(int n, int k) { return tiles_of_walk(n, k, true) < tiles_of_walk(n, k, false); }
__device__ __forceinline__ int tiles_of(int n, int k, bool split) {
  if (n <= 0) return 0;
  if (min(k, n - 1) == 0) return 1;
  return tiles_of_walk(n, k, split && split_pays(n, k));
}

__global__ void __launch_bounds__(1024) build_tiles_kernel(const int* __restrict__ mol_ptr, int n_mols, int k, int split, int4* __restrict__ tiles,
                                                           int* __restrict__ n_tiles) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_thread = (n_mols + 1023) / 1024;
  const int m0 = min(n_mols, tid * per_thread), m1 = min(n_mols, m0 + per_thread);
  int cnt = 0;
  for (int m = m0; m < m1; ++m) cnt += tiles_of(mol_ptr[m + 1] - mol_ptr[m], k, split != 0);
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += up;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += up;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  int off = inc - cnt + (warp > 0 ? s_warp[warp - 1] : 0);
  if (tid == 1023) *n_tiles = s_warp[31];
  for (int m = m0; m < m1; ++m) {
    const int a0 = mol_ptr[m], n = mol_ptr[m + 1] - a0;
    if (n <= 0) continue;
    const int deg = min(k, n - 1);
    if (deg == 0) { tiles[off++] = make_int4(a0, m, n | (1 << 16), 0); continue; }   // single atom: one empty tile
    const int recip = 65536 / deg + 1;
    TileWalk w(n, k, split != 0 && split_pays(n, k));
    while (!w.done()) {
      int e0, rows, d0, nd;
      w.next(e0, rows, d0, nd);
      tiles[off++] = make_int4(a0, m | ((e0 - d0 * deg) << 24), n | (d0 << 8) | (nd << 16) | (deg << 24), recip | (rows << 20));
    }
  }
}

// ---- destinations split between two tiles (ROLE_XV): VN linear maps + BatchNorm terms once both parts' o sums are in the vn row
__global__ void __launch_bounds__(256) xv_split_finish_kernel(EdgeArgs a, int row0) {
  __shared__ float s_w[2 * kHeads * kVnStride];
  __shared__ float s_o[8][kHeads * 3];
  __shared__ float s_bn[8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int p = tid; p < kHeads * kVnStride; p += 256) { s_w[p] = a.vn_feat[p]; s_w[kHeads * kVnStride + p] = a.vn_dir[p]; }
  __syncthreads();
  const int n_tiles = *a.n_tiles;
  const int ch = lane & 15, which = lane >> 4;
  float bn_s = 0.f, bn_q = 0.f;
  for (int t = blockIdx.x * 8 + warp; t < n_tiles; t += gridDim.x * 8) {   // fixed assignment: deterministic partial sums
    const Tile T(__ldg(a.tiles + t));
    if (T.deg == 0 || !T.tail_split()) continue;
    const int atom = T.a0 + T.d0 + T.nd - 1;
    float* row = a.vn + (size_t)atom * kVnRow;
    float* so = s_o[warp];
    so[lane] = row[3 + lane];
    if (lane < 16) so[32 + lane] = row[35 + lane];
    __syncwarp();
    const float* w = s_w + (which * kHeads + ch) * kVnStride;
    const float* xp = a.x + (size_t)atom * 3;
    const float xi = __ldg(xp), yi = __ldg(xp + 1), zi = __ldg(xp + 2);
    float vx = w[0] * xi, vy = w[0] * yi, vz = w[0] * zi;
#pragma unroll
    for (int cc = 0; cc < kHeads; ++cc) {
      const float wc = w[1 + cc];
      vx = fmaf(wc, so[cc * 3], vx); vy = fmaf(wc, so[cc * 3 + 1], vy); vz = fmaf(wc, so[cc * 3 + 2], vz);
    }
    {
      const float* vs = a.vn_shape + (size_t)T.mol * (2 * kHeads * 3) + (which * kHeads + ch) * 3;
      vx += __ldg(vs); vy += __ldg(vs + 1); vz += __ldg(vs + 2);
    }
    float sm = 0.f;
    if (lane < 3) {
#pragma unroll
      for (int cc = 0; cc < kHeads; ++cc) sm += so[cc * 3 + lane];
    }
    __syncwarp();
    row[3 + which * 48 + ch * 3] = vx; row[4 + which * 48 + ch * 3] = vy; row[5 + which * 48 + ch * 3] = vz;
    if (lane < 3) row[lane] = sm * (1.f / kHeads);
    if (which == 0) {
      const float nu = sqrtf(vx * vx + vy * vy + vz * vz) + 1e-6f;
      bn_s += nu; bn_q = fmaf(nu, nu, bn_q);
    }
    __syncwarp();
  }
  // lanes < 16 hold channel sums: row layout [sum | sum of squares] like the edge kernel's partial rows
  const float qv = __shfl_sync(0xffffffffu, bn_q, lane & 15);
  s_bn[warp][lane] = lane < 16 ? bn_s : qv;
  __syncthreads();
  if (warp == 0) {
    float acc = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) acc += s_bn[w8][lane];
    a.bn_partial[(size_t)(row0 + blockIdx.x) * 32 + lane] = acc;
  }
}

template <int ROLE>
int launch_ws(const EdgeArgs& a_in, int* bn_rows_out, cudaStream_t st) {
  EdgeArgs a = a_in;
#ifdef SMB_DEBUG
  static const int dbg = getenv("SMB_WS_DBG") ? atoi(getenv("SMB_WS_DBG")) : 0;
  a.dbg = dbg;
#else
  a.dbg = 0;
#endif
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem(edge_ws_kernel<ROLE>, Plan<ROLE>::total, configured)) return rc;
  int grid = device_sm_count();
  if (grid > kEdgeMaxCtas) grid = kEdgeMaxCtas;
  if (bn_rows_out) *bn_rows_out = grid;
  edge_ws_kernel<ROLE><<<grid, THREADS, Plan<ROLE>::total, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

#ifdef SMB_DEBUG
int debug_ws_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * TRACE_EVENTS * TRACE_TILES);
}
#endif

bool edge_ws_supported(const smb_model_dims& d, int n_max) {
#ifdef SMB_DEBUG
  static const bool off = getenv("SMB_EDGE_LEGACY") != nullptr;   // debugging aid: mma.sync kernels in bf16 mode
  if (off) return false;
#endif
  return d.precision == SMB_PREC_BF16 && d.hidden == H && n_max >= 1 && n_max <= G && d.k >= 1;
}

int launch_build_tiles(const int* mol_ptr, int n_mols, int k, bool split, int4* tiles, int* n_tiles, cudaStream_t st) {
#ifdef SMB_DEBUG
  static const int allow = getenv("SMB_TILE_SPLIT") ? atoi(getenv("SMB_TILE_SPLIT")) : 1;   // 0: whole destinations everywhere (A/B runs)
#else
  const int allow = 1;
#endif
  build_tiles_kernel<<<1, 1024, 0, st>>>(mol_ptr, n_mols, k, split && allow ? 1 : 0, tiles, n_tiles);
  return (int)cudaGetLastError();
}

int launch_xv_split_finish(const EdgeArgs& a, int bn_rows_in, int* bn_rows_out, cudaStream_t st) {
  // one warp per tile with a dependent load chain (descriptor -> vn row -> stores): many small CTAs hide its latency.  Every CTA
  // writes one BatchNorm partial row (fixed tile assignment: deterministic), the workspace holds kEdgeMaxCtas * kEdgeWarps rows.
  int grid = device_sm_count() * 8;
  if (grid > kEdgeMaxCtas * (kEdgeWarps - 1)) grid = kEdgeMaxCtas * (kEdgeWarps - 1);
  xv_split_finish_kernel<<<grid, 256, 0, st>>>(a, bn_rows_in);
  if (bn_rows_out) *bn_rows_out = bn_rows_in + grid;
  return (int)cudaGetLastError();
}

int launch_edge_ws(int role, const EdgeArgs& a, int* bn_rows_out, cudaStream_t st) {
  switch (role) {
    case ROLE_GATE: return launch_ws<ROLE_GATE>(a, bn_rows_out, st);
    case ROLE_K: return launch_ws<ROLE_K>(a, bn_rows_out, st);
    case ROLE_V: return launch_ws<ROLE_V>(a, bn_rows_out, st);
    case ROLE_XV: return launch_ws<ROLE_XV>(a, bn_rows_out, st);
    default: set_error_msg("launch_edge_ws: unsupported role"); return SMB_E_UNSUPPORTED;
  }
}

}  // namespace smb
