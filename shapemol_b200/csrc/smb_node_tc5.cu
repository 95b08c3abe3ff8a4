// Node-level projections on tcgen05 / TMEM (plain-bf16 contraction mode): the per-atom halves of the edge MLPs'
// first Linear and the query MLP of one attention block, for the warp-specialised edge pipeline.
//
//   X  = [h (128) | inv[mol] (32) | 1 | 1 | 0 ...]                       [128 atoms x 176]   bf16, K-major, smem
//   Y1 = X W1^T : 640 columns in 5 chunks of 128: the query MLP's hidden layer first, then the four pass-through
//        blocks A_k | B_k | A_v | B_v  (dst / src parts of hk/hv or xk/xv, models/uni_transformer.py:61-71,133-140)
//        bias = two extra K columns (bf16 hi | lo) against the constant-one columns of X
//   pass-through blocks -> bf16, stored in the per-molecule MN-major operand image that the edge pipeline bulk-copies
//   hidden chunk        -> LayerNorm (folded, see smb_host.cu) -> ReLU -> z (bf16, TMEM) -> GEMM2 (A from TMEM) + b2 -> q
//
// One persistent CTA per SM, 128-atom tiles.  Roles: warp 0 streams the 45 KB weight chunks L2 -> smem with bulk
// copies (2-slot ring), warp 1 issues the MMAs (accumulators in a 3-deep TMEM ring), warps 2-5 convert the fp32
// activations of the next tile into the operand layout (2 slots), warps 6-13 are the epilogue (thread = row x
// column half).  Bound: HBM (0.5 KB read + 1.5 KB written per atom); the tensor pipe needs ~10 us per launch.
#include "smb_common.cuh"
#include "smb_kernels.h"
#include "smb_tc.cuh"

#include <cstdlib>

namespace smb {

namespace {

using namespace tc;

constexpr int H = 128, TM = 128;
constexpr int W2_BYTES = H * H * 2;
constexpr int W_WARP = 0, MMA_WARP = 1, X_WARP0 = 2, X_WARPS = 4, E_WARP0 = 6, E_WARPS = 8, WARPS = 14, THREADS = WARPS * 32;
constexpr int E_THREADS = E_WARPS * 32, X_THREADS = X_WARPS * 32;
constexpr int BAR_LN = 1;
constexpr uint32_t Z_COL = 384;

enum { B_X_FULL = 0, B_X_FREE = 2, B_W_FULL = 4, B_W_FREE = 6, B_D_FULL = 8, B_D_FREE = 11, B_Z_FULL = 14, N_BARS = 15 };

// MODE 0 (pre): X = [h | inv | 1 1 0..] (K = 176), five streamed weight chunks, two X slots; accumulator productions per
//               tile: hidden, A_k, B_k, A_v, B_v, GEMM2 (q)
// MODE 1 (out): X = [agg | h | 1 1 0..] (K = 272), node_output MLP (uni_transformer.py:82,87-88): one resident weight
//               block, one X slot; productions per tile: hidden, GEMM2 (+ b2 + residual h -> h')
template <int MODE>
struct Cfg {
  static constexpr int KX = MODE == 0 ? kNodeKx : kNodeOutKx;
  static constexpr int X_SBO = (KX / 8) * 128;
  static constexpr int X_BYTES = TM * KX * 2;
  static constexpr int WCH_BYTES = 128 * KX * 2;
  static constexpr int N_CH = MODE == 0 ? 5 : 1;
  static constexpr int NXS = MODE == 0 ? 2 : 1, NWS = MODE == 0 ? 2 : 1;
  static constexpr int STEPS = MODE == 0 ? 6 : 2, G2_STEP = MODE == 0 ? 5 : 1;   // GEMM2 last: the MMA warp never waits for the LayerNorm
  static constexpr int o_bar = 0;
  static constexpr int o_tmem = 128;
  static constexpr int o_beta = 256;                  // beta / |gamma| of the LayerNorm
  static constexpr int o_b2 = o_beta + 512;
  static constexpr int o_stat = o_b2 + 512;           // float[2 tiles][2 halves][128 rows]
  static constexpr int o_w2 = o_stat + 2048;          // 3328
  static constexpr int o_x = o_w2 + W2_BYTES;
  static constexpr int o_w = o_x + NXS * X_BYTES;
  static constexpr int SMEM_TOTAL = o_w + NWS * WCH_BYTES;
  static_assert(o_w2 % 128 == 0 && SMEM_TOTAL <= 227 * 1024, "shared memory plan");
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) node_tc5_kernel(NodeArgs a) {
  using C = Cfg<MODE>;
  constexpr int KX = C::KX, X_SBO = C::X_SBO, X_BYTES = C::X_BYTES, WCH_BYTES = C::WCH_BYTES, N_CH = C::N_CH;
  constexpr int NXS = C::NXS, NWS = C::NWS, STEPS = C::STEPS, G2_STEP = C::G2_STEP;
  constexpr int o_bar = C::o_bar, o_tmem = C::o_tmem, o_beta = C::o_beta, o_b2 = C::o_b2, o_stat = C::o_stat, o_w2 = C::o_w2, o_x = C::o_x, o_w = C::o_w;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + o_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + o_tmem);
  float* s_beta = reinterpret_cast<float*>(smem + o_beta);
  float* s_b2 = reinterpret_cast<float*>(smem + o_b2);
  float* s_stat = reinterpret_cast<float*>(smem + o_stat);
  unsigned char* s_w2 = smem + o_w2;
  unsigned char* s_x = smem + o_x;
  unsigned char* s_w = smem + o_w;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = (a.n_atoms + TM - 1) / TM;
  const int nt = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles b, b + grid, ...

  // ---- once per CTA ----
  {
    const uint4* s2 = reinterpret_cast<const uint4*>(a.w2_t);
    uint4* d2 = reinterpret_cast<uint4*>(s_w2);
    for (int p = tid; p < W2_BYTES / 16; p += THREADS) d2[p] = s2[p];
    if (tid < H) { s_beta[tid] = a.beta_t[tid]; s_b2[tid] = a.b2[tid]; }
    // operand slots start at zero; the constant-one bias columns (k = 160, 161) are written once
    uint4* z0 = reinterpret_cast<uint4*>(s_x);
    for (int p = tid; p < NXS * X_BYTES / 16; p += THREADS) z0[p] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int p = tid; p < NXS * TM; p += THREADS) {
    const int slot = p >> 7, r = p & 127;
    *reinterpret_cast<uint32_t*>(s_x + slot * X_BYTES + (r >> 3) * X_SBO + ((KX - 16) / 8) * 128 + (r & 7) * 16) = 0x3F803F80u;
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 32) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar + B_X_FULL + b, X_THREADS);
      mbar_init(bar + B_X_FREE + b, 1);
      mbar_init(bar + B_W_FULL + b, 1);
      mbar_init(bar + B_W_FREE + b, 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(bar + B_D_FULL + b, 1);
      mbar_init(bar + B_D_FREE + b, E_THREADS);
    }
    mbar_init(bar + B_Z_FULL, E_THREADS);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_WARP) {
    // =========================== weight chunks: L2 -> smem ring ===========================
    if (lane == 0) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.w1_t);
      const int jobs = MODE == 0 ? nt * N_CH : (nt > 0 ? 1 : 0);   // MODE 1: the single weight block stays resident
      int c = 0;
      for (int j = 0; j < jobs; ++j) {
        const int slot = j % NWS;
        if (j >= NWS) mbar_wait(bar + B_W_FREE + slot, ((j / NWS) - 1) & 1);
        if (SMB_DBG(a, 1) && j >= NWS) { mbar_arrive(bar + B_W_FULL + slot); c = c + 1 == N_CH ? 0 : c + 1; continue; }   // timing: no weight streaming
        mbar_arrive_expect_tx(bar + B_W_FULL + slot, WCH_BYTES);
        bulk_g2s(s_w + slot * WCH_BYTES, src + (size_t)c * WCH_BYTES, WCH_BYTES, bar + B_W_FULL + slot);
        c = c + 1 == N_CH ? 0 : c + 1;
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issue ===========================
    constexpr uint32_t IDESC = idesc_bf16(H, false);
    const uint32_t x_base = smem_u32(s_x), w_base = smem_u32(s_w), w2_base = smem_u32(s_w2);
    int g = 0, j = 0;
    if (MODE == 1 && nt > 0) mbar_wait(bar + B_W_FULL, 0);
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
      const int xs = it % NXS;
      mbar_wait(bar + B_X_FULL + xs, (it / NXS) & 1);
#pragma unroll 1
      for (int s = 0; s < STEPS; ++s, ++g) {
        const int db = g % 3;
        const uint32_t d = tmem + (uint32_t)db * 128u;
        if (s == G2_STEP) mbar_wait(bar + B_Z_FULL, it & 1);
        else if (MODE == 0) mbar_wait(bar + B_W_FULL + (j & 1), (j >> 1) & 1);
        if (g >= 3) mbar_wait(bar + B_D_FREE + db, (g / 3 - 1) & 1);
        fence_after_sync();
        if (lane == 0) {
          if (s != G2_STEP) {
            const uint32_t xa = x_base + xs * X_BYTES, wb = w_base + (MODE == 0 ? (j & 1) : 0) * WCH_BYTES;
#pragma unroll
            for (int ks = 0; ks < KX / 16; ++ks)
              mma_ss(d, smem_desc(xa + ks * 256, 128, X_SBO), smem_desc(wb + ks * 256, 128, X_SBO), IDESC, ks > 0);
            if (MODE == 0) mma_commit(bar + B_W_FREE + (j & 1));
          } else {
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks)
              mma_ts(d, tmem + Z_COL + ks * 8, smem_desc(w2_base + ks * 256, 128, 2048), IDESC, ks > 0);
          }
          mma_commit(bar + B_D_FULL + db);
          if (s == (MODE == 0 ? STEPS - 2 : 0)) mma_commit(bar + B_X_FREE + xs);   // last GEMM that reads this X slot
        }
        __syncwarp();
        if (s != G2_STEP) ++j;
      }
    }
  } else if (warp < E_WARP0) {
    // =========================== X tiles: fp32 -> bf16 operand layout ===========================
    const int xw = warp - X_WARP0;
    const int rr = lane & 7, kq = lane >> 3;
    constexpr int KSTEPS = MODE == 0 ? 5 : 8;          // groups of 4 x (8 columns) per row: 128 h + 32 inv | 128 agg + 128 h
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int xs = it % NXS;
      unsigned char* xb = s_x + xs * X_BYTES;
      int mol[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const int grow = tile * TM + xw * 32 + rb * 8 + rr;
        mol[rb] = grow < a.n_atoms ? (MODE == 0 ? __ldg(a.atom_mol + grow) : 0) : -1;
      }
      if (it >= NXS) mbar_wait(bar + B_X_FREE + xs, ((it / NXS) - 1) & 1);
      constexpr int RPB = MODE == 0 ? 2 : 1;           // row blocks in flight per thread (register budget)
#pragma unroll
      for (int rp = 0; rp < 4 / RPB; ++rp) {
        float4 v[RPB][KSTEPS][2];
#pragma unroll
        for (int q = 0; q < RPB; ++q) {
          const int rb = rp * RPB + q;
          const int grow = tile * TM + xw * 32 + rb * 8 + rr;
          if (mol[rb] >= 0 && !SMB_DBG(a, 4)) {
            // row-major [N][128] or tile image [block][32 column groups][128 rows][4] (NodeArgs::xa_image): eight consecutive
            // columns are two float4 pieces `pst` apart; the next 32 columns are 8 pieces further
            const int rt = grow - tile * TM;
            const bool ia = MODE == 0 && a.xa_image;
            const float4* hp = reinterpret_cast<const float4*>(ia ? a.xa + (size_t)tile * (TM * H) + (size_t)(kq * 2) * (TM * 4) + rt * 4
                                                                  : a.xa + (size_t)grow * H + kq * 8);
            const int pst = ia ? TM : 1;
#pragma unroll
            for (int st = 0; st < 4; ++st) { v[q][st][0] = __ldg(hp + st * 8 * pst); v[q][st][1] = __ldg(hp + (st * 8 + 1) * pst); }
            if (MODE == 0) {
              const float4* ip = reinterpret_cast<const float4*>(a.xb + (size_t)mol[rb] * kShape + kq * 8);
              v[q][4][0] = __ldg(ip); v[q][4][1] = __ldg(ip + 1);
            } else {
              const bool ib = a.xb_image != 0;
              const float4* h2 = reinterpret_cast<const float4*>(ib ? a.xb + (size_t)tile * (TM * H) + (size_t)(kq * 2) * (TM * 4) + rt * 4
                                                                    : a.xb + (size_t)grow * H + kq * 8);
              const int pst2 = ib ? TM : 1;
#pragma unroll
              for (int st = 0; st < 4; ++st) { v[q][4 + st][0] = __ldg(h2 + st * 8 * pst2); v[q][4 + st][1] = __ldg(h2 + (st * 8 + 1) * pst2); }
            }
          } else {
#pragma unroll
            for (int st = 0; st < KSTEPS; ++st) v[q][st][0] = v[q][st][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int q = 0; q < RPB; ++q) {
          const int r = xw * 32 + (rp * RPB + q) * 8 + rr;
          unsigned char* row = xb + (r >> 3) * X_SBO + (r & 7) * 16;
#pragma unroll
          for (int st = 0; st < KSTEPS; ++st) {
            const int kg = st * 4 + kq;
            const float4 lo = v[q][st][0], hi = v[q][st][1];
            *reinterpret_cast<uint4*>(row + kg * 128) =
                make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
          }
        }
      }
      fence_async_smem();
      mbar_arrive(bar + B_X_FULL + xs);
    }
  } else {
    // =========================== epilogue: thread = (row, column half) ===========================
    const int ew = warp - E_WARP0;
    const int qd = warp & 3, half = ew >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
    unsigned char* img = reinterpret_cast<unsigned char*>(a.out1_h);
    int g = 0;
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int grow = tile * TM + r;
      const bool valid = grow < a.n_atoms;
      int ma = 0, mn = 1;
      if (MODE == 0 && valid) { const int m = __ldg(a.atom_mol + grow); ma = __ldg(a.mol_ptr + m); mn = __ldg(a.mol_ptr + m + 1) - ma; }
      unsigned char* img_row = img + (size_t)ma * 1024 + (size_t)(grow - ma) * 16;
#pragma unroll 1
      for (int s = 0; s < STEPS; ++s, ++g) {
        const int db = g % 3;
        mbar_wait(bar + B_D_FULL + db, (g / 3) & 1);
        fence_after_sync();
        uint32_t v[64];
        tmem_ld32(lane_addr + db * 128 + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(lane_addr + db * 128 + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        wait_ld();
        fence_before_sync();
        mbar_arrive(bar + B_D_FREE + db);
        if (s == 0) {
          // LayerNorm folded into the operands: the accumulator is centred and carries sign(gamma)
          const float sq = ln_sumsq64(v);
          float* st = s_stat + (it & 1) * (2 * TM);
          st[half * TM + r] = sq;
          named_sync(BAR_LN + qd, 64);   // the row's other column half is in the warp of the same TMEM lane quadrant
          const float rstd = rsqrtf((sq + st[(half ^ 1) * TM + r]) * (1.f / H) + 1e-5f);
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            uint32_t zp[16];
            ln_apply32(v, hp * 32, rstd, s_beta + half * 64 + hp * 32, zp);
            tmem_st16(lane_addr + Z_COL + half * 32 + hp * 16, zp);
          }
          wait_st();
          fence_before_sync();
          mbar_arrive(bar + B_Z_FULL);
        } else if (s == G2_STEP) {
          if (MODE == 0) {
            // q goes out as bf16, pre-multiplied by log2(e) / sqrt(head_dim) (the K edge role's softmax works in base 2), in the
            // chunk image [128-row block][16 heads][128 rows][16 bytes] (smb_layout.h): a warp's 32 rows store 512 contiguous
            // bytes per head, and the K role cp.asyncs one 16-byte chunk per (destination, head) straight into its MMA operand
            if (valid && !SMB_DBG(a, 2)) {
              constexpr float qs = 0.35355339059327373f * 1.4426950408889634f;
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.out2) + (size_t)tile * kQChunkBlockBytes +
                                                    (size_t)(half * 8) * 2048 + r * 16);
#pragma unroll
              for (int hd = 0; hd < 8; ++hd) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_b2 + half * 64 + hd * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(s_b2 + half * 64 + hd * 8 + 4);
                const int e = hd * 8;
                dst[hd * 128] = make_uint4(pack_bf16((__uint_as_float(v[e]) + b0.x) * qs, (__uint_as_float(v[e + 1]) + b0.y) * qs),
                                           pack_bf16((__uint_as_float(v[e + 2]) + b0.z) * qs, (__uint_as_float(v[e + 3]) + b0.w) * qs),
                                           pack_bf16((__uint_as_float(v[e + 4]) + b1.x) * qs, (__uint_as_float(v[e + 5]) + b1.y) * qs),
                                           pack_bf16((__uint_as_float(v[e + 6]) + b1.z) * qs, (__uint_as_float(v[e + 7]) + b1.w) * qs));
              }
            }
          } else if (valid && !SMB_DBG(a, 2)) {
            // h' : row-major [N][128] or the tile image [128-row block][32 column groups][128 rows][4 floats] between layers
            const bool oi = a.out2_image;
            float4* dst = oi ? reinterpret_cast<float4*>(a.out2 + (size_t)tile * (TM * H) + (size_t)(half * 16) * (TM * 4) + r * 4)
                             : reinterpret_cast<float4*>(a.out2 + (size_t)grow * H + half * 64);
            const int DSTEP = oi ? TM : 1;   // float4 stride between consecutive column groups
            const bool ri = a.xb_image;      // the residual is the h operand (same layout)
            const float4* res = ri ? reinterpret_cast<const float4*>(a.residual + (size_t)tile * (TM * H) + (size_t)(half * 16) * (TM * 4) + r * 4)
                                   : reinterpret_cast<const float4*>(a.residual + (size_t)grow * H + half * 64);
            const int RSTEP = ri ? TM : 1;
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
              float4 bb = *reinterpret_cast<const float4*>(s_b2 + half * 64 + e);
              const float4 rv = __ldg(res + (e / 4) * RSTEP);
              bb.x += rv.x; bb.y += rv.y; bb.z += rv.z; bb.w += rv.w;
              dst[(e / 4) * DSTEP] = make_float4(__uint_as_float(v[e]) + bb.x, __uint_as_float(v[e + 1]) + bb.y, __uint_as_float(v[e + 2]) + bb.z,
                                       __uint_as_float(v[e + 3]) + bb.w);
            }
          }
        } else if (MODE == 0 && valid && !SMB_DBG(a, 2)) {
          const int part = s - 1;
          unsigned char* dst = img_row + (size_t)part * mn * 256 + (size_t)(half * 8) * mn * 16;
#pragma unroll
          for (int cg = 0; cg < 8; ++cg) {
            const int e = cg * 8;
            *reinterpret_cast<uint4*>(dst + (size_t)cg * mn * 16) =
                make_uint4(pack_bf16(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), pack_bf16(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])),
                           pack_bf16(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])), pack_bf16(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])));
          }
        }
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

}  // namespace

bool node_tc5_supported(const smb_model_dims& d, int n_max) {
  // always together with the warp-specialised edge pipeline: q travels between them as a pre-scaled bf16 chunk image
  return edge_ws_supported(d, n_max);
}

template <int MODE>
static int launch_tc5(const NodeArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 0) return 0;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem(node_tc5_kernel<MODE>, Cfg<MODE>::SMEM_TOTAL, configured)) return rc;
  const int n_sm = device_sm_count();
  const int n_tiles = (a.n_atoms + TM - 1) / TM;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  NodeArgs b = a;
#ifdef SMB_DEBUG
  static const int dbg = getenv("SMB_NODE_DBG") ? atoi(getenv("SMB_NODE_DBG")) : 0;   // timing ablations (results are wrong when set)
  b.dbg = dbg;
#else
  b.dbg = 0;
#endif
  node_tc5_kernel<MODE><<<grid, THREADS, Cfg<MODE>::SMEM_TOTAL, st>>>(b);
  return (int)cudaGetLastError();
}

int launch_node_pre_tc5(const NodeArgs& a, cudaStream_t st) { return launch_tc5<0>(a, st); }
int launch_node_out_tc5(const NodeArgs& a, cudaStream_t st) { return launch_tc5<1>(a, st); }

}  // namespace smb
