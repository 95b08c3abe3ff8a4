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
constexpr int KX = kNodeKx;                  // 176
constexpr int X_SBO = (KX / 8) * 128;        // 2816
constexpr int X_BYTES = TM * KX * 2;         // 45056
constexpr int WCH_BYTES = kNodeChunkBytes;   // 45056
constexpr int N_CH = 5;                      // hidden | A_k | B_k | A_v | B_v
constexpr int W2_BYTES = H * H * 2;
constexpr int W_WARP = 0, MMA_WARP = 1, X_WARP0 = 2, X_WARPS = 4, E_WARP0 = 6, E_WARPS = 8, WARPS = 14, THREADS = WARPS * 32;
constexpr int E_THREADS = E_WARPS * 32, X_THREADS = X_WARPS * 32;
constexpr int BAR_LN = 1;
constexpr uint32_t Z_COL = 384;

enum { B_X_FULL = 0, B_X_FREE = 2, B_W_FULL = 4, B_W_FREE = 6, B_D_FULL = 8, B_D_FREE = 11, B_Z_FULL = 14, N_BARS = 15 };

constexpr int o_bar = 0;
constexpr int o_tmem = 128;
constexpr int o_beta = 256;                  // beta / |gamma| of the query MLP's LayerNorm
constexpr int o_b2 = o_beta + 512;
constexpr int o_stat = o_b2 + 512;           // float[2 tiles][2 halves][128 rows]
constexpr int o_w2 = o_stat + 2048;          // 3328 -> aligned 128
constexpr int o_x = o_w2 + W2_BYTES;
constexpr int o_w = o_x + 2 * X_BYTES;
constexpr int SMEM_TOTAL = o_w + 2 * WCH_BYTES;
static_assert(o_w2 % 128 == 0 && SMEM_TOTAL <= 227 * 1024, "shared memory plan");

__global__ void __launch_bounds__(THREADS, 1) node_pre_tc5_kernel(NodeArgs a) {
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
    for (int p = tid; p < 2 * X_BYTES / 16; p += THREADS) z0[p] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int p = tid; p < 2 * TM; p += THREADS) {
    const int slot = p >> 7, r = p & 127;
    *reinterpret_cast<uint32_t*>(s_x + slot * X_BYTES + (r >> 3) * X_SBO + (160 / 8) * 128 + (r & 7) * 16) = 0x3F803F80u;
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
      const int jobs = nt * N_CH;
      int c = 0;
      for (int j = 0; j < jobs; ++j) {
        const int slot = j & 1;
        if (j >= 2) mbar_wait(bar + B_W_FREE + slot, ((j >> 1) - 1) & 1);
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
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
      const int xs = it & 1;
      mbar_wait(bar + B_X_FULL + xs, (it >> 1) & 1);
#pragma unroll 1
      for (int s = 0; s < 6; ++s, ++g) {
        const int db = g % 3;
        const uint32_t d = tmem + (uint32_t)db * 128u;
        if (s != 3) mbar_wait(bar + B_W_FULL + (j & 1), (j >> 1) & 1);
        else mbar_wait(bar + B_Z_FULL, it & 1);
        if (g >= 3) mbar_wait(bar + B_D_FREE + db, (g / 3 - 1) & 1);
        fence_after_sync();
        if (lane == 0) {
          if (s != 3) {
            const uint32_t xa = x_base + xs * X_BYTES, wb = w_base + (j & 1) * WCH_BYTES;
#pragma unroll
            for (int ks = 0; ks < KX / 16; ++ks)
              mma_ss(d, smem_desc(xa + ks * 256, 128, X_SBO), smem_desc(wb + ks * 256, 128, X_SBO), IDESC, ks > 0);
            mma_commit(bar + B_W_FREE + (j & 1));
          } else {
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks)
              mma_ts(d, tmem + Z_COL + ks * 8, smem_desc(w2_base + ks * 256, 128, 2048), IDESC, ks > 0);
          }
          mma_commit(bar + B_D_FULL + db);
          if (s == 5) mma_commit(bar + B_X_FREE + xs);
        }
        __syncwarp();
        if (s != 3) ++j;
      }
    }
  } else if (warp < E_WARP0) {
    // =========================== X tiles: fp32 -> bf16 operand layout ===========================
    const int xw = warp - X_WARP0;
    const int rr = lane & 7, kq = lane >> 3;
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int xs = it & 1;
      unsigned char* xb = s_x + xs * X_BYTES;
      int mol[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const int grow = tile * TM + xw * 32 + rb * 8 + rr;
        mol[rb] = grow < a.n_atoms ? __ldg(a.atom_mol + grow) : -1;
      }
      if (it >= 2) mbar_wait(bar + B_X_FREE + xs, ((it >> 1) - 1) & 1);
#pragma unroll
      for (int rp = 0; rp < 2; ++rp) {
        float4 v[2][5][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rb = rp * 2 + q;
          const int grow = tile * TM + xw * 32 + rb * 8 + rr;
          if (mol[rb] >= 0) {
            const float4* hp = reinterpret_cast<const float4*>(a.xa + (size_t)grow * H + kq * 8);
#pragma unroll
            for (int st = 0; st < 4; ++st) { v[q][st][0] = __ldg(hp + st * 8); v[q][st][1] = __ldg(hp + st * 8 + 1); }
            const float4* ip = reinterpret_cast<const float4*>(a.xb + (size_t)mol[rb] * kShape + kq * 8);
            v[q][4][0] = __ldg(ip); v[q][4][1] = __ldg(ip + 1);
          } else {
#pragma unroll
            for (int st = 0; st < 5; ++st) v[q][st][0] = v[q][st][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int r = xw * 32 + (rp * 2 + q) * 8 + rr;
          unsigned char* row = xb + (r >> 3) * X_SBO + (r & 7) * 16;
#pragma unroll
          for (int st = 0; st < 5; ++st) {
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
      if (valid) { const int m = __ldg(a.atom_mol + grow); ma = __ldg(a.mol_ptr + m); mn = __ldg(a.mol_ptr + m + 1) - ma; }
      unsigned char* img_row = img + (size_t)ma * 1024 + (size_t)(grow - ma) * 16;
#pragma unroll 1
      for (int s = 0; s < 6; ++s, ++g) {
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
          named_sync(BAR_LN, E_THREADS);
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
        } else if (s == 3) {
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.out2 + (size_t)grow * H + half * 64);
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(s_b2 + half * 64 + e);
              dst[e / 4] = make_float4(__uint_as_float(v[e]) + bb.x, __uint_as_float(v[e + 1]) + bb.y, __uint_as_float(v[e + 2]) + bb.z,
                                       __uint_as_float(v[e + 3]) + bb.w);
            }
          }
        } else if (valid) {
          const int part = s < 3 ? s - 1 : s - 2;
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
  static const bool off = getenv("SMB_NODE_LEGACY") != nullptr;   // debugging aid
  return !off && edge_ws_supported(d, n_max);
}

int launch_node_pre_tc5(const NodeArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 0) return 0;
  static bool configured = false;
  static int n_sm = 148;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(node_pre_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) n_sm = n;
    configured = true;
  }
  const int n_tiles = (a.n_atoms + TM - 1) / TM;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  node_pre_tc5_kernel<<<grid, THREADS, SMEM_TOTAL, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace smb
