// Gathered fp32 GEMM on tcgen05 / TMEM with split-bf16 ("bf16x3") products:
//
//     C[M x N] (+)= [ A_0[idx_0[m]] | A_1[idx_1[m]] | ... ] . W[N][K]^T + bias          (fp32 in, fp32 out, fp32 accumulate)
//
// The A operand is the CONCATENATION of up to four gathered row segments -- exactly how the reference builds the edge-MLP input
// kv = [r | h_dst | h_src | inv_dst] (models/uni_transformer.py:61-63) -- so one launch replaces the four accumulating SGEMMs
// of the generic path and its [E, hidden] intermediate is written once.  Also used (batched, one cloud per grid.z) for the
// VN-DGCNN encoder's node GEMMs and Gram matrices (models/shape_vn_layers.py:257-292).
//
// Every operand element is split into bf16 pieces while it is staged to shared memory.  Two pieces x = hi + lo (16 significant
// bits), three MMAs per K = 16 step (hi.hi + lo.hi + hi.lo): products agree with fp32 to ~2^-16, which keeps the 1e-3 parity
// bar of the denoiser with a wide margin.  Three pieces (TcGemmArgs::split3; 24 significant bits, six MMAs): fp32-level
// products, for the encoder, whose dynamic kNN graphs are selected on these values -- a 1e-5 perturbation of the features
// flips near-tie neighbours against the reference's fp32 graph.  The kernel is tensor-pipe bound by construction (3 x 128 x N x 16 per
// step against ~2.5 staged elements per thread and step).
//
// CTA = 128 rows x N_TILE <= 256 columns, K consumed in chunks of 32 (16 with three pieces) through a two-stage shared-memory
// ring: eight warps stage (gather with a one-chunk register prefetch, split, K-major store) and arrive on the stage's "full"
// mbarrier, a ninth warp issues the MMAs and commits the stage's "free" mbarrier; two CTAs per SM (<= 96 KB of shared memory,
// 256 TMEM columns each), so one CTA's epilogue overlaps the other's main loop.
#include "smb_common.cuh"
#include "smb_kernels.h"
#include "smb_tc.cuh"

#include <type_traits>

namespace smb {

namespace {

using namespace tc;

constexpr int TM = 128, NT_MAX = 256, STAGERS = 256, THREADS = STAGERS + 32;   // 8 staging / epilogue warps + the MMA issuer

// SPLIT = 2: x = hi + lo            (16 significant bits), products hh + lh + hl                    (~2^-16 per product)
// SPLIT = 3: x = hi + mid + lo      (24 significant bits), products hh + hm + mh + mm + hl + lh     (~2^-23: fp32 level)
template <int SPLIT>
struct Cfg {
  static constexpr int KC = SPLIT == 3 ? 16 : 32;          // k per stage: keeps two CTAs (<= 96 KB each) per SM
  static constexpr int A_BYTES = TM * KC * 2;              // one piece
  static constexpr int W_BYTES = NT_MAX * KC * 2;
  static constexpr int STAGE_BYTES = SPLIT * (A_BYTES + W_BYTES);
  static constexpr int SBO = (KC / 8) * 128;
  static constexpr int SMEM_TOTAL = 128 + 4 * NT_MAX * 4 + 2 * STAGE_BYTES;   // mbarriers | bias, LN gamma, LN beta of the column tile, LN partial sums | two stages
  static constexpr int N_TERMS = SPLIT == 3 ? 6 : 3;
};

// eight consecutive values -> one 16-byte K-major chunk per piece
template <int SPLIT>
__device__ __forceinline__ void split_store(const float (&v)[8], unsigned char* dst, int piece_stride) {
  uint32_t p[SPLIT][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a = v[2 * q], b = v[2 * q + 1];
#pragma unroll
    for (int s = 0; s < SPLIT; ++s) {
      p[s][q] = pack_bf16(a, b);
      if (s + 1 < SPLIT) { a -= __uint_as_float(p[s][q] << 16); b -= __uint_as_float(p[s][q] & 0xffff0000u); }
    }
  }
#pragma unroll
  for (int s = 0; s < SPLIT; ++s) *reinterpret_cast<uint4*>(dst + s * piece_stride) = make_uint4(p[s][0], p[s][1], p[s][2], p[s][3]);
}

// four consecutive values -> 8 bytes (half a K-major chunk) per piece
template <int SPLIT>
__device__ __forceinline__ void split_store4(const float4& x, unsigned char* dst, int piece_stride) {
  float a = x.x, b = x.y, c = x.z, d = x.w;
#pragma unroll
  for (int s = 0; s < SPLIT; ++s) {
    const uint32_t p0 = pack_bf16(a, b), p1 = pack_bf16(c, d);
    *reinterpret_cast<uint2*>(dst + s * piece_stride) = make_uint2(p0, p1);
    if (s + 1 < SPLIT) {
      a -= __uint_as_float(p0 << 16); b -= __uint_as_float(p0 & 0xffff0000u);
      c -= __uint_as_float(p1 << 16); d -= __uint_as_float(p1 & 0xffff0000u);
    }
  }
}

__device__ __forceinline__ float4 load4(const float* p, int k, int kmax, bool vec) {
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && k + 3 < kmax) return __ldg(reinterpret_cast<const float4*>(p + k));
  if (k < kmax) x.x = __ldg(p + k);
  if (k + 1 < kmax) x.y = __ldg(p + k + 1);
  if (k + 2 < kmax) x.z = __ldg(p + k + 2);
  if (k + 3 < kmax) x.w = __ldg(p + k + 3);
  return x;
}

// chunk c of the concatenated K range -> (segment, first k inside it)
__device__ __forceinline__ void chunk_of(const TcGemmArgs& g, int c, int KC, int& s, int& k0) {
  s = 0; k0 = c * KC;
  while (k0 >= g.seg[s].k) { k0 -= (g.seg[s].k + KC - 1) / KC * KC; ++s; }
}

// Pre-split image of a shared W: for every (column tile j, chunk c) the SPLIT x [256 n][KC k] bf16 K-major blocks exactly as a
// pipeline stage holds them, so that a CTA fetches a chunk of W with ONE bulk copy.  grid = (chunks, column tiles), 256 threads.
template <int SPLIT>
__global__ void __launch_bounds__(256) tc_wsplit_kernel(TcGemmArgs g, int n_chunks) {
  using C = Cfg<SPLIT>;
  constexpr int KC = C::KC, W_BYTES = C::W_BYTES, SBO = C::SBO;
  const int c = blockIdx.x, n0 = blockIdx.y * NT_MAX;
  int s, k0;
  chunk_of(g, c, KC, s, k0);
  const TcGemmSeg& sg = g.seg[s];
  unsigned char* dst = reinterpret_cast<unsigned char*>(g.w_img) + ((size_t)blockIdx.y * n_chunks + c) * (SPLIT * W_BYTES);
  for (int it = threadIdx.x; it < NT_MAX * (KC / 8); it += 256) {
    const int n = it / (KC / 8), kg = it % (KC / 8);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + kg * 8 + e;
      v[e] = (n0 + n < g.N && k < sg.k) ? __ldg(g.W + (long long)(n0 + n) * g.ldw + sg.w_off + k) : 0.f;
    }
    split_store<SPLIT>(v, dst + (n >> 3) * SBO + kg * 128 + (n & 7) * 16, W_BYTES);
  }
}

template <int SPLIT, bool IMG>
__global__ void __launch_bounds__(THREADS, 2) tc_gemm_kernel(TcGemmArgs g) {
  using C = Cfg<SPLIT>;
  constexpr int KC = C::KC, A_BYTES = C::A_BYTES, W_BYTES = C::W_BYTES, STAGE_BYTES = C::STAGE_BYTES, SBO = C::SBO;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);            // [0], [1]: stage free; [2]: all MMAs done; [3], [4]: stage full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  float* s_bias = reinterpret_cast<float*>(smem + 128);
  float* s_gamma = s_bias + NT_MAX;
  float* s_beta = s_gamma + NT_MAX;
  float* s_red = s_beta + NT_MAX;      // [2 column halves][128 rows]
  unsigned char* stages = smem + 128 + 4 * NT_MAX * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * NT_MAX;
  const long long bz = blockIdx.z;
  const int n_left = g.N - n0;
  const int n_tile = n_left >= NT_MAX ? NT_MAX : (n_left + 15) & ~15;

  for (int i = tid; i < NT_MAX; i += THREADS) {
    const bool in = blockIdx.y * NT_MAX + i < g.N;
    s_bias[i] = (g.bias && in) ? __ldg(g.bias + blockIdx.y * NT_MAX + i) : 0.f;
    s_gamma[i] = (g.ln_gamma && in) ? __ldg(g.ln_gamma + i) : 0.f;
    s_beta[i] = (g.ln_gamma && in) ? __ldg(g.ln_beta + i) : 0.f;
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  if (tid == 32) {
    mbar_init(bar + 0, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1);
    mbar_init(bar + 3, STAGERS); mbar_init(bar + 4, STAGERS);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_bf16(n_tile, false);

  // Staging map (coalesced): a chunk row is KC / 4 pieces of 16 bytes; consecutive lanes take consecutive pieces of a row, so a
  // warp's load touches 4 (8) whole rows = 4 (8) cache lines instead of one line per lane.  Thread t: piece t % PPR of the rows
  // t / PPR + (256 / PPR) i.  The global loads of chunk c + 1 are issued (into registers) right after chunk c has been stored
  // to shared memory, so their latency overlaps the MMA issue and the wait for the stage to drain.
  constexpr int PPR = KC / 4;                        // pieces per row
  constexpr int RSTEP = STAGERS / PPR;               // rows covered by one pass of the 256 staging threads
  constexpr int NA = TM / RSTEP, NW = NT_MAX / RSTEP;   // pieces per thread and chunk
  const int sp = tid % PPR, sr = tid / PPR;
  int n_chunks = 0;
  for (int s = 0; s < g.n_segs; ++s) n_chunks += (g.seg[s].k + KC - 1) / KC;

  if (warp == STAGERS / 32) {
    // =========== MMA issuer: its own warp, so that issuing (~60 cycles per MMA, ~150 per commit) never delays the staging ===========
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int st = c & 1;
      mbar_wait<32>(bar + 3 + st, (c >> 1) & 1);
      fence_after_sync();
      if (lane == 0) {
        const uint32_t a0 = smem_u32(stages + st * STAGE_BYTES), w0 = a0 + SPLIT * A_BYTES;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks)
#pragma unroll
          for (int t = 0; t < C::N_TERMS; ++t) {
            const int pa = SPLIT == 3 ? (t == 1 || t == 3 ? 1 : t == 5 ? 2 : 0) : (t == 1 ? 1 : 0);
            const int pw = SPLIT == 3 ? (t == 2 || t == 3 ? 1 : t == 4 ? 2 : 0) : (t == 2 ? 1 : 0);
            mma_ss(tmem, smem_desc(a0 + pa * A_BYTES + ks * 256, 128, SBO), smem_desc(w0 + pw * W_BYTES + ks * 256, 128, SBO), idesc,
                   (c | ks | t) > 0);
          }
        mma_commit(bar + st);
        if (c + 1 == n_chunks) mma_commit(bar + 2);
      }
      __syncwarp();
    }
  } else {

  // this thread's pieces of the staged chunks (fp32).  Pre-split W (IMG): only A is staged by threads, and its gathered
  // rows are fetched TWO chunks ahead (ncu: the first use of a prefetched piece was the kernel's top stall with one chunk).
  constexpr int DEPTH = IMG ? 2 : 1;
  float4 ra[DEPTH][NA], rw[IMG ? 1 : NW];
  int arow[NA], arow_seg = -1;          // gathered row of each of this thread's pieces in the current segment (-1: none)
  auto fetch = [&](int c, auto SLOT) {  // chunk c -> (segment, k0); registers of slot SLOT
    constexpr int S = decltype(SLOT)::value;
    int s, k0;
    chunk_of(g, c, KC, s, k0);
    const TcGemmSeg& sg = g.seg[s];
    if (s != arow_seg) {                // the row indices change with the segment only: one dependent load per segment, not per chunk
      arow_seg = s;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int m = m0 + sr + RSTEP * i;
        arow[i] = m < g.M ? (sg.idx ? __ldg(sg.idx + bz * g.idx_batch + m) : m) : -1;
      }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i)
      ra[S][i] = arow[i] >= 0 ? load4(sg.a + bz * g.a_batch + (long long)arow[i] * sg.lda, k0 + 4 * sp, sg.k, g.vec != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (IMG) return;
    const float* wb = g.W + bz * g.w_batch + sg.w_off;
#pragma unroll
    for (int i = 0; i < (IMG ? 1 : NW); ++i) {
      const int n = sr + RSTEP * i;
      rw[i] = (n < n_tile && n0 + n < g.N) ? load4(wb + (long long)(n0 + n) * g.ldw, k0 + 4 * sp, sg.k, g.vec != 0)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  const int poff = (sp >> 1) * 128 + (sp & 1) * 8;    // piece position inside a K-major row
  auto stage_chunk = [&](int c, auto SLOT) {
    constexpr int S = decltype(SLOT)::value;
    const int st = c & 1;
    unsigned char* sa = stages + st * STAGE_BYTES;     // A pieces, then W pieces
    unsigned char* sw = sa + SPLIT * A_BYTES;
    if (c >= 2) mbar_wait(bar + st, ((c >> 1) - 1) & 1);   // the MMAs of chunk c - 2 have read this stage
    if (IMG && tid == 0) {
      mbar_expect_tx(bar + 3 + st, SPLIT * W_BYTES);
      bulk_g2s(sw, reinterpret_cast<const unsigned char*>(g.w_img) + ((size_t)blockIdx.y * n_chunks + c) * (SPLIT * W_BYTES), SPLIT * W_BYTES,
               bar + 3 + st);
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int r = sr + RSTEP * i;
      split_store4<SPLIT>(ra[S][i], sa + (r >> 3) * SBO + (r & 7) * 16 + poff, A_BYTES);
    }
    if (!IMG) {
#pragma unroll
      for (int i = 0; i < (IMG ? 1 : NW); ++i) {
        const int n = sr + RSTEP * i;
        if (n < n_tile) split_store4<SPLIT>(rw[i], sw + (n >> 3) * SBO + (n & 7) * 16 + poff, W_BYTES);
      }
    }
    if (c + DEPTH < n_chunks) fetch(c + DEPTH, SLOT);
    fence_async_smem();
    mbar_arrive(bar + 3 + st);
  };
  using I0 = std::integral_constant<int, 0>;
  using I1 = std::integral_constant<int, DEPTH - 1>;
  if (n_chunks > 0) fetch(0, I0{});
  if (DEPTH > 1 && n_chunks > 1) fetch(1, I1{});
  for (int c = 0; c < n_chunks; c += 2) {
    stage_chunk(c, I0{});
    if (c + 1 < n_chunks) stage_chunk(c + 1, I1{});
  }
  }   // staging warps
  if (n_chunks > 0) mbar_wait(bar + 2, 0);
  fence_after_sync();

  // ---- epilogue: thread = row (TMEM lane quadrant warp & 3), the two warps of a quadrant alternate 16-column chunks ----
  if (warp < STAGERS / 32) {
    const int row = (warp & 3) * 32 + lane, m = m0 + row;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* crow = g.C + bz * g.c_batch + (long long)m * g.ldc + n0;
    // fused LayerNorm (the whole row is in this CTA's accumulator: N <= 256): the row's two threads -- the warps of a TMEM lane
    // quadrant alternate 16-column chunks -- exchange partial sums through shared memory; mean, then centred sum of squares
    // (two more passes over TMEM), like the separate row kernel it replaces
    float mean = 0.f, rstd = 1.f;
    if (g.ln_gamma) {
      const int half = warp >> 2;
      auto row_sum = [&](auto f) {
        float acc = 0.f;
        for (int cb = half * 16; cb < n_tile; cb += 32) {
          uint32_t v[16];
          tmem_ld16(lane_addr + cb, v);
          wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) acc = f(acc, __uint_as_float(v[e]) + s_bias[cb + e]);
        }
        s_red[half * TM + row] = acc;
        named_sync(1 + (warp & 3), 64);
        const float tot = acc + s_red[(half ^ 1) * TM + row];
        named_sync(1 + (warp & 3), 64);       // both threads have read before the next pass overwrites
        return tot;
      };
      mean = row_sum([](float a, float x) { return a + x; }) / (float)g.N;
      const float m_ = mean;
      rstd = 1.f / sqrtf(row_sum([m_](float a, float x) { const float d = x - m_; return fmaf(d, d, a); }) / (float)g.N + 1e-5f);
    }
    for (int cb = (warp >> 2) * 16; cb < n_tile; cb += 32) {
      uint32_t v[16];
      tmem_ld16(lane_addr + cb, v);
      float4 bq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bq[q] = *reinterpret_cast<const float4*>(s_bias + cb + 4 * q);   // staged once per CTA (zeros without a bias)
      wait_ld();
      if (m < g.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = cb + 4 * q;
          float o[4];
          const float4 b4 = bq[q];
          o[0] = __uint_as_float(v[4 * q]) + b4.x; o[1] = __uint_as_float(v[4 * q + 1]) + b4.y;
          o[2] = __uint_as_float(v[4 * q + 2]) + b4.z; o[3] = __uint_as_float(v[4 * q + 3]) + b4.w;
          if (g.ln_gamma) {
            const float4 ga = *reinterpret_cast<const float4*>(s_gamma + n), be = *reinterpret_cast<const float4*>(s_beta + n);
            o[0] = fmaxf(fmaf((o[0] - mean) * rstd, ga.x, be.x), 0.f); o[1] = fmaxf(fmaf((o[1] - mean) * rstd, ga.y, be.y), 0.f);
            o[2] = fmaxf(fmaf((o[2] - mean) * rstd, ga.z, be.z), 0.f); o[3] = fmaxf(fmaf((o[3] - mean) * rstd, ga.w, be.w), 0.f);
          }
          if (g.vec_c && n0 + n + 3 < g.N) {
            float4* dst = reinterpret_cast<float4*>(crow + n);
            if (g.accumulate) { const float4 p = *dst; o[0] += p.x; o[1] += p.y; o[2] += p.z; o[3] += p.w; }
            *dst = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n0 + n + e < g.N) crow[n + e] = g.accumulate ? crow[n + e] + o[e] : o[e];
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Persistent variant for big GEMMs with a pre-split shared W and N <= 256 (one column tile): one CTA per SM walks the 128-row
// tiles; the accumulator is double-buffered in TMEM (2 x 256 columns) and the epilogue has its own eight warps, so the epilogue
// of tile t (with the fused LayerNorm: three passes over TMEM plus the stores -- half of a tile's time in the kernel above, where
// the two co-resident CTAs run in lockstep) overlaps the main loop of tile t + 1.  Four-stage operand ring: the staging warps run up
// to four chunks ahead of the MMA issuer, which also hides the latency of W's bulk copies.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int P_STAGES = 4, P_EPI = 256, P_THREADS = STAGERS + 32 + P_EPI;
template <int SPLIT>
struct PCfg {
  static constexpr int SMEM_TOTAL = 256 + 4 * NT_MAX * 4 + P_STAGES * Cfg<SPLIT>::STAGE_BYTES;
};
enum { PB_FULL = 0, PB_FREE = 4, PB_ACC_FULL = 8, PB_ACC_FREE = 10, PB_N = 12 };

template <int SPLIT>
__global__ void __launch_bounds__(P_THREADS, 1) tc_gemm_persistent_kernel(TcGemmArgs g) {
  using C = Cfg<SPLIT>;
  constexpr int KC = C::KC, A_BYTES = C::A_BYTES, W_BYTES = C::W_BYTES, STAGE_BYTES = C::STAGE_BYTES, SBO = C::SBO;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  float* s_bias = reinterpret_cast<float*>(smem + 256);
  float* s_gamma = s_bias + NT_MAX;
  float* s_beta = s_gamma + NT_MAX;
  float* s_red = s_beta + NT_MAX;      // [2 column halves][128 rows]
  unsigned char* stages = smem + 256 + 4 * NT_MAX * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tile = g.N >= NT_MAX ? NT_MAX : (g.N + 15) & ~15;
  const int n_row_tiles = (g.M + TM - 1) / TM;
  int n_chunks = 0;
  for (int s = 0; s < g.n_segs; ++s) n_chunks += (g.seg[s].k + KC - 1) / KC;

  for (int i = tid; i < NT_MAX; i += P_THREADS) {
    const bool in = i < g.N;
    s_bias[i] = (g.bias && in) ? __ldg(g.bias + i) : 0.f;
    s_gamma[i] = (g.ln_gamma && in) ? __ldg(g.ln_gamma + i) : 0.f;
    s_beta[i] = (g.ln_gamma && in) ? __ldg(g.ln_beta + i) : 0.f;
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  if (tid == 32) {
    for (int b = 0; b < P_STAGES; ++b) { mbar_init(bar + PB_FULL + b, STAGERS); mbar_init(bar + PB_FREE + b, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar + PB_ACC_FULL + b, 1); mbar_init(bar + PB_ACC_FREE + b, P_EPI); }
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_bf16(n_tile, false);

  if (warp < STAGERS / 32) {
    // =========================== A staging (the map of tc_gemm_kernel) + W bulk copies ===========================
    constexpr int PPR = KC / 4, RSTEP = STAGERS / PPR, NA = TM / RSTEP;
    const int sp = tid % PPR, sr = tid / PPR;
    const int poff = (sp >> 1) * 128 + (sp & 1) * 8;
    float4 ra[2][NA];
    int arow[NA], arow_seg;
    int gc = 0;                                  // chunks staged so far (all tiles): ring position
    for (int tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
      const int m0 = tile * TM;
      arow_seg = -1;
      auto fetch = [&](int c, auto SLOT) {
        constexpr int S = decltype(SLOT)::value;
        int s, k0;
        chunk_of(g, c, KC, s, k0);
        const TcGemmSeg& sg = g.seg[s];
        if (s != arow_seg) {
          arow_seg = s;
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            const int m = m0 + sr + RSTEP * i;
            arow[i] = m < g.M ? (sg.idx ? __ldg(sg.idx + m) : m) : -1;
          }
        }
#pragma unroll
        for (int i = 0; i < NA; ++i)
          ra[S][i] = arow[i] >= 0 ? load4(sg.a + (long long)arow[i] * sg.lda, k0 + 4 * sp, sg.k, g.vec != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto stage_chunk = [&](int c, auto SLOT) {
        constexpr int S = decltype(SLOT)::value;
        const int st = gc % P_STAGES;
        unsigned char* sa = stages + st * STAGE_BYTES;
        unsigned char* sw = sa + SPLIT * A_BYTES;
        if (gc >= P_STAGES) mbar_wait(bar + PB_FREE + st, ((gc / P_STAGES) - 1) & 1);   // the MMAs of the chunk P_STAGES back have read this stage
        if (tid == 0) {
          mbar_expect_tx(bar + PB_FULL + st, SPLIT * W_BYTES);
          bulk_g2s(sw, reinterpret_cast<const unsigned char*>(g.w_img) + (size_t)c * (SPLIT * W_BYTES), SPLIT * W_BYTES, bar + PB_FULL + st);
        }
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const int r = sr + RSTEP * i;
          split_store4<SPLIT>(ra[S][i], sa + (r >> 3) * SBO + (r & 7) * 16 + poff, A_BYTES);
        }
        if (c + 2 < n_chunks) fetch(c + 2, SLOT);
        fence_async_smem();
        mbar_arrive(bar + PB_FULL + st);
        ++gc;
      };
      using I0 = std::integral_constant<int, 0>;
      using I1 = std::integral_constant<int, 1>;
      if (n_chunks > 0) fetch(0, I0{});
      if (n_chunks > 1) fetch(1, I1{});
      for (int c = 0; c < n_chunks; c += 2) {
        stage_chunk(c, I0{});
        if (c + 1 < n_chunks) stage_chunk(c + 1, I1{});
      }
    }
  } else if (warp == STAGERS / 32) {
    // =========================== MMA issuer ===========================
    int gc = 0, it = 0;
    for (int tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      if (it >= 2) mbar_wait<32>(bar + PB_ACC_FREE + ab, ((it >> 1) - 1) & 1);   // the epilogue of tile it - 2 has drained this accumulator
      fence_after_sync();
      const uint32_t d = tmem + (uint32_t)ab * NT_MAX;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c, ++gc) {
        const int st = gc % P_STAGES;
        mbar_wait<32>(bar + PB_FULL + st, (gc / P_STAGES) & 1);
        fence_after_sync();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(stages + st * STAGE_BYTES), w0 = a0 + SPLIT * A_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)
#pragma unroll
            for (int t = 0; t < C::N_TERMS; ++t) {
              const int pa = SPLIT == 3 ? (t == 1 || t == 3 ? 1 : t == 5 ? 2 : 0) : (t == 1 ? 1 : 0);
              const int pw = SPLIT == 3 ? (t == 2 || t == 3 ? 1 : t == 4 ? 2 : 0) : (t == 2 ? 1 : 0);
              mma_ss(d, smem_desc(a0 + pa * A_BYTES + ks * 256, 128, SBO), smem_desc(w0 + pw * W_BYTES + ks * 256, 128, SBO), idesc,
                     (c | ks | t) > 0);
            }
          mma_commit(bar + PB_FREE + st);
          if (c + 1 == n_chunks) mma_commit(bar + PB_ACC_FULL + ab);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue warps: thread = (row, 16-column chunks of one parity) ===========================
    const int ew = warp - (STAGERS / 32 + 1);          // 0..7
    const int qd = warp & 3, half = ew >> 2 & 1;       // TMEM lane quadrant = warp id mod 4
    // the two warps of a quadrant must differ in `half`: ew and ew ^ 4 share (warp & 3)
    const int row = qd * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1, m = tile * TM + row;
      mbar_wait(bar + PB_ACC_FULL + ab, (it >> 1) & 1);
      fence_after_sync();
      const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)ab * NT_MAX;
      float* crow = g.C + (long long)m * g.ldc;
      float mean = 0.f, rstd = 1.f;
      if (g.ln_gamma) {
        auto row_sum = [&](auto f) {
          float acc = 0.f;
          for (int cb = half * 16; cb < n_tile; cb += 32) {
            uint32_t v[16];
            tmem_ld16(lane_addr + cb, v);
            wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) acc = f(acc, __uint_as_float(v[e]) + s_bias[cb + e]);
          }
          s_red[half * TM + row] = acc;
          named_sync(1 + qd, 64);
          const float tot = acc + s_red[(half ^ 1) * TM + row];
          named_sync(1 + qd, 64);
          return tot;
        };
        mean = row_sum([](float a, float x) { return a + x; }) / (float)g.N;
        const float m_ = mean;
        rstd = 1.f / sqrtf(row_sum([m_](float a, float x) { const float dd = x - m_; return fmaf(dd, dd, a); }) / (float)g.N + 1e-5f);
      }
      for (int cb = half * 16; cb < n_tile; cb += 32) {
        uint32_t v[16];
        tmem_ld16(lane_addr + cb, v);
        float4 bq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bq[q] = *reinterpret_cast<const float4*>(s_bias + cb + 4 * q);
        wait_ld();
        if (m < g.M) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = cb + 4 * q;
            float o[4];
            const float4 b4 = bq[q];
            o[0] = __uint_as_float(v[4 * q]) + b4.x; o[1] = __uint_as_float(v[4 * q + 1]) + b4.y;
            o[2] = __uint_as_float(v[4 * q + 2]) + b4.z; o[3] = __uint_as_float(v[4 * q + 3]) + b4.w;
            if (g.ln_gamma) {
              const float4 ga = *reinterpret_cast<const float4*>(s_gamma + n), be = *reinterpret_cast<const float4*>(s_beta + n);
              o[0] = fmaxf(fmaf((o[0] - mean) * rstd, ga.x, be.x), 0.f); o[1] = fmaxf(fmaf((o[1] - mean) * rstd, ga.y, be.y), 0.f);
              o[2] = fmaxf(fmaf((o[2] - mean) * rstd, ga.z, be.z), 0.f); o[3] = fmaxf(fmaf((o[3] - mean) * rstd, ga.w, be.w), 0.f);
            }
            if (g.vec_c && n + 3 < g.N) {
              float4* dst = reinterpret_cast<float4*>(crow + n);
              if (g.accumulate) { const float4 p = *dst; o[0] += p.x; o[1] += p.y; o[2] += p.z; o[3] += p.w; }
              *dst = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < g.N) crow[n + e] = g.accumulate ? crow[n + e] + o[e] : o[e];
            }
          }
        }
      }
      fence_before_sync();
      mbar_arrive(bar + PB_ACC_FREE + ab);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool tc_gemm_can_fuse_ln(int N, bool accumulate) { return N <= NT_MAX && N % 16 == 0 && !accumulate; }

long long tc_gemm_w_img_bytes(int N, const int* seg_k, int n_segs, bool split3) {
  const int KC = split3 ? Cfg<3>::KC : Cfg<2>::KC;
  long long chunks = 0;
  for (int s = 0; s < n_segs; ++s) chunks += (seg_k[s] + KC - 1) / KC;
  return (long long)((N + NT_MAX - 1) / NT_MAX) * chunks * (split3 ? 3 * Cfg<3>::W_BYTES : 2 * Cfg<2>::W_BYTES);
}

int launch_tc_gemm(const TcGemmArgs& g_in, int n_batch, cudaStream_t st) {
  if (g_in.M <= 0 || g_in.N <= 0 || n_batch <= 0) return 0;
  if (g_in.n_segs < 1 || g_in.n_segs > 4) { set_error_msg("tc_gemm: 1..4 operand segments"); return SMB_E_BADARG; }
  TcGemmArgs g = g_in;
  if ((g.ln_gamma != nullptr) != (g.ln_beta != nullptr) || (g.ln_gamma && !tc_gemm_can_fuse_ln(g.N, g.accumulate != 0))) {
    set_error_msg("tc_gemm: the fused LayerNorm needs gamma and beta, N <= 256, N % 16 == 0 and no accumulate");
    return SMB_E_BADARG;
  }
  // 16-byte vector loads / stores only where every row start is 16-byte aligned
  bool vec = aligned16(g.W) && g.ldw % 4 == 0 && g.w_batch % 4 == 0;
  for (int s = 0; s < g.n_segs; ++s)
    vec = vec && aligned16(g.seg[s].a) && g.seg[s].lda % 4 == 0 && g.seg[s].w_off % 4 == 0 && g.a_batch % 4 == 0;
  g.vec = vec ? 1 : 0;
  g.vec_c = (aligned16(g.C) && g.ldc % 4 == 0 && g.c_batch % 4 == 0 && (!g.bias || aligned16(g.bias))) ? 1 : 0;
  const dim3 grid((unsigned)((g.M + TM - 1) / TM), (unsigned)((g.N + NT_MAX - 1) / NT_MAX), (unsigned)n_batch);
  {
    int seg_k[4];
    for (int s = 0; s < g.n_segs; ++s) seg_k[s] = g.seg[s].k;
    const long long need = tc_gemm_w_img_bytes(g.N, seg_k, g.n_segs, g.split3 != 0);
    g.use_img = (g.w_img && g.w_batch == 0 && need <= g.w_img_bytes && (reinterpret_cast<uintptr_t>(g.w_img) & 127) == 0) ? 1 : 0;
    if (g.use_img) {
      const int KC = g.split3 ? Cfg<3>::KC : Cfg<2>::KC;
      int n_chunks = 0;
      for (int s = 0; s < g.n_segs; ++s) n_chunks += (g.seg[s].k + KC - 1) / KC;
      const dim3 sgrid((unsigned)n_chunks, grid.y);
      if (g.split3) tc_wsplit_kernel<3><<<sgrid, 256, 0, st>>>(g, n_chunks);
      else tc_wsplit_kernel<2><<<sgrid, 256, 0, st>>>(g, n_chunks);
    }
  }
  if (g.use_img && n_batch == 1 && g.N <= NT_MAX && g.M >= TM * device_sm_count()) {
    // big GEMM with shared weights: persistent CTAs, epilogue of a tile under the main loop of the next
    const int n_row_tiles = (g.M + TM - 1) / TM;
    const int pgrid = n_row_tiles < device_sm_count() ? n_row_tiles : device_sm_count();
    if (g.split3) {
      static size_t configured_p[kMaxDevices] = {};
      if (int rc = ensure_dynamic_smem(tc_gemm_persistent_kernel<3>, PCfg<3>::SMEM_TOTAL, configured_p)) return rc;
      tc_gemm_persistent_kernel<3><<<pgrid, P_THREADS, PCfg<3>::SMEM_TOTAL, st>>>(g);
    } else {
      static size_t configured_p[kMaxDevices] = {};
      if (int rc = ensure_dynamic_smem(tc_gemm_persistent_kernel<2>, PCfg<2>::SMEM_TOTAL, configured_p)) return rc;
      tc_gemm_persistent_kernel<2><<<pgrid, P_THREADS, PCfg<2>::SMEM_TOTAL, st>>>(g);
    }
    return (int)cudaGetLastError();
  }
  if (g.split3) {
    static size_t configured[kMaxDevices] = {};
    static size_t configured_i[kMaxDevices] = {};
    if (g.use_img) {
      if (int rc = ensure_dynamic_smem(tc_gemm_kernel<3, true>, Cfg<3>::SMEM_TOTAL, configured_i)) return rc;
      tc_gemm_kernel<3, true><<<grid, THREADS, Cfg<3>::SMEM_TOTAL, st>>>(g);
    } else {
      if (int rc = ensure_dynamic_smem(tc_gemm_kernel<3, false>, Cfg<3>::SMEM_TOTAL, configured)) return rc;
      tc_gemm_kernel<3, false><<<grid, THREADS, Cfg<3>::SMEM_TOTAL, st>>>(g);
    }
  } else {
    static size_t configured[kMaxDevices] = {};
    static size_t configured_i[kMaxDevices] = {};
    if (g.use_img) {
      if (int rc = ensure_dynamic_smem(tc_gemm_kernel<2, true>, Cfg<2>::SMEM_TOTAL, configured_i)) return rc;
      tc_gemm_kernel<2, true><<<grid, THREADS, Cfg<2>::SMEM_TOTAL, st>>>(g);
    } else {
      if (int rc = ensure_dynamic_smem(tc_gemm_kernel<2, false>, Cfg<2>::SMEM_TOTAL, configured)) return rc;
      tc_gemm_kernel<2, false><<<grid, THREADS, Cfg<2>::SMEM_TOTAL, st>>>(g);
    }
  }
  return (int)cudaGetLastError();
}

}  // namespace smb
