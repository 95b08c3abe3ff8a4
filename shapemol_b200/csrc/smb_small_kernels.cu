// Bandwidth-bound glue kernels of the denoising step (fp32 SIMT):
//   prep      : per-molecule time embedding + invariant shape embedding
//   knn       : batched fixed-k kNN, dense [N,k+1] output, no atomics
//   embed     : atom embedding h0
//   bn_final  : deterministic BatchNorm statistics reduction (+ running-stat update)
//   vn_apply  : VN batch-norm / leaky-ReLU epilogue and coordinate update
//   posterior : position posterior + log-categorical posterior + Gumbel-argmax (+ Philox noise)
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

// -------------------------------------------------------------------------------------------
// prep: one warp per molecule.
//   tau  = Linear(16->8)(SiLU(Linear(8->16)([sin(t w) | cos(t w)])))   (molopt_score_model.py:159-166,247-252)
//   inv  = MLP32->32->32( shape . m/(|m|^2+1e-6) ),  m = mean_c shape   (uni_transformer.py:181-189)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prep_kernel(PrepArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= a.n_mols) return;
  // ---- time embedding (lanes 0..15 hold the hidden layer) ----
  {
    const float t = (float)a.t[m];
    float emb = 0.f;
    if (lane < 8) {
      const float ph = t * a.time_freq[lane & 3];
      emb = lane < 4 ? sinf(ph) : cosf(ph);
    }
    float hid = 0.f;
    if (lane < 16) hid = a.time_b1[lane];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float e = __shfl_sync(0xffffffffu, emb, k);
      if (lane < 16) hid = fmaf(a.time_w1[lane * 8 + k], e, hid);
    }
    hid = hid / (1.f + expf(-hid));   // SiLU
    float out = lane < 8 ? a.time_b2[lane] : 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float e = __shfl_sync(0xffffffffu, hid, k);
      if (lane < 8) out = fmaf(a.time_w2[lane * 16 + k], e, out);
    }
    if (lane < 8) a.tau[m * 8 + lane] = out;
  }
  // ---- invariant shape embedding (lane = channel) ----
  if (a.do_shape) {
    const float* s = a.shape + (size_t)m * kShape * 3 + lane * 3;
    const float sx = s[0], sy = s[1], sz = s[2];
    const float inv32 = 1.f / 32.f;
    const float mx = warp_sum(sx) * inv32, my = warp_sum(sy) * inv32, mz = warp_sum(sz) * inv32;
    const float den = (mx * mx + my * my + mz * mz) + 1e-6f;
    const float q = sx * (mx / den) + sy * (my / den) + sz * (mz / den);
    float y = a.inv_b1[lane];
#pragma unroll 8
    for (int c = 0; c < 32; ++c) y = fmaf(a.inv_w1[lane * 32 + c], __shfl_sync(0xffffffffu, q, c), y);
    const float mean = warp_sum(y) * inv32;
    const float dlt = y - mean;
    const float var = warp_sum(dlt * dlt) * inv32;
    float z = dlt * (1.f / sqrtf(var + 1e-5f)) * a.inv_g[lane] + a.inv_bb[lane];
    z = fmaxf(z, 0.f);
    float o = a.inv_b2[lane];
#pragma unroll 8
    for (int c = 0; c < 32; ++c) o = fmaf(a.inv_w2[lane * 32 + c], __shfl_sync(0xffffffffu, z, c), o);
    a.inv[m * kShape + lane] = o;
    // ---- shape part of the VN linear maps: lane = (feat | dir, channel), three components each ----
    if (a.vn_shape) {
      const int which = lane >> 4, ch = lane & 15;
      for (int l = 0; l < a.n_layers; ++l) {
        const float* w = a.vn_w[l][which] + ch * kVnStride + 1 + kHeads;
        float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll 8
        for (int c = 0; c < kShape; ++c) {
          const float wc = __ldg(w + c);
          vx = fmaf(wc, __shfl_sync(0xffffffffu, sx, c), vx);
          vy = fmaf(wc, __shfl_sync(0xffffffffu, sy, c), vy);
          vz = fmaf(wc, __shfl_sync(0xffffffffu, sz, c), vz);
        }
        float* o3 = a.vn_shape + ((size_t)l * a.n_mols + m) * 96 + lane * 3;
        o3[0] = vx; o3[1] = vy; o3[2] = vz;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// knn: one 64-thread CTA per molecule (n <= 64).  Canonical distance:
// d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)), no FMA contraction; order key (d2, index); the first
// min(k+1, n) entries are kept, then self is dropped (torch_cluster.knn + knn_graph(loop=False)).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) knn_kernel(const float* __restrict__ x, const int* __restrict__ mol_ptr,
                                                 int n_mols, int k, int* __restrict__ nbr, int* __restrict__ deg) {
  __shared__ float xs[SMB_MAX_ATOMS_PER_MOL][3];
  __shared__ float d2[SMB_MAX_ATOMS_PER_MOL][SMB_MAX_ATOMS_PER_MOL + 1];
  const int m = blockIdx.x;
  if (m >= n_mols) return;
  const int a0 = mol_ptr[m], n = mol_ptr[m + 1] - a0, i = threadIdx.x;
  if (i < n) { xs[i][0] = x[(a0 + i) * 3]; xs[i][1] = x[(a0 + i) * 3 + 1]; xs[i][2] = x[(a0 + i) * 3 + 2]; }
  __syncthreads();
  if (i >= n) return;
  const float xi = xs[i][0], yi = xs[i][1], zi = xs[i][2];
  for (int j = 0; j < n; ++j) {
    const float dx = __fsub_rn(xi, xs[j][0]), dy = __fsub_rn(yi, xs[j][1]), dz = __fsub_rn(zi, xs[j][2]);
    d2[i][j] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  }
  const int KS = k + 1;
  const int keep = KS < n ? KS : n;
  // rank of self
  const float dself = d2[i][i];
  int rank_self = 0;
  for (int j = 0; j < n; ++j) { const float v = d2[i][j]; rank_self += (v < dself) || (v == dself && j < i); }
  int* out = nbr + (size_t)(a0 + i) * KS;
  int cnt = 0;
  for (int j = 0; j < n; ++j) {
    if (j == i) continue;
    const float dj = d2[i][j];
    int rank = 0;
    for (int jj = 0; jj < n; ++jj) { const float v = d2[i][jj]; rank += (v < dj) || (v == dj && jj < j); }
    if (rank < keep) { out[rank - (rank_self < rank ? 1 : 0)] = j; ++cnt; }
  }
  for (int s = cnt; s < KS; ++s) out[s] = -1;
  deg[a0 + i] = cnt;
}

// -------------------------------------------------------------------------------------------
// embed: h0[i] = W_emb [onehot(v_i) | tau_mol(i)] + b   (molopt_score_model.py:292-301)
// one thread per (atom, 4 channels)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_kernel(EmbedArgs a) {
  const int H4 = a.H / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.n_atoms * H4) return;
  const int i = (int)(idx / H4), c4 = (int)(idx % H4);
  const int v = a.v[i], m = a.atom_mol[i];
  const float4* wT = reinterpret_cast<const float4*>(a.emb_wT);
  float4 acc = reinterpret_cast<const float4*>(a.emb_b)[c4];
  const float4 w = wT[(size_t)v * H4 + c4];
  acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float tk = a.tau[m * 8 + k];
    const float4 wk = wT[(size_t)(a.classes + k) * H4 + c4];
    acc.x = fmaf(wk.x, tk, acc.x); acc.y = fmaf(wk.y, tk, acc.y);
    acc.z = fmaf(wk.z, tk, acc.z); acc.w = fmaf(wk.w, tk, acc.w);
  }
  reinterpret_cast<float4*>(a.h)[idx] = acc;
  if (a.h0) reinterpret_cast<float4*>(a.h0)[idx] = acc;
}

// -------------------------------------------------------------------------------------------
// bn_final: one CTA.  Reduces the per-warp partial sums (sum nu, sum nu^2 per channel) in a fixed
// order in fp64, produces scale/shift of BatchNorm1d(16) (shape_vn_layers.py:41-61) and, in
// training mode, updates the running statistics like nn.BatchNorm1d (momentum 0.1, unbiased var).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_final_kernel(BnArgs a) {
  __shared__ double red[8][32];
  const int c = threadIdx.x & 31, part = threadIdx.x >> 5;
  if (a.training) {
    double s = 0.0;
    for (int r = part; r < a.rows; r += 8) s += (double)a.partial[(size_t)r * 32 + c];
    red[part][c] = s;
    __syncthreads();
    if (part == 0) {
      double tot = 0.0;
      for (int p = 0; p < 8; ++p) tot += red[p][c];
      red[0][c] = tot;
    }
    __syncthreads();
  }
  if (threadIdx.x < kHeads) {
    const int ch = threadIdx.x;
    float mean, var;
    if (a.training) {
      const double n = (double)a.n_atoms;
      const double mu = red[0][ch] / n;
      double v = red[0][16 + ch] / n - mu * mu;
      if (v < 0.0) v = 0.0;
      mean = (float)mu; var = (float)v;
      const double unb = a.n_atoms > 1 ? v * n / (n - 1.0) : v;
      a.running_mean[ch] = 0.9f * a.running_mean[ch] + 0.1f * mean;
      a.running_var[ch] = 0.9f * a.running_var[ch] + 0.1f * (float)unb;
      if (ch == 0 && a.num_batches_tracked) a.num_batches_tracked[0] += 1;
    } else {
      mean = a.running_mean[ch]; var = a.running_var[ch];
    }
    const float scale = a.weight[ch] / sqrtf(var + 1e-5f);
    a.param[ch] = scale;
    a.param[16 + ch] = a.bias[ch] - mean * scale;
  }
}

// -------------------------------------------------------------------------------------------
// vn_apply: per atom, finish VNLinearLeakyReLU (shape_vn_layers.py:100-110) and update x
// (uni_transformer.py:155-156,326):  p <- p/nu * BN(nu);  leaky-ReLU along d;  x += mean_a o + mean_a p.
// 16 lanes per atom (lane = channel).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vn_apply_kernel(const float* __restrict__ vn, const float* __restrict__ bn_param,
                                                      float* __restrict__ x, float* __restrict__ x_out, int n_atoms) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gid >> 4, ch = gid & 15;
  const bool live = i < n_atoms;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (live) {
    const float* row = vn + (size_t)i * kVnRow;
    float px = row[3 + ch * 3], py = row[4 + ch * 3], pz = row[5 + ch * 3];
    const float dx = row[51 + ch * 3], dy = row[52 + ch * 3], dz = row[53 + ch * 3];
    const float nu = sqrtf(px * px + py * py + pz * pz) + 1e-6f;
    const float nbn = nu * bn_param[ch] + bn_param[16 + ch];
    const float sc = nbn / nu;
    px *= sc; py *= sc; pz *= sc;
    const float dot = px * dx + py * dy + pz * dz;
    const float dn = dx * dx + dy * dy + dz * dz;
    if (!(dot >= 0.f)) {
      const float f = dot / (dn + 1e-6f);
      const float qx = px - f * dx, qy = py - f * dy, qz = pz - f * dz;
      rx = 0.2f * px + 0.8f * qx; ry = 0.2f * py + 0.8f * qy; rz = 0.2f * pz + 0.8f * qz;
    } else {
      rx = px; ry = py; rz = pz;   // 0.2 p + 0.8 p
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    rx += __shfl_xor_sync(0xffffffffu, rx, o);
    ry += __shfl_xor_sync(0xffffffffu, ry, o);
    rz += __shfl_xor_sync(0xffffffffu, rz, o);
  }
  if (live && ch < 3) {
    const float* row = vn + (size_t)i * kVnRow;
    const float r = ch == 0 ? rx : (ch == 1 ? ry : rz);
    const float nx = x[i * 3 + ch] + row[ch] + r * (1.f / 16.f);
    x[i * 3 + ch] = nx;
    if (x_out) x_out[i * 3 + ch] = nx;
  }
}

// -------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (own implementation; perf-mode noise)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }   // [0,1)

__device__ __forceinline__ float log_add_exp(float a, float b) {
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}

// posterior: one thread per atom (molopt_score_model.py:655-673).
__global__ void __launch_bounds__(128) posterior_kernel(PosteriorArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_atoms) return;
  const int C = a.classes;
  int t = a.t[a.atom_mol[i]];
  t = t < 0 ? 0 : (t >= a.timesteps ? a.timesteps - 1 : t);   // table bounds
  const int tm1 = t - 1 < 0 ? 0 : t - 1;
  // ---- noise ----
  float eps[3], u[16];
  if (a.noise_pos) {
    eps[0] = a.noise_pos[i * 3]; eps[1] = a.noise_pos[i * 3 + 1]; eps[2] = a.noise_pos[i * 3 + 2];
    for (int c = 0; c < C; ++c) u[c] = a.noise_u[(size_t)i * C + c];
  } else {
    const uint64_t gi = (uint64_t)(a.atom_offset + i);
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    uint4 r[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) r[q] = philox4x32(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)t, (uint32_t)q), key);
    // Box-Muller on r[0], r[1].xy
    const float u1 = 1.0f - u01(r[0].x), u2 = u01(r[0].y), u3 = 1.0f - u01(r[0].z), u4 = u01(r[0].w);
    const float ra = sqrtf(-2.f * logf(u1)), rb = sqrtf(-2.f * logf(u3));
    eps[0] = ra * cospif(2.f * u2); eps[1] = ra * sinpif(2.f * u2); eps[2] = rb * cospif(2.f * u4);
    const uint32_t* rr = reinterpret_cast<const uint32_t*>(&r[1]);
    for (int c = 0; c < 16; ++c) u[c] = u01(rr[c]);
  }
  // ---- position posterior ----
  const float c0 = a.c0[t], ct = a.ct[t];
  const float sig = t == 0 ? 0.f : expf(0.5f * a.logvar[t]);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // separately rounded products/sums, like the reference's elementwise torch ops (:402-403,662)
    const float mean = __fadd_rn(__fmul_rn(c0, a.pred_pos[i * 3 + d]), __fmul_rn(ct, a.pos[i * 3 + d]));
    a.pos[i * 3 + d] = __fadd_rn(mean, __fmul_rn(sig, eps[d]));
  }
  // ---- log-categorical posterior ----
  float lg[16];
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) { lg[c] = a.pred_v[(size_t)i * C + c]; mx = fmaxf(mx, lg[c]); }
  float se = 0.f;
  for (int c = 0; c < C; ++c) se += expf(lg[c] - mx);
  const float lse = mx + logf(se);
  const int vt = a.v[i];
  const float lnC = logf((float)C);
  const float lac = a.log_ac[tm1], l1mac = a.log_1m_ac[tm1] - lnC;
  const float la = a.log_a[t], l1ma = a.log_1m_a[t] - lnC;
  const float off = -69.07755279f;   // log(1e-30)
  float un[16];
  float m2 = -INFINITY;
  for (int c = 0; c < C; ++c) {
    const float lv0 = lg[c] - lse;
    if (a.log_v0) a.log_v0[(size_t)i * C + c] = lv0;
    const float A = log_add_exp(lv0 + lac, l1mac);
    const float Bq = log_add_exp((c == vt ? 0.f : off) + la, l1ma);
    un[c] = A + Bq;
    m2 = fmaxf(m2, un[c]);
  }
  float s2 = 0.f;
  for (int c = 0; c < C; ++c) s2 += expf(un[c] - m2);
  const float lse2 = m2 + logf(s2);
  int best = 0;
  float bestv = -INFINITY;
  for (int c = 0; c < C; ++c) {
    const float post = un[c] - lse2;
    if (a.log_post) a.log_post[(size_t)i * C + c] = post;
    const float gum = -logf(-logf(u[c] + 1e-30f) + 1e-30f);
    const float sc = gum + post;
    if (sc > bestv) { bestv = sc; best = c; }
  }
  a.v[i] = best;
}

__global__ void decrement_t_kernel(int* t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] -= 1;
}

// ---- launchers -------------------------------------------------------------------------------
int launch_prep(const PrepArgs& a, cudaStream_t st) {
  if (a.n_mols <= 0) return 0;
  prep_kernel<<<(a.n_mols + 3) / 4, 128, 0, st>>>(a);
  return (int)cudaGetLastError();
}
int launch_knn(const float* x, const int* mol_ptr, int n_mols, int k, int* nbr, int* deg, cudaStream_t st) {
  if (n_mols <= 0) return 0;
  knn_kernel<<<n_mols, 64, 0, st>>>(x, mol_ptr, n_mols, k, nbr, deg);
  return (int)cudaGetLastError();
}
int launch_embed(const EmbedArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 0) return 0;
  const size_t total = (size_t)a.n_atoms * (a.H / 4);
  embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  return (int)cudaGetLastError();
}
int launch_bn_final(const BnArgs& a, cudaStream_t st) {
  bn_final_kernel<<<1, 256, 0, st>>>(a);
  return (int)cudaGetLastError();
}
int launch_vn_apply(const float* vn, const float* bn_param, float* x, float* x_out, int n_atoms, cudaStream_t st) {
  if (n_atoms <= 0) return 0;
  vn_apply_kernel<<<(n_atoms * 16 + 255) / 256, 256, 0, st>>>(vn, bn_param, x, x_out, n_atoms);
  return (int)cudaGetLastError();
}
int launch_posterior(const PosteriorArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 0) return 0;
  posterior_kernel<<<(a.n_atoms + 127) / 128, 128, 0, st>>>(a);
  return (int)cudaGetLastError();
}
// -------------------------------------------------------------------------------------------
// Point-cloud shape guidance: one thread per atom (models/molopt_score_model.py:699-740).  Float64 in the reference's
// (numpy / sklearn) evaluation order, every operation individually rounded: no FMA contraction.
// -------------------------------------------------------------------------------------------
struct Nn3 { double d[3]; int i[3]; };

__device__ __forceinline__ void three_nn(const double* __restrict__ cloud, int c0, int c1, double px, double py, double pz, Nn3& r) {
  r.d[0] = r.d[1] = r.d[2] = INFINITY;
  r.i[0] = r.i[1] = r.i[2] = c0;
  for (int c = c0; c < c1; ++c) {
    const double dx = __dsub_rn(px, cloud[3 * c]), dy = __dsub_rn(py, cloud[3 * c + 1]), dz = __dsub_rn(pz, cloud[3 * c + 2]);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 < r.d[2]) {                       // strict: the lower index wins ties
      if (d2 < r.d[1]) {
        r.d[2] = r.d[1]; r.i[2] = r.i[1];
        if (d2 < r.d[0]) { r.d[1] = r.d[0]; r.i[1] = r.i[0]; r.d[0] = d2; r.i[0] = c; }
        else { r.d[1] = d2; r.i[1] = c; }
      } else { r.d[2] = d2; r.i[2] = c; }
    }
  }
}
__device__ __forceinline__ double mean3(double a, double b, double c) { return __ddiv_rn(__dadd_rn(__dadd_rn(a, b), c), 3.0); }

__global__ void __launch_bounds__(128) guidance_kernel(smb_guidance_io io, int n_atoms, const int* __restrict__ atom_mol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const int m = atom_mol ? atom_mol[i] : 0;
  int tt = io.step;
  if (io.t) { tt = io.t[m]; if (tt <= io.grad_step) return; }
  const int c0 = io.cloud_ptr ? io.cloud_ptr[m] : 0, c1 = io.cloud_ptr ? io.cloud_ptr[m + 1] : io.n_cloud;
  if (c1 - c0 < 3) return;
  double px = (double)io.pos[3 * i], py = (double)io.pos[3 * i + 1], pz = (double)io.pos[3 * i + 2];
  Nn3 nn;
  three_nn(io.cloud, c0, c1, px, py, pz, nn);
  if (!(mean3(sqrt(nn.d[0]), sqrt(nn.d[1]), sqrt(nn.d[2])) > io.radius)) return;
  const double span = __dsub_rn(0.8, io.ratio);
  const uint64_t gi = (uint64_t)(io.atom_offset + i);
  const uint2 key = make_uint2((uint32_t)io.seed, (uint32_t)(io.seed >> 32));
  for (int j = 0; j < 5; ++j) {
    double uj;
    if (io.u) uj = io.u[(size_t)j * n_atoms + i];
    else {   // 53 random bits -> [0, 1), as numpy's random_sample
      const uint4 r = philox4x32(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)tt, 0x47554944u + (uint32_t)j), key);
      uj = (double)(((uint64_t)(r.x >> 5) << 26) | (uint64_t)(r.y >> 6)) * (1.0 / 9007199254740992.0);
    }
    const double* a = io.cloud + 3 * (size_t)nn.i[0];
    const double* b = io.cloud + 3 * (size_t)nn.i[1];
    const double* c = io.cloud + 3 * (size_t)nn.i[2];
    const double s = __dadd_rn(__dmul_rn(uj, span), io.ratio);
    px = __dsub_rn(px, __dmul_rn(s, __dsub_rn(px, mean3(a[0], b[0], c[0]))));
    py = __dsub_rn(py, __dmul_rn(s, __dsub_rn(py, mean3(a[1], b[1], c[1]))));
    pz = __dsub_rn(pz, __dmul_rn(s, __dsub_rn(pz, mean3(a[2], b[2], c[2]))));
    three_nn(io.cloud, c0, c1, px, py, pz, nn);
    if (mean3(sqrt(nn.d[0]), sqrt(nn.d[1]), sqrt(nn.d[2])) < io.radius) break;
  }
  io.pos[3 * i] = (float)px; io.pos[3 * i + 1] = (float)py; io.pos[3 * i + 2] = (float)pz;
}

int launch_guidance(const smb_guidance_io& io, int n_atoms, const int* atom_mol, cudaStream_t st) {
  if (n_atoms <= 0) return 0;
  guidance_kernel<<<(n_atoms + 127) / 128, 128, 0, st>>>(io, n_atoms, atom_mol);
  return (int)cudaGetLastError();
}
// -------------------------------------------------------------------------------------------
// Shape Tanimoto of get_ROCS (utils/evaluation/shaep_utils.py:59-83): one warp per molecule, float64.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ double overlap_sum(const double* a, int na, const double* b, int nb, double k, double coef, double den, int lane) {
  double acc = 0.0;
  for (int p = lane; p < na * nb; p += 32) {
    const int i = p / nb, j = p - i * nb;
    const double dx = a[3 * i] - b[3 * j], dy = a[3 * i + 1] - b[3 * j + 1], dz = a[3 * i + 2] - b[3 * j + 2];
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    acc += coef * exp(-k * (r * r)) / den;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(128) tanimoto_kernel(const float* __restrict__ pos, const int* __restrict__ mol_ptr, int n_mols,
                                                       const double* __restrict__ ref, const int* __restrict__ ref_ptr, int n_ref, double k,
                                                       double coef, double den, double* __restrict__ out) {
  __shared__ double s_a[4][SMB_MAX_ATOMS_PER_MOL * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= n_mols) return;
  const int a0 = mol_ptr[m], na = mol_ptr[m + 1] - a0;
  const int r0 = ref_ptr ? ref_ptr[m] : 0, nr = ref_ptr ? ref_ptr[m + 1] - r0 : n_ref;
  for (int p = lane; p < na * 3; p += 32) s_a[warp][p] = (double)pos[(size_t)a0 * 3 + p];
  __syncwarp();
  const double* b = ref + (size_t)r0 * 3;
  const double vaa = overlap_sum(s_a[warp], na, s_a[warp], na, k, coef, den, lane);
  const double vbb = overlap_sum(b, nr, b, nr, k, coef, den, lane);
  const double vab = overlap_sum(s_a[warp], na, b, nr, k, coef, den, lane);
  if (lane == 0) out[m] = vab / (vaa + vbb - vab);
}

int launch_tanimoto(const float* pos, const int* mol_ptr, int n_mols, const double* ref, const int* ref_ptr, int n_ref, double k,
                    double coef, double den, double* out, cudaStream_t st) {
  if (n_mols <= 0) return 0;
  tanimoto_kernel<<<(n_mols + 3) / 4, 128, 0, st>>>(pos, mol_ptr, n_mols, ref, ref_ptr, n_ref, k, coef, den, out);
  return (int)cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// check_stability / get_bond_order (utils/evaluation/analyze.py:249-297): one warp per molecule, lane = atom (two per lane
// for molecules of more than 32 atoms).  float32 distances in numpy's evaluation order, every operation individually rounded.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stability_kernel(const float* __restrict__ pos, const int* __restrict__ mol_ptr, int n_mols,
                                                        const int* __restrict__ elem, const int* __restrict__ thr,
                                                        const int* __restrict__ allowed, int n_elem, int hs, int* __restrict__ nr_bonds,
                                                        int* __restrict__ stable_atoms) {
  __shared__ float s_x[4][SMB_MAX_ATOMS_PER_MOL][3];
  __shared__ int s_e[4][SMB_MAX_ATOMS_PER_MOL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= n_mols) return;
  const int a0 = mol_ptr[m], n = mol_ptr[m + 1] - a0;
  for (int i = lane; i < n; i += 32) {
    s_x[warp][i][0] = pos[(size_t)(a0 + i) * 3]; s_x[warp][i][1] = pos[(size_t)(a0 + i) * 3 + 1]; s_x[warp][i][2] = pos[(size_t)(a0 + i) * 3 + 2];
    s_e[warp][i] = elem[a0 + i];
  }
  __syncwarp();
  const int ee = n_elem * n_elem;
  int stable = 0;
  for (int i = lane; i < n; i += 32) {
    const float xi = s_x[warp][i][0], yi = s_x[warp][i][1], zi = s_x[warp][i][2];
    const int ei = s_e[warp][i];
    int nb = 0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const float dx = __fsub_rn(xi, s_x[warp][j][0]), dy = __fsub_rn(yi, s_x[warp][j][1]), dz = __fsub_rn(zi, s_x[warp][j][2]);
      const float d = __fmul_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 100.f);
      const int* t = thr + ei * n_elem + s_e[warp][j];
      if (d < (float)t[0]) { nb += 1; if (d < (float)t[ee]) { nb += 1; if (d < (float)t[2 * ee]) nb += 1; } }
    }
    if (nr_bonds) nr_bonds[a0 + i] = nb;
    const int al = allowed[ei];
    stable += hs ? (al == nb) : (al >= nb && nb > 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) stable += __shfl_xor_sync(0xffffffffu, stable, o);
  if (lane == 0) stable_atoms[m] = stable;
}

int launch_stability(const float* pos, const int* mol_ptr, int n_mols, const int* elem, const int* thr, const int* allowed, int n_elem,
                     int hs, int* nr_bonds, int* stable_atoms, cudaStream_t st) {
  if (n_mols <= 0) return 0;
  stability_kernel<<<(n_mols + 3) / 4, 128, 0, st>>>(pos, mol_ptr, n_mols, elem, thr, allowed, n_elem, hs, nr_bonds, stable_atoms);
  return (int)cudaGetLastError();
}

int launch_decrement_t(int* t, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  decrement_t_kernel<<<(n + 127) / 128, 128, 0, st>>>(t, n);
  return (int)cudaGetLastError();
}

}  // namespace smb
