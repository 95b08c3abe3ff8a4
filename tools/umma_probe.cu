// Hardware probe for the tcgen05 building blocks the edge kernels rely on (sm_100a).  Not part of the
// product: it pins down, on a real B200, the shared-memory descriptor layouts (no-swizzle K-major and
// MN-major), the A-from-TMEM packing, and the tcgen05.ld/st register<->TMEM mapping.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  return d;          // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}

struct Params {
  int K, N;
  int a_tmem;      // 1: A operand written to TMEM with tcgen05.st (packed bf16x2) and used by the TS form
  int b_mn;        // 1: B operand is MN-major
  int a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;   // bytes
  const uint8_t* a_img; int a_bytes;
  const uint8_t* b_img; int b_bytes;
  const __nv_bfloat16* a_rows;   // [128][K] row-major (for the TMEM path)
  float* d;                      // [128][N]
};

__global__ void __launch_bounds__(128, 1) probe(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(p.a_img)[i];
  for (int i = tid; i < p.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(p.b_img)[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t d_tmem = tmem;          // columns [0, N)
  const uint32_t a_tm = tmem + 256;      // columns [256, 256 + K/2)

  if (p.a_tmem) {
    // thread = row: pack (k, k+1) into one 32-bit column, k in the low half
    const __nv_bfloat16* row = p.a_rows + (size_t)tid * p.K;
    for (int c0 = 0; c0 < p.K / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 t; t.x = row[2 * (c0 + j)]; t.y = row[2 * (c0 + j) + 1];
        v[j] = *reinterpret_cast<uint32_t*>(&t);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "r"(a_tm + lane_base + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.b_mn << 16) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < p.K / 16; ++ks) {
      const uint64_t bd = make_desc(smem_u32(sb) + ks * p.b_kstep, p.b_lbo, p.b_sbo);
      if (p.a_tmem) mma_ts(d_tmem, a_tm + ks * 8, bd, idesc, ks > 0);
      else mma_ss(d_tmem, make_desc(smem_u32(sa) + ks * p.a_kstep, p.a_lbo, p.a_sbo), bd, idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(d_tmem + lane_base + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) p.d[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// variant bit 0: swap the roles of LBO and SBO for A; bit 1: same for B
static double run(int K, int N, int a_tmem, int b_mn, int variant) {
  std::vector<float> A(128 * K), B((size_t)N * K);
  srand(1234 + K + N);
  for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = bf((rand() % 2001 - 1000) / 1000.f);
  const int LBO = 128, SBO_A = (K / 8) * 128, SBO_B = (K / 8) * 128;
  std::vector<uint8_t> ai((size_t)128 * K * 2), bi((size_t)N * K * 2);
  std::vector<__nv_bfloat16> arows(128 * K);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) {
    __nv_bfloat16 h = __float2bfloat16(A[r * K + k]);
    arows[r * K + k] = h;
    size_t off = (size_t)(r / 8) * SBO_A + (size_t)(k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(&ai[off]) = h;
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
    __nv_bfloat16 h = __float2bfloat16(B[(size_t)n * K + k]);
    size_t off;
    if (b_mn) off = (size_t)(n / 8) * SBO_B + (size_t)(k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2;
    else off = (size_t)(n / 8) * SBO_B + (size_t)(k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(&bi[off]) = h;
  }
  Params p{};
  p.K = K; p.N = N; p.a_tmem = a_tmem; p.b_mn = b_mn;
  p.a_lbo = (variant & 1) ? SBO_A : LBO; p.a_sbo = (variant & 1) ? LBO : SBO_A;
  p.b_lbo = (variant & 2) ? SBO_B : LBO; p.b_sbo = (variant & 2) ? LBO : SBO_B;
  p.a_kstep = 2 * LBO; p.b_kstep = 2 * LBO;
  uint8_t *da, *db; __nv_bfloat16* dr; float* dd;
  CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dr, arows.size() * 2)); CK(cudaMalloc(&dd, (size_t)128 * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr, arows.data(), arows.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, (size_t)128 * N * 4));
  p.a_img = da; p.a_bytes = (int)ai.size(); p.b_img = db; p.b_bytes = (int)bi.size(); p.a_rows = dr; p.d = dd;
  const int smem = ((p.a_bytes + 1023) / 1024) * 1024 + p.b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)128 * N);
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  double err = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
    double s = 0;
    for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * B[(size_t)n * K + k];
    err = fmax(err, fabs(s - D[(size_t)r * N + n]));
  }
  cudaFree(da); cudaFree(db); cudaFree(dr); cudaFree(dd);
  return err;
}

int main() {
  struct T { int K, N, a_tmem, b_mn; } tests[] = {
      {64, 128, 0, 0}, {64, 128, 0, 1}, {64, 128, 1, 0}, {128, 16, 1, 0}, {80, 256, 0, 1}, {128, 128, 1, 0}, {32, 64, 0, 0}};
  for (auto& t : tests)
    for (int v = 0; v < 4; ++v) {
      if (t.a_tmem && (v & 1)) continue;
      double e = run(t.K, t.N, t.a_tmem, t.b_mn, v);
      printf("K=%3d N=%3d a_tmem=%d b_mn=%d variant=%d  max_abs_err=%.3e %s\n", t.K, t.N, t.a_tmem, t.b_mn, v, e, e < 1e-3 ? "OK" : "");
      fflush(stdout);
    }
  return 0;
}
