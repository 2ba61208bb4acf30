// Hardware probe (not part of the product): validates tcgen05 operand conventions that the
// attention / backward kernels rely on, against a host reference, on a real B200:
//   T0  SS, A and B K-major, SWIZZLE_64B            (sanity: same as gemm_tc.cu)
//   T1  SS, B MN-major (B stored (K, N) row-major), SWIZZLE_64B boxes [64 k x 32 n]
//   T2  TS, A read from TMEM (packed bf16 pairs written with tcgen05.st), B K-major
//   T3  SS, A MN-major (A stored (K, M) row-major), B MN-major
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 probe_umma.cu
//        ../../lightningfastspeech2_b200/csrc/gemm_tc.cu ../../lightningfastspeech2_b200/csrc/elementwise.cu -o probe_umma
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../lightningfastspeech2_b200/csrc/tc_common.cuh"

using namespace lfs2;
using namespace lfs2::tc;

constexpr int M = 128, N = 128, K = 64;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct ProbeParams {
  int mode;
  uint32_t lbo, sbo, kstep;  // MN-major descriptor fields (bytes) and per-k16 start advance (bytes)
  const __nv_bfloat16* a_glob;  // (M, K) row-major, for TS mode
  float* c;                      // (M, N)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sA = smem;            // 16 KB
  uint8_t* sB = smem + 16384;    // 16 KB
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const bool a_mn = p.mode == 3, b_mn = p.mode == 1 || p.mode == 3, ts = p.mode == 2;

  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (!ts) bytes += 16384;
    bytes += 16384;
    mbar_expect_tx(&full_bar, bytes);
    if (!ts) {
      if (!a_mn) {
        for (int s = 0; s < K / 32; ++s) tma_load_3d(sA + s * 8192, &map_a, &full_bar, s * 32, 0, 0);  // [128 m x 32 k]
      } else {
        for (int j = 0; j < M / 32; ++j) tma_load_3d(sA + j * 4096, &map_a, &full_bar, j * 32, 0, 0);  // [64 k x 32 m]
      }
    }
    if (!b_mn) {
      for (int s = 0; s < K / 32; ++s) tma_load_3d(sB + s * 8192, &map_b, &full_bar, s * 32, 0, 0);    // [128 n x 32 k]
    } else {
      for (int j = 0; j < N / 32; ++j) tma_load_3d(sB + j * 4096, &map_b, &full_bar, j * 32, 0, 0);    // [64 k x 32 n]
    }
  }
  if (ts) {
    // thread r = row r: 64 bf16 = 32 packed words -> TMEM columns [128, 160)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.a_glob + (size_t)threadIdx.x * K);
    float v[32];
    uint32_t* vi = reinterpret_cast<uint32_t*>(v);
    for (int j = 0; j < 32; ++j) vi[j] = src[j];
    tmem_st32(tmem_base + 128 + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    uint32_t idesc = make_idesc(kFmtBF16, M, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint64_t ad, bd;
      if (!a_mn) ad = make_smem_desc(smem_u32(sA) + (k16 / 2) * 8192 + (k16 % 2) * 32, 16, 512, kSwizzle64);
      else ad = make_smem_desc(smem_u32(sA) + k16 * p.kstep, p.lbo, p.sbo, kSwizzle64);
      if (!b_mn) bd = make_smem_desc(smem_u32(sB) + (k16 / 2) * 8192 + (k16 % 2) * 32, 16, 512, kSwizzle64);
      else bd = make_smem_desc(smem_u32(sB) + k16 * p.kstep, p.lbo, p.sbo, kSwizzle64);
      if (ts) umma_f16_ts(tmem_base, tmem_base + 128 + k16 * 8, bd, idesc, k16 ? 1u : 0u);
      else umma_f16(tmem_base, ad, bd, idesc, k16 ? 1u : 0u);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  float v[32];
  for (int c = 0; c < N / 32; ++c) {
    tmem_ld32(tmem_base + c * 32 + ((uint32_t)(warp * 32) << 16), v);
    for (int j = 0; j < 32; ++j) p.c[(size_t)threadIdx.x * N + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  std::vector<float> A(M * K), B(N * K), C(M * N);
  srand(1);
  for (auto& x : A) x = bf((rand() % 17 - 8) / 8.0f);
  for (auto& x : B) x = bf((rand() % 13 - 6) / 4.0f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      C[m * N + n] = s;
    }
  std::vector<__nv_bfloat16> a_mk(M * K), a_km(K * M), b_nk(N * K), b_kn(K * N);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      a_mk[m * K + k] = __float2bfloat16_rn(A[m * K + k]);
      a_km[k * M + m] = a_mk[m * K + k];
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      b_nk[n * K + k] = __float2bfloat16_rn(B[n * K + k]);
      b_kn[k * N + n] = b_nk[n * K + k];
    }
  __nv_bfloat16 *d_amk, *d_akm, *d_bnk, *d_bkn;
  float* d_c;
  cudaMalloc(&d_amk, M * K * 2); cudaMalloc(&d_akm, M * K * 2); cudaMalloc(&d_bnk, N * K * 2); cudaMalloc(&d_bkn, N * K * 2);
  cudaMalloc(&d_c, M * N * 4);
  cudaMemcpy(d_amk, a_mk.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_akm, a_km.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_bnk, b_nk.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_bkn, b_kn.data(), N * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap m_amk, m_akm, m_bnk, m_bkn;
  bool ok = make_tmap_3d(&m_amk, d_amk, K, M, 1, 32, 128, 64) && make_tmap_3d(&m_bnk, d_bnk, K, N, 1, 32, 128, 64) &&
            make_tmap_3d(&m_akm, d_akm, M, K, 1, 32, 64, 64) && make_tmap_3d(&m_bkn, d_bkn, N, K, 1, 32, 64, 64);
  if (!ok) { printf("tensor map creation failed\n"); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  struct Cand { int mode; uint32_t lbo, sbo, kstep; const char* name; };
  Cand cands[] = {
      {0, 0, 0, 0, "T0 SS K-major/K-major"},
      {1, 4096, 512, 1024, "T1 B MN-major lbo=4096 sbo=512 kstep=1024"},
      {1, 512, 4096, 1024, "T1 B MN-major lbo=512 sbo=4096 kstep=1024"},
      {2, 0, 0, 0, "T2 TS A-from-TMEM (packed pairs, 8 cols per k16)"},
      {3, 4096, 512, 1024, "T3 A,B MN-major lbo=4096 sbo=512 kstep=1024"},
      {3, 512, 4096, 1024, "T3 A,B MN-major lbo=512 sbo=4096 kstep=1024"},
  };
  std::vector<float> out(M * N);
  int ci = -1;
  for (auto& cd : cands) {
    if (++ci != only && only >= 0) continue;
    cudaMemset(d_c, 0xff, M * N * 4);
    ProbeParams p{cd.mode, cd.lbo, cd.sbo, cd.kstep, d_amk, d_c};
    const CUtensorMap& ma = cd.mode == 3 ? m_akm : m_amk;
    const CUtensorMap& mb = (cd.mode == 1 || cd.mode == 3) ? m_bkn : m_bnk;
    probe_kernel<<<1, 128, 40960>>>(ma, mb, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", cd.name, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), d_c, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) {
      double d = fabs((double)out[i] - C[i]);
      if (!(d <= 1e-3)) ++bad;
      if (d > maxerr || d != d) maxerr = d;
    }
    printf("%-55s maxerr=%.4g bad=%d/%d %s\n", cd.name, maxerr, bad, M * N, bad == 0 ? "PASS" : "FAIL");
  }
  return 0;
}
