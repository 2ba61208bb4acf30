// fp32 CUDA-core GEMM  c (m,n) = a (m,k) . w (n,k)^T + bias [, relu]   and its implicit-GEMM
// Conv1d form (k = ksize*d, A rows gathered from time-shifted frames with zero fill at the
// sequence ends).  128x128x16 tiles, 8x8 register micro-tiles, double-buffered smem.
//
// This is the exact-fp32 GEMM of the path: it is the on-device ground truth the tcgen05
// (bf16x3 split) GEMMs are checked against, and it serves the shapes those kernels do
// not cover.
#include "common.cuh"

namespace lfs2 {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
constexpr int kGemmThreads = 256;

struct ConvMap {
  int t;      // frames per utterance (0 => plain GEMM)
  int d;      // channels per tap
  int half;   // (ksize-1)/2
};

__global__ void __launch_bounds__(kGemmThreads)
gemm_f32_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ c, int m, int n, int k, int lda, int relu, ConvMap cm) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2;          // 0..63 (+64)
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int row = m0 + lrow + h * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < m) {
        if (cm.t == 0) {
          v = *reinterpret_cast<const float4*>(a + (size_t)row * lda + k0 + lk);
        } else {
          int tap = k0 / cm.d, c0 = k0 - tap * cm.d;
          int tt = row % cm.t + tap - cm.half;
          if (tt >= 0 && tt < cm.t)
            v = *reinterpret_cast<const float4*>(a + (size_t)(row + tap - cm.half) * cm.d + c0 + lk);
        }
      }
      ra[h] = v;
      int col = n0 + lrow + h * 64;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < n) u = *reinterpret_cast<const float4*>(w + (size_t)col * k + k0 + lk);
      rb[h] = u;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x;
      As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z;
      As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x;
      Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z;
      Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int nk = k / BK;
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= m) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int col = n0 + jh * 64 + tx * 4;
      if (col >= n) continue;  // n % 4 == 0
      float4 o;
      float4 bz = bias ? *reinterpret_cast<const float4*>(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      o.x = acc[i][jh * 4 + 0] + bz.x;
      o.y = acc[i][jh * 4 + 1] + bz.y;
      o.z = acc[i][jh * 4 + 2] + bz.z;
      o.w = acc[i][jh * 4 + 3] + bz.w;
      if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      *reinterpret_cast<float4*>(c + (size_t)row * n + col) = o;
    }
  }
}

static int launch_gemm(const float* a, const float* w, const float* bias, float* c, int m, int n, int k, int lda,
                       int relu, ConvMap cm, void* stream, const char* name) {
  dim3 grid(ceil_div(n, BN), ceil_div(m, BM));
  LFS2_REQUIRE(grid.y <= 65535, LFS2_ERR_UNSUPPORTED, "%s: m=%d too large", name, m);
  gemm_f32_kernel<<<grid, kGemmThreads, 0, (cudaStream_t)stream>>>(a, w, bias, c, m, n, k, lda, relu, cm);
  LFS2_CHECK_LAUNCH(name);
  return LFS2_OK;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_linear(const float* a, const float* w, const float* bias, float* c, int m, int n, int k, int relu,
                void* stream) {
  LFS2_REQUIRE(a && w && c, LFS2_ERR_INVALID_ARG, "linear: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && n > 0 && k > 0, LFS2_ERR_INVALID_ARG, "linear: bad shape");
  LFS2_REQUIRE(k % BK == 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "linear: need k %% 16 == 0 and n %% 4 == 0 (k=%d n=%d)", k, n);
  LFS2_REQUIRE(aligned16(a) && aligned16(w) && aligned16(c) && (!bias || aligned16(bias)), LFS2_ERR_INVALID_ARG,
               "linear: pointers must be 16-byte aligned");
  ConvMap cm{0, 0, 0};
  return launch_gemm(a, w, bias, c, m, n, k, k, relu, cm, stream, "linear");
}

int lfs2_conv1d_dense(const float* x, const float* wp, const float* bias, float* c, int batch, int t, int d, int n,
                      int ksize, int relu, void* stream) {
  LFS2_REQUIRE(x && wp && c, LFS2_ERR_INVALID_ARG, "conv1d_dense: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && n > 0, LFS2_ERR_INVALID_ARG, "conv1d_dense: bad shape");
  LFS2_REQUIRE(ksize > 0 && ksize % 2 == 1, LFS2_ERR_UNSUPPORTED, "conv1d_dense: kernel size %d must be odd", ksize);
  LFS2_REQUIRE(d % BK == 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "conv1d_dense: need d %% 16 == 0 and n %% 4 == 0");
  LFS2_REQUIRE(aligned16(x) && aligned16(wp) && aligned16(c) && (!bias || aligned16(bias)), LFS2_ERR_INVALID_ARG,
               "conv1d_dense: pointers must be 16-byte aligned");
  ConvMap cm{t, d, (ksize - 1) / 2};
  return launch_gemm(x, wp, bias, c, batch * t, n, ksize * d, d, relu, cm, stream, "conv1d_dense");
}

}  // extern "C"
