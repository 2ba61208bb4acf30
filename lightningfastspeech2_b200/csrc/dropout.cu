// Dropout of the train step (reference: nn.Dropout in PositionalEncoding model.py:42,55, the three
// residual-path dropouts and the attention-probability dropout of the FFTBlock model.py:111-122,
// VarianceConvolutionLayer model.py:539,557).  Masks are never stored: keep/drop of element i is a
// pure function of (seed, site, i) through Philox4x32-10, so the backward pass regenerates exactly
// the forward mask by calling the same kernel on the gradient with the same (seed, site).
// The random stream cannot match PyTorch's (different generator layout); parity is statistical
// (keep rate, 1/(1-p) scaling) plus exact forward/backward mask consistency.
#include "common.cuh"

namespace lfs2 {

__global__ void dropout_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4, uint32_t threshold,
                               float inv_keep, uint2 key, uint32_t site) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 s = dropout_scale4(i, threshold, inv_keep, key, site);
  float4 v = x[i];
  v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
  y[i] = v;
}

__device__ __forceinline__ void dsplit2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}

// bf16 hi/lo planes (attention probabilities): out = planes(in * mask / (1-p))
__global__ void dropout_planes_kernel(const uint2* __restrict__ in_hi, const uint2* __restrict__ in_lo,
                                      uint2* __restrict__ out_hi, uint2* __restrict__ out_lo, size_t n4,
                                      uint32_t threshold, float inv_keep, uint2 key, uint32_t site) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 s = dropout_scale4(i, threshold, inv_keep, key, site);
  const uint2 h = in_hi[i];
  const uint2 l = in_lo ? in_lo[i] : make_uint2(0u, 0u);
  const float v0 = (__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16)) * s.x;
  const float v1 = (__uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u)) * s.y;
  const float v2 = (__uint_as_float(h.y << 16) + __uint_as_float(l.y << 16)) * s.z;
  const float v3 = (__uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u)) * s.w;
  uint2 oh, ol;
  dsplit2(v0, v1, oh.x, ol.x);
  dsplit2(v2, v3, oh.y, ol.y);
  out_hi[i] = oh;
  if (out_lo) out_lo[i] = ol;
}

static inline bool dropout_args(float p, uint32_t* threshold, float* inv_keep) {
  if (!(p >= 0.f && p < 1.f)) return false;
  double t = (double)p * 4294967296.0;
  *threshold = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  *inv_keep = 1.f / (1.f - p);
  return true;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_dropout(const float* x, float* y, long long n, float p, unsigned long long seed, unsigned int site,
                 void* stream) {
  LFS2_REQUIRE(x && y, LFS2_ERR_INVALID_ARG, "dropout: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0 && aligned16(x) && aligned16(y), LFS2_ERR_UNSUPPORTED,
               "dropout: n must be a positive multiple of 4 and the pointers 16-byte aligned");
  uint32_t thr;
  float inv;
  LFS2_REQUIRE(dropout_args(p, &thr, &inv), LFS2_ERR_INVALID_ARG, "dropout: p=%f must be in [0, 1)", p);
  const size_t n4 = (size_t)n / 4;
  dropout_kernel<<<ceil_div((long long)n4, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (float4*)y, n4, thr, inv, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), site);
  LFS2_CHECK_LAUNCH("dropout");
  return LFS2_OK;
}

int lfs2_dropout_planes(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, long long n, float p,
                        unsigned long long seed, unsigned int site, void* stream) {
  LFS2_REQUIRE(in_hi && out_hi, LFS2_ERR_INVALID_ARG, "dropout_planes: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "dropout_planes: n must be a positive multiple of 4");
  uint32_t thr;
  float inv;
  LFS2_REQUIRE(dropout_args(p, &thr, &inv), LFS2_ERR_INVALID_ARG, "dropout_planes: p=%f must be in [0, 1)", p);
  const size_t n4 = (size_t)n / 4;
  dropout_planes_kernel<<<ceil_div((long long)n4, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint2*)in_hi, (const uint2*)in_lo, (uint2*)out_hi, (uint2*)out_lo, n4, thr, inv,
      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), site);
  LFS2_CHECK_LAUNCH("dropout_planes");
  return LFS2_OK;
}

}  // extern "C"
