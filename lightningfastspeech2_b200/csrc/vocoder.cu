// Streaming glue kernels of the HiFi-GAN generator (reference litfass/third_party/hifigan/models.py:112-174): the
// convolutions themselves run on lfs2_gemm_tc_ex (dense / dilated / polyphase-transposed Conv1d as tap-shifted tcgen05
// GEMMs with bias + leaky-ReLU + residual epilogues); what is left is HBM-bound element-wise work on bf16 hi/lo planes:
//   mel_to_planes      (B, 80, T) channels-first fp32 mel -> (B, T, 96) channels-last planes, zero beyond each length
//   lrelu_planes       y = leaky_relu(x)                       (models.py:87,89: the activations in front of c1 / c2)
//   mean3_lrelu_planes y = leaky_relu((a + b + c) / 3)         (models.py:157-166: MRF average, then the next stage's lrelu)
//   conv_post_tanh     wav = tanh(Conv1d(C -> 1, k)(leaky_relu(x, 0.01)))   (models.py:167-169)
// 16-byte coalesced accesses throughout; thread = 8 consecutive elements (one uint4 per plane).
#include <math.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace lfs2 {

__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float (&x)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    x[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void pack8(const float (&x)[8], uint4& h, uint4& l) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hw[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    const float a = x[2 * i] - __uint_as_float(hw[i] << 16), b = x[2 * i + 1] - __uint_as_float(hw[i] & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lw[i]) : "f"(b), "f"(a));
  }
  h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

__global__ void lrelu_planes_kernel(const uint4* __restrict__ in_hi, const uint4* __restrict__ in_lo,
                                    uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, size_t n8, float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float x[8];
  unpack8(in_hi[i], in_lo[i], x);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = lrelu(x[j], slope);
  uint4 h, l;
  pack8(x, h, l);
  out_hi[i] = h;
  out_lo[i] = l;
}

__global__ void mean3_lrelu_planes_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                          const uint4* __restrict__ b_hi, const uint4* __restrict__ b_lo,
                                          const uint4* __restrict__ c_hi, const uint4* __restrict__ c_lo,
                                          uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, size_t n8, float scale,
                                          float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float a[8], b[8], c[8];
  unpack8(a_hi[i], a_lo[i], a);
  unpack8(b_hi[i], b_lo[i], b);
  unpack8(c_hi[i], c_lo[i], c);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = lrelu(((a[j] + b[j]) + c[j]) * scale, slope);  // the reference's order: (r1 + r2) + r3, / 3
  uint4 h, l;
  pack8(a, h, l);
  out_hi[i] = h;
  out_lo[i] = l;
}

// mel (B, C, T) fp32 channels-first -> planes (B, T, CP) channels-last, zero for c >= C and for t >= lengths[b].
// One thread per (b, t, 8 channels); reads are strided by T (a 96 x T transpose of a small tensor: not worth tiling).
__global__ void mel_to_planes_kernel(const float* __restrict__ mel, const int* __restrict__ lengths,
                                     uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, int batch, int c, int t,
                                     int cp8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)batch * t * cp8) return;
  const int g = (int)(i % cp8);
  const size_t row = i / cp8;
  const int tt = (int)(row % t), b = (int)(row / t);
  const bool live = !lengths || tt < lengths[b];
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = g * 8 + j;
    x[j] = (live && ch < c) ? mel[((size_t)b * c + ch) * t + tt] : 0.f;
  }
  uint4 h, l;
  pack8(x, h, l);
  out_hi[i] = h;
  out_lo[i] = l;
}

// wav[b, t] = tanh(bias + sum_j sum_c lrelu(x[b, t + j - (k-1)/2, c], slope) * w[j * C + c]); zero "same" padding at
// the ends of the (B, T) tensor, samples at or beyond lengths[b] are written as 0 and read as 0.
// One warp per 32 consecutive samples: lane = sample; the k x C weights sit in shared memory.
__global__ void conv_post_tanh_kernel(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      const int* __restrict__ lengths, float slope, float* __restrict__ out, int batch,
                                      int t, int c, int k) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < k * c; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)batch * t) return;
  const int tt = (int)(i % t), b = (int)(i / t);
  const int len = lengths ? min(lengths[b], t) : t;
  if (tt >= len) {
    out[i] = 0.f;
    return;
  }
  const int half = (k - 1) / 2;
  float acc = bias[0];
  for (int j = 0; j < k; ++j) {
    const int ts = tt + j - half;
    if (ts < 0 || ts >= len) continue;
    const uint4* rh = reinterpret_cast<const uint4*>(x_hi + ((size_t)b * t + ts) * c);
    const uint4* rl = reinterpret_cast<const uint4*>(x_lo + ((size_t)b * t + ts) * c);
    const float* wj = ws + j * c;
    for (int g = 0; g < c / 8; ++g) {
      float x[8];
      unpack8(rh[g], rl[g], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(lrelu(x[e], slope), wj[g * 8 + e], acc);
    }
  }
  out[i] = tanhf(acc);
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_lrelu_planes(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, long long n, float slope,
                      void* stream) {
  LFS2_REQUIRE(in_hi && in_lo && out_hi && out_lo, LFS2_ERR_INVALID_ARG, "lrelu_planes: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 8 == 0, LFS2_ERR_UNSUPPORTED, "lrelu_planes: n must be a positive multiple of 8");
  LFS2_REQUIRE(aligned16(in_hi) && aligned16(in_lo) && aligned16(out_hi) && aligned16(out_lo), LFS2_ERR_INVALID_ARG,
               "lrelu_planes: pointers must be 16-byte aligned");
  const size_t n8 = (size_t)n / 8;
  lrelu_planes_kernel<<<ceil_div(n8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in_hi, (const uint4*)in_lo,
                                                                          (uint4*)out_hi, (uint4*)out_lo, n8, slope);
  LFS2_CHECK_LAUNCH("lrelu_planes");
  return LFS2_OK;
}

int lfs2_mean3_lrelu_planes(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const void* c_hi,
                            const void* c_lo, void* out_hi, void* out_lo, long long n, float scale, float slope,
                            void* stream) {
  LFS2_REQUIRE(a_hi && a_lo && b_hi && b_lo && c_hi && c_lo && out_hi && out_lo, LFS2_ERR_INVALID_ARG,
               "mean3_lrelu_planes: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 8 == 0, LFS2_ERR_UNSUPPORTED, "mean3_lrelu_planes: n must be a positive multiple of 8");
  LFS2_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(c_hi) &&
                   aligned16(c_lo) && aligned16(out_hi) && aligned16(out_lo),
               LFS2_ERR_INVALID_ARG, "mean3_lrelu_planes: pointers must be 16-byte aligned");
  const size_t n8 = (size_t)n / 8;
  mean3_lrelu_planes_kernel<<<ceil_div(n8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)a_hi, (const uint4*)a_lo, (const uint4*)b_hi, (const uint4*)b_lo, (const uint4*)c_hi,
      (const uint4*)c_lo, (uint4*)out_hi, (uint4*)out_lo, n8, scale, slope);
  LFS2_CHECK_LAUNCH("mean3_lrelu_planes");
  return LFS2_OK;
}

int lfs2_mel_to_planes(const float* mel, const int* lengths, void* out_hi, void* out_lo, int batch, int c, int t,
                       int c_padded, void* stream) {
  LFS2_REQUIRE(mel && out_hi && out_lo, LFS2_ERR_INVALID_ARG, "mel_to_planes: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && c > 0 && c_padded >= c && c_padded % 8 == 0, LFS2_ERR_INVALID_ARG,
               "mel_to_planes: bad shape (padded channel count must be a multiple of 8 and >= channels)");
  LFS2_REQUIRE(aligned16(out_hi) && aligned16(out_lo), LFS2_ERR_INVALID_ARG, "mel_to_planes: outputs must be 16-byte aligned");
  const size_t n = (size_t)batch * t * (c_padded / 8);
  mel_to_planes_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mel, lengths, (uint4*)out_hi, (uint4*)out_lo,
                                                                          batch, c, t, c_padded / 8);
  LFS2_CHECK_LAUNCH("mel_to_planes");
  return LFS2_OK;
}

int lfs2_conv_post_tanh(const void* x_hi, const void* x_lo, const float* w, const float* bias, const int* lengths,
                        float slope, float* out, int batch, int t, int c, int ksize, void* stream) {
  LFS2_REQUIRE(x_hi && x_lo && w && bias && out, LFS2_ERR_INVALID_ARG, "conv_post_tanh: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && c > 0 && c % 8 == 0 && ksize > 0 && ksize % 2 == 1 && ksize * c <= 8192,
               LFS2_ERR_UNSUPPORTED, "conv_post_tanh: channels must be a multiple of 8, the kernel odd, k*c <= 8192");
  LFS2_REQUIRE(aligned16(x_hi) && aligned16(x_lo), LFS2_ERR_INVALID_ARG, "conv_post_tanh: planes must be 16-byte aligned");
  const size_t n = (size_t)batch * t;
  conv_post_tanh_kernel<<<ceil_div(n, 256), 256, ksize * c * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w, bias, lengths, slope, out, batch, t, c, ksize);
  LFS2_CHECK_LAUNCH("conv_post_tanh");
  return LFS2_OK;
}

}  // extern "C"
