// Streaming glue of the FastDiff variance adaptor (reference litfass/fastspeech2/fastdiff_variances.py):
// the noise-predicting networks are the same VarianceConvolutionLayer stacks as the plain adaptor's predictors (dwconv +
// tcgen05 GEMM with ReLU + LayerNorm epilogue); what differs is the front end -- the noisy scalar track is lifted to d
// channels, the encoder output (the condition) and a diffusion-step embedding are added -- and the DDPM update between
// the network calls.
//   diffusion_step_embed   [sin(t e_i) | cos(t e_i)], e_i = 10000^(-i / (half - 1))   (third_party/fastdiff/module/util.py:318-343)
//   swish                  x * sigmoid(x)                                              (FastDiff.py swish, fastdiff_variances.py:196-197)
//   diffusion_input        out[b,t,:] = x[b,t] * w_in + b_in + c[b,t,:] + noise_embed[b,:]   (fastdiff_variances.py:199-208)
//   diffusion_mix          out[b,t] = a[b] * x[b,t] + d[b] * z[b,t]   (q(x_t | x_0), :185-190; with a = 1/sqrt(1-beta) ... the
//                          reverse update x <- (x - k eps) / sqrt(1 - beta) [+ sigma * noise], util.py:224-228)
#include <math.h>

#include "common.cuh"

namespace lfs2 {

__global__ void diffusion_step_embed_kernel(const float* __restrict__ steps, float* __restrict__ out, int batch, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i % dim, half = dim / 2;
  const int k = j < half ? j : j - half;
  // the reference builds the frequencies in fp32: exp(arange(half) * -(log(10000) / (half - 1)))
  const float e = expf((float)k * -(logf(10000.f) / (float)(half - 1)));
  const float a = steps[b] * e;
  out[i] = j < half ? sinf(a) : cosf(a);
}

__global__ void swish_kernel(float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = x[i];
    x[i] = v / (1.f + expf(-v));
  }
}

// one thread per (row, float4 of channels)
__global__ void diffusion_input_kernel(const float* __restrict__ xt, const float4* __restrict__ w_in,
                                       const float4* __restrict__ b_in, const float4* __restrict__ c,
                                       const float4* __restrict__ ne, float4* __restrict__ out, int batch, int t, int d4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)batch * t * d4) return;
  const int ch = (int)(i % d4);
  const size_t row = i / d4;
  const int b = (int)(row / t);
  const float x = xt[row];
  const float4 w = w_in[ch], bb = b_in[ch], cc = c[i], n = ne[(size_t)b * d4 + ch];
  float4 o;
  // the reference's order: (linear_in(x) + c) + noise_embed
  o.x = (fmaf(x, w.x, bb.x) + cc.x) + n.x;
  o.y = (fmaf(x, w.y, bb.y) + cc.y) + n.y;
  o.z = (fmaf(x, w.z, bb.z) + cc.z) + n.z;
  o.w = (fmaf(x, w.w, bb.w) + cc.w) + n.w;
  out[i] = o;
}

// out[b, t] = (a[b] * x[b, t] + e[b] * y[b, t]) * s[b] + g[b] * z[b, t] + add   (y, z and the per-row vectors may be
// null); positions with zero_mask[b, t] != 0 are written as 0
__global__ void diffusion_mix_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                     const float* __restrict__ a, const float* __restrict__ e, const float* __restrict__ s,
                                     const float* __restrict__ g, float add, const uint8_t* __restrict__ zero_mask,
                                     float* __restrict__ out, int batch, int t) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)batch * t) return;
  const int b = (int)(i / t);
  float v = (a ? a[b] : 1.f) * x[i];
  if (y) v += e[b] * y[i];
  if (s) v *= s[b];
  if (z) v += g[b] * z[i];
  v += add;
  out[i] = (zero_mask && zero_mask[i]) ? 0.f : v;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_diffusion_step_embed(const float* steps, float* out, int batch, int dim, void* stream) {
  LFS2_REQUIRE(steps && out, LFS2_ERR_INVALID_ARG, "diffusion_step_embed: null pointer");
  if (batch == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && dim >= 4 && dim % 2 == 0, LFS2_ERR_INVALID_ARG, "diffusion_step_embed: dim must be even and >= 4");
  diffusion_step_embed_kernel<<<ceil_div((long long)batch * dim, 256), 256, 0, (cudaStream_t)stream>>>(steps, out, batch, dim);
  LFS2_CHECK_LAUNCH("diffusion_step_embed");
  return LFS2_OK;
}

int lfs2_swish(float* x, long long n, void* stream) {
  LFS2_REQUIRE(x || n == 0, LFS2_ERR_INVALID_ARG, "swish: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0, LFS2_ERR_INVALID_ARG, "swish: bad size");
  swish_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n);
  LFS2_CHECK_LAUNCH("swish");
  return LFS2_OK;
}

int lfs2_diffusion_input(const float* xt, const float* w_in, const float* b_in, const float* c, const float* noise_embed,
                         float* out, int batch, int t, int d, void* stream) {
  LFS2_REQUIRE(xt && w_in && b_in && c && noise_embed && out, LFS2_ERR_INVALID_ARG, "diffusion_input: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "diffusion_input: d must be a multiple of 4");
  LFS2_REQUIRE(aligned16(w_in) && aligned16(b_in) && aligned16(c) && aligned16(noise_embed) && aligned16(out),
               LFS2_ERR_INVALID_ARG, "diffusion_input: pointers must be 16-byte aligned");
  const size_t n = (size_t)batch * t * (d / 4);
  diffusion_input_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      xt, (const float4*)w_in, (const float4*)b_in, (const float4*)c, (const float4*)noise_embed, (float4*)out, batch, t, d / 4);
  LFS2_CHECK_LAUNCH("diffusion_input");
  return LFS2_OK;
}

int lfs2_diffusion_mix(const float* x, const float* y, const float* z, const float* a, const float* e, const float* s,
                       const float* g, float add, const uint8_t* zero_mask, float* out, int batch, int t, void* stream) {
  LFS2_REQUIRE(x && out && (!y || e) && (!z || g), LFS2_ERR_INVALID_ARG, "diffusion_mix: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0, LFS2_ERR_INVALID_ARG, "diffusion_mix: bad shape");
  diffusion_mix_kernel<<<ceil_div((long long)batch * t, 256), 256, 0, (cudaStream_t)stream>>>(x, y, z, a, e, s, g, add,
                                                                                           zero_mask, out, batch, t);
  LFS2_CHECK_LAUNCH("diffusion_mix");
  return LFS2_OK;
}

}  // extern "C"
