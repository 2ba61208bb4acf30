// Stochastic duration predictor, inference direction: the pieces that are not a GEMM.
// Reference (relative to litfass/): third_party/stochastic_duration_predictor/sdp.py:11-70 (dilated depth-separable conv
// stack), :73-95 (element-wise affine flow, reverse), :98-164 (spline coupling flow, reverse), transforms.py:50-212
// (monotone rational-quadratic spline with linear tails, inverse), normalization.py (LayerNorm over channels);
// fastspeech2/model.py:299-309 (durations from the predicted log-durations).
// Everything runs at phoneme level (B x Tp rows of <= 256 channels): a few hundred KB per call, so these are plain
// streaming kernels; the 1x1 convolutions between them are lfs2_linear.  Activations are channels-last (B, T, C); the
// 2-channel flow state z is (B, T, 2).  PAD rows (mask != 0) are read as zeros by the convolution, exactly like the
// reference's `x * x_mask`, and written as zeros at the end of every flow.
#include <math.h>

#include "common.cuh"

namespace lfs2 {

// out[b,t,c] = bias[c] + sum_j wt[j,c] * x[b, t + (j - (k-1)/2) * dil, c]   (rows outside [0,T) or with mask != 0 are zero)
__global__ void sdp_dwconv_kernel(const float4* __restrict__ x, const uint8_t* __restrict__ pad_mask,
                                  const float4* __restrict__ wt, const float4* __restrict__ bias, float4* __restrict__ out,
                                  int t, int c4, int ksize, int dil, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % c4);
  const size_t row = i / c4;
  const int tt = (int)(row % t);
  const size_t b = row / t;
  float4 acc = bias[c];
  const int half = (ksize - 1) / 2;
  for (int j = 0; j < ksize; ++j) {
    const int ts = tt + (j - half) * dil;
    if (ts < 0 || ts >= t) continue;
    if (pad_mask && pad_mask[b * t + ts]) continue;
    const float4 xv = x[(b * t + ts) * c4 + c];
    const float4 w = wt[(size_t)j * c4 + c];
    acc.x = fmaf(w.x, xv.x, acc.x);
    acc.y = fmaf(w.y, xv.y, acc.y);
    acc.z = fmaf(w.z, xv.z, acc.z);
    acc.w = fmaf(w.w, xv.w, acc.w);
  }
  out[i] = acc;
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// out[m,:] = [res[m,:] +] gelu(LayerNorm(y[m,:]; gamma, beta, eps)); one warp per row, two-pass statistics, c <= 1024
constexpr int kSdpMaxVec = 8;
__global__ void sdp_ln_gelu_kernel(const float4* __restrict__ y, const float4* __restrict__ gamma,
                                   const float4* __restrict__ beta, float eps, const float4* __restrict__ res,
                                   float4* __restrict__ out, int m, int c4) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  const float4* yr = y + (size_t)row * c4;
  float4 v[kSdpMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kSdpMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < c4) {
      v[i] = yr[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)(4 * c4);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kSdpMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < c4) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)(4 * c4) + eps);
#pragma unroll
  for (int i = 0; i < kSdpMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < c4) {
      const float4 g = gamma[c], bt = beta[c];
      float4 o;
      o.x = gelu_erf((v[i].x - mean) * rstd * g.x + bt.x);
      o.y = gelu_erf((v[i].y - mean) * rstd * g.y + bt.y);
      o.z = gelu_erf((v[i].z - mean) * rstd * g.z + bt.z);
      o.w = gelu_erf((v[i].w - mean) * rstd * g.w + bt.w);
      if (res) {
        const float4 r = res[(size_t)row * c4 + c];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      out[(size_t)row * c4 + c] = o;
    }
  }
}

// h[m,:] = z[m, ch] * w[:] + bias[:] + g[m,:]   (the flow's Conv1d(1, C, 1) on its untouched half, plus the conditioning)
__global__ void sdp_flow_pre_kernel(const float* __restrict__ z, int ch, const float4* __restrict__ w,
                                    const float4* __restrict__ bias, const float4* __restrict__ g, float4* __restrict__ out,
                                    int c4, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % c4);
  const size_t row = i / c4;
  const float x0 = z[row * 2 + ch];
  const float4 wv = w[c], bv = bias[c], gv = g[i];
  out[i] = make_float4((x0 * wv.x + bv.x) + gv.x, (x0 * wv.y + bv.y) + gv.y, (x0 * wv.z + bv.z) + gv.z,
                       (x0 * wv.w + bv.w) + gv.w);
}

constexpr int kSplineBins = 10;
// knots of one axis: softmax -> floor + rescale -> cumulative sum -> [-bound, bound] with exact end points
__device__ __forceinline__ void spline_knots(const float* u, float div, float bound, float (&cum)[kSplineBins + 1]) {
  float mx = u[0] / div;
#pragma unroll
  for (int i = 1; i < kSplineBins; ++i) mx = fmaxf(mx, u[i] / div);
  float e[kSplineBins], sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSplineBins; ++i) {
    e[i] = expf(u[i] / div - mx);
    sum += e[i];
  }
  const float kMin = 1e-3f;
  float run = 0.f;
  cum[0] = -bound;
#pragma unroll
  for (int i = 0; i < kSplineBins; ++i) {
    run += kMin + (1.f - kMin * kSplineBins) * (e[i] / sum);
    cum[i + 1] = 2.f * bound * run - bound;
  }
  cum[kSplineBins] = bound;
}
__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// one flow step, reverse direction, in place on z (m, 2): the half `x1c` goes through the inverse spline whose 29
// parameters per row are h[m, 0:29] (row stride hs), the other half is kept; both are zeroed on PAD rows
__global__ void sdp_spline_inverse_kernel(float* __restrict__ z, int x1c, const float* __restrict__ h, int hs,
                                          const uint8_t* __restrict__ pad_mask, float sqrt_hidden, float bound, int m) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  if (pad_mask && pad_mask[row]) {
    z[2 * row] = 0.f;
    z[2 * row + 1] = 0.f;
    return;
  }
  const float* p = h + (size_t)row * hs;
  float u[3 * kSplineBins - 1];
#pragma unroll
  for (int i = 0; i < 3 * kSplineBins - 1; ++i) u[i] = p[i];
  const float y = z[2 * row + x1c];
  if (!(y >= -bound && y <= bound)) return;  // identity tails (NaN stays NaN)
  float cw[kSplineBins + 1], chh[kSplineBins + 1];
  spline_knots(u, sqrt_hidden, bound, cw);
  spline_knots(u + kSplineBins, sqrt_hidden, bound, chh);
  int bin = -1;
#pragma unroll
  for (int i = 0; i <= kSplineBins; ++i) bin += (y >= (i == kSplineBins ? chh[i] + 1e-6f : chh[i])) ? 1 : 0;
  bin = min(max(bin, 0), kSplineBins - 1);
  // derivatives at the bin's two knots; the outermost knots have derivative min + softplus(log(exp(1 - min) - 1)) = 1
  const float kMinD = 1e-3f;
  const float edge = logf(expf(1.f - kMinD) - 1.f);
  float cwb = 0.f, wb = 0.f, chb = 0.f, hb = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < kSplineBins; ++i) {
    if (i == bin) {
      cwb = cw[i];
      wb = cw[i + 1] - cw[i];
      chb = chh[i];
      hb = chh[i + 1] - chh[i];
      d0 = kMinD + softplus_f(i == 0 ? edge : u[2 * kSplineBins + i - 1]);
      d1 = kMinD + softplus_f(i == kSplineBins - 1 ? edge : u[2 * kSplineBins + i]);
    }
  }
  const float delta = hb / wb;
  const float tt = y - chb;
  const float s2 = d0 + d1 - 2.f * delta;
  const float qa = tt * s2 + hb * (delta - d0);
  const float qb = hb * d0 - tt * s2;
  const float qc = -delta * tt;
  const float root = (2.f * qc) / (-qb - sqrtf(qb * qb - 4.f * qa * qc));
  z[2 * row + x1c] = root * wb + cwb;
}

// element-wise affine flow, reverse: logical channel c = physical channel c ^ flip
__global__ void sdp_affine_reverse_kernel(float* __restrict__ z, const float* __restrict__ translation,
                                          const float* __restrict__ log_scale, const uint8_t* __restrict__ pad_mask,
                                          int flip, int m) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  const bool pad = pad_mask && pad_mask[row];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int pc = c ^ flip;
    z[2 * row + pc] = pad ? 0.f : (z[2 * row + pc] - translation[c]) * expf(-log_scale[c]);
  }
}

// durations of the stochastic branch (model.py:302-309): ceil(exp(logw + 1e-9)), 0 where logw == 0, clamp, int32, and
// the all-ones guard; one block per utterance
__global__ void sdp_duration_kernel(const float* __restrict__ logw, const uint8_t* __restrict__ src_mask,
                                    int32_t* __restrict__ dur, int tp) {
  const int b = blockIdx.x;
  const float* p = logw + (size_t)b * tp;
  const uint8_t* mk = src_mask + (size_t)b * tp;
  int32_t* o = dur + (size_t)b * tp;
  __shared__ long long s_total;
  __shared__ int s_valid;
  if (threadIdx.x == 0) {
    s_total = 0;
    s_valid = 0;
  }
  __syncthreads();
  long long total = 0;
  int nvalid = 0;
  for (int i = threadIdx.x; i < tp; i += blockDim.x) {
    float v = p[i] == 0.f ? 0.f : ceilf(expf(__fadd_rn(p[i], 1e-9f)));
    v = fmaxf(v, 0.f);
    const int32_t di = v >= 2147483520.f ? 2147483647 : (int32_t)v;
    o[i] = di;
    if (!mk[i]) {
      total += di;
      ++nvalid;
    }
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(&s_total), (unsigned long long)total);
  atomicAdd(&s_valid, nvalid);
  __syncthreads();
  if (s_total <= (long long)(s_valid / 2)) {
    for (int i = threadIdx.x; i < tp; i += blockDim.x)
      if (!mk[i]) o[i] = 1;
  }
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_sdp_dwconv(const float* x, const uint8_t* pad_mask, const float* wt, const float* bias, float* out, int batch,
                    int t, int c, int ksize, int dilation, void* stream) {
  LFS2_REQUIRE(x && wt && bias && out, LFS2_ERR_INVALID_ARG, "sdp_dwconv: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && c > 0 && c % 4 == 0 && ksize > 0 && ksize % 2 == 1 && dilation > 0, LFS2_ERR_UNSUPPORTED,
               "sdp_dwconv: need c %% 4 == 0, an odd kernel size and a positive dilation");
  LFS2_REQUIRE(aligned16(x) && aligned16(wt) && aligned16(bias) && aligned16(out), LFS2_ERR_INVALID_ARG,
               "sdp_dwconv: pointers must be 16-byte aligned");
  const size_t total = (size_t)batch * t * (c / 4);
  sdp_dwconv_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, pad_mask, (const float4*)wt,
                                                                            (const float4*)bias, (float4*)out, t, c / 4,
                                                                            ksize, dilation, total);
  LFS2_CHECK_LAUNCH("sdp_dwconv");
  return LFS2_OK;
}

int lfs2_sdp_ln_gelu(const float* y, const float* gamma, const float* beta, float eps, const float* res, float* out, int m,
                     int c, void* stream) {
  LFS2_REQUIRE(y && gamma && beta && out, LFS2_ERR_INVALID_ARG, "sdp_ln_gelu: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && c > 0 && c % 4 == 0 && c <= 128 * kSdpMaxVec, LFS2_ERR_UNSUPPORTED,
               "sdp_ln_gelu: c=%d must be a multiple of 4 and <= %d", c, 128 * kSdpMaxVec);
  LFS2_REQUIRE(aligned16(y) && aligned16(gamma) && aligned16(beta) && aligned16(res) && aligned16(out), LFS2_ERR_INVALID_ARG,
               "sdp_ln_gelu: pointers must be 16-byte aligned");
  sdp_ln_gelu_kernel<<<ceil_div((long long)m * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)y, (const float4*)gamma, (const float4*)beta, eps, (const float4*)res, (float4*)out, m, c / 4);
  LFS2_CHECK_LAUNCH("sdp_ln_gelu");
  return LFS2_OK;
}

int lfs2_sdp_flow_pre(const float* z, int channel, const float* w, const float* bias, const float* g, float* out, int m,
                      int c, void* stream) {
  LFS2_REQUIRE(z && w && bias && g && out && (channel == 0 || channel == 1), LFS2_ERR_INVALID_ARG, "sdp_flow_pre: bad argument");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && c > 0 && c % 4 == 0, LFS2_ERR_UNSUPPORTED, "sdp_flow_pre: c must be a multiple of 4");
  LFS2_REQUIRE(aligned16(w) && aligned16(bias) && aligned16(g) && aligned16(out), LFS2_ERR_INVALID_ARG,
               "sdp_flow_pre: pointers must be 16-byte aligned");
  const size_t total = (size_t)m * (c / 4);
  sdp_flow_pre_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(z, channel, (const float4*)w, (const float4*)bias,
                                                                              (const float4*)g, (float4*)out, c / 4, total);
  LFS2_CHECK_LAUNCH("sdp_flow_pre");
  return LFS2_OK;
}

int lfs2_sdp_spline_inverse(float* z, int x1_channel, const float* h, int h_stride, const uint8_t* pad_mask,
                            int hidden_channels, float tail_bound, int m, void* stream) {
  LFS2_REQUIRE(z && h && (x1_channel == 0 || x1_channel == 1), LFS2_ERR_INVALID_ARG, "sdp_spline_inverse: bad argument");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && h_stride >= 3 * kSplineBins - 1 && hidden_channels > 0 && tail_bound > 0.f, LFS2_ERR_INVALID_ARG,
               "sdp_spline_inverse: bad shape");
  sdp_spline_inverse_kernel<<<ceil_div(m, 128), 128, 0, (cudaStream_t)stream>>>(
      z, x1_channel, h, h_stride, pad_mask, (float)sqrt((double)hidden_channels), tail_bound, m);
  LFS2_CHECK_LAUNCH("sdp_spline_inverse");
  return LFS2_OK;
}

int lfs2_sdp_affine_reverse(float* z, const float* translation, const float* log_scale, const uint8_t* pad_mask, int flip,
                            int m, void* stream) {
  LFS2_REQUIRE(z && translation && log_scale && (flip == 0 || flip == 1), LFS2_ERR_INVALID_ARG, "sdp_affine_reverse: bad argument");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0, LFS2_ERR_INVALID_ARG, "sdp_affine_reverse: bad shape");
  sdp_affine_reverse_kernel<<<ceil_div(m, 128), 128, 0, (cudaStream_t)stream>>>(z, translation, log_scale, pad_mask, flip, m);
  LFS2_CHECK_LAUNCH("sdp_affine_reverse");
  return LFS2_OK;
}

int lfs2_sdp_durations(const float* logw, const uint8_t* src_mask, int32_t* dur, int batch, int tp, void* stream) {
  LFS2_REQUIRE(logw && src_mask && dur, LFS2_ERR_INVALID_ARG, "sdp_durations: null pointer");
  if (batch == 0 || tp == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && tp > 0, LFS2_ERR_INVALID_ARG, "sdp_durations: bad shape");
  sdp_duration_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(logw, src_mask, dur, tp);
  LFS2_CHECK_LAUNCH("sdp_durations");
  return LFS2_OK;
}

}  // extern "C"
