// fp32 CUDA-core flash attention (streaming softmax; the (B,h,T,T) logits the reference
// materialises -- 3.9 GB per decoder layer at C2 -- never exist).  One CTA = 64 queries of
// one (utterance, head); K/V tiles of 64 keys stream through shared memory; 4x4 register
// micro-tiles for S = QK^T and 4 x DH/16 for O += P.V; row statistics by 16-lane shuffles.
// KV tiles that contain only PAD keys are skipped (block-uniform test), which removes the
// padding waste of ragged batches without changing any valid result.
//
// Exact-fp32 arithmetic: this is the on-device ground truth for the tensor-core attention.
#include <math.h>

#include "common.cuh"

namespace lfs2 {

constexpr int AQ = 64, AK = 64, APAD = 4;
constexpr int kAttnThreads = 256;

template <int DH>
__global__ void __launch_bounds__(kAttnThreads)
attention_f32_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ kpm, float* __restrict__ ctx,
                     float* __restrict__ lse, int t, int d, float scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = DH + APAD;
  constexpr int NG = DH / 64;  // float4 column groups per thread in the PV product
  float* Qs = smem;                       // [AQ][LD]
  float* KVs = Qs + AQ * LD;              // [AK][LD]
  float* Ps = KVs + AK * LD;              // [AQ][AK+APAD]
  int* s_valid = reinterpret_cast<int*>(Ps + AQ * (AK + APAD));  // [AK]

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const size_t ld_qkv = 3 * (size_t)d;
  const float* base = qkv + (size_t)b * t * ld_qkv + (size_t)h * DH;

  // Q tile, pre-scaled like torch (q * dh^-1/2 before the product)
  for (int i = tid; i < AQ * (DH / 4); i += kAttnThreads) {
    int r = i / (DH / 4), c4 = i % (DH / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < t) v = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld_qkv + c4 * 4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(Qs + r * LD + c4 * 4) = v;
  }

  float m_i[4], l_i[4];
  float4 o[4][NG];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_i[i] = -INFINITY;
    l_i[i] = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) o[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int k0 = 0; k0 < t; k0 += AK) {
    __syncthreads();  // previous iteration finished with KVs / Ps / s_valid
    int valid = 0;
    if (tid < AK) {
      int key = k0 + tid;
      valid = (key < t) && !(kpm && kpm[(size_t)b * t + key]);
      s_valid[tid] = valid;
    }
    if (!__syncthreads_or(valid)) continue;  // tile of PAD keys only
    for (int i = tid; i < AK * (DH / 4); i += kAttnThreads) {
      int r = i / (DH / 4), c4 = i % (DH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + r < t) v = *reinterpret_cast<const float4*>(base + d + (size_t)(k0 + r) * ld_qkv + c4 * 4);
      *reinterpret_cast<float4*>(KVs + r * LD + c4 * 4) = v;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int c = 0; c < DH; c += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * LD + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(KVs + (tx + 16 * j) * LD + c);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
    // mask, online softmax update (rows are shared by the 16 lanes with equal ty)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!s_valid[tx + 16 * j]) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float m_new = fmaxf(m_i[i], mx);
      float alpha = expf(m_i[i] - m_new);  // NaN when every key so far is masked, like torch
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = expf(s[i][j] - m_new);
        Ps[(ty + 16 * i) * (AK + APAD) + tx + 16 * j] = p;
        rs += p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_i[i] = l_i[i] * alpha + rs;
      m_i[i] = m_new;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        o[i][g].x *= alpha; o[i][g].y *= alpha; o[i][g].z *= alpha; o[i][g].w *= alpha;
      }
    }
    __syncthreads();  // K consumed, P visible
    for (int i = tid; i < AK * (DH / 4); i += kAttnThreads) {
      int r = i / (DH / 4), c4 = i % (DH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + r < t) v = *reinterpret_cast<const float4*>(base + 2 * d + (size_t)(k0 + r) * ld_qkv + c4 * 4);
      *reinterpret_cast<float4*>(KVs + r * LD + c4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < AK; ++j) {
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Ps[(ty + 16 * i) * (AK + APAD) + j];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float4 v = *reinterpret_cast<const float4*>(KVs + j * LD + g * 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i][g].x = fmaf(p[i], v.x, o[i][g].x);
          o[i][g].y = fmaf(p[i], v.y, o[i][g].y);
          o[i][g].z = fmaf(p[i], v.z, o[i][g].z);
          o[i][g].w = fmaf(p[i], v.w, o[i][g].w);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int q = q0 + ty + 16 * i;
    if (q >= t) continue;
    float inv = 1.f / l_i[i];  // 0/0 -> NaN for fully masked rows, like the reference
    // log-sum-exp of the scaled logits, saved for the backward pass: P = exp(S - lse)
    if (lse && tx == 0) lse[((size_t)b * gridDim.y + h) * t + q] = m_i[i] + logf(l_i[i]);
    float* orow = ctx + ((size_t)b * t + q) * d + (size_t)h * DH;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float4 v = o[i][g];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      *reinterpret_cast<float4*>(orow + g * 64 + tx * 4) = v;
    }
  }
}

template <int DH>
static int launch_attention(const float* qkv, const uint8_t* kpm, float* ctx, float* lse, int batch, int t, int d,
                            int nhead, cudaStream_t s) {
  size_t smem = sizeof(float) * (AQ * (DH + APAD) + AK * (DH + APAD) + AQ * (AK + APAD)) + sizeof(int) * AK;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_f32_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      set_error("attention: cannot reserve %zu bytes of shared memory", smem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(ceil_div(t, AQ), nhead, batch);
  float scale = 1.0f / sqrtf((float)DH);
  attention_f32_kernel<DH><<<grid, kAttnThreads, smem, s>>>(qkv, kpm, ctx, lse, t, d, scale);
  LFS2_CHECK_LAUNCH("attention");
  return LFS2_OK;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" int lfs2_attention(const float* qkv, const uint8_t* key_padding_mask, float* ctx, int batch, int t, int d,
                              int nhead, void* stream) {
  return lfs2_attention_lse(qkv, key_padding_mask, ctx, nullptr, batch, t, d, nhead, stream);
}

extern "C" int lfs2_attention_lse(const float* qkv, const uint8_t* key_padding_mask, float* ctx, float* lse, int batch,
                                  int t, int d, int nhead, void* stream) {
  LFS2_REQUIRE(qkv && ctx, LFS2_ERR_INVALID_ARG, "attention: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && nhead > 0 && d % nhead == 0, LFS2_ERR_INVALID_ARG, "attention: bad shape");
  LFS2_REQUIRE(batch <= 65535 && nhead <= 65535, LFS2_ERR_UNSUPPORTED, "attention: batch/nhead > 65535");
  LFS2_REQUIRE(aligned16(qkv) && aligned16(ctx), LFS2_ERR_INVALID_ARG, "attention: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int dh = d / nhead;
  switch (dh) {
    case 64: return launch_attention<64>(qkv, key_padding_mask, ctx, lse, batch, t, d, nhead, s);
    case 128: return launch_attention<128>(qkv, key_padding_mask, ctx, lse, batch, t, d, nhead, s);
    case 384: return launch_attention<384>(qkv, key_padding_mask, ctx, lse, batch, t, d, nhead, s);
    default:
      set_error("attention: head_dim %d not supported (64, 128, 384)", dh);
      return LFS2_ERR_UNSUPPORTED;
  }
}
