// Row-wise kernels of the GEMM-decomposed attention (any head_dim that is a multiple of 32; forward
// for the shapes the fused tcgen05 flash kernel does not cover, and the backward of every shape):
//
//   forward   S = Q.K^T            lfs2_gemm_tc2 (K-major x K-major)       -> fp32 (Z, T, Tp)
//             P = softmax(S*scale) attn_softmax_planes (this file)         -> bf16 hi/lo planes (Z, T, Tp) + lse
//             O = P.V              lfs2_gemm_tc2 (K-major x MN-major)      -> ctx (B, T, d)
//   backward  dP = dO.V^T          lfs2_gemm_tc2 (K-major x K-major)       -> fp32 (Z, T, Tp)
//             dS = scale * P o (dP - delta), delta = rowsum(dO o O)        attn_ds_planes (this file)
//             dV = P^T.dO, dK = dS^T.Q  (MN-major x MN-major), dQ = dS.K  (K-major x MN-major)
//
// Z = B * nhead, Tp = T rounded up to 8 (TMA row pitch); the pad columns and PAD keys hold P = dS = 0.
// The probabilities stay resident (2 x 2 bytes per logit) between forward and backward: with 180 GB
// of HBM the (Z, T, T) planes of a 76 M-parameter train step fit many times over, and the backward
// saves the S recomputation.  All HBM-bound streaming kernels: one warp per row, 16-byte accesses.
#include <math.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace lfs2 {

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}
__device__ __forceinline__ float2 merge2(uint32_t hi, uint32_t lo) {
  return make_float2(__uint_as_float(hi << 16) + __uint_as_float(lo << 16),
                     __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u));
}

// one warp per (z, query) row.  s: (Z, T, Tp) raw logits; kpm: (B, T) 1 = PAD key
__global__ void __launch_bounds__(256)
attn_softmax_planes_kernel(const float* __restrict__ s, const uint8_t* __restrict__ kpm, uint32_t* __restrict__ p_hi,
                           uint32_t* __restrict__ p_lo, uint32_t* __restrict__ pm_hi, uint32_t* __restrict__ pm_lo,
                           float* __restrict__ lse, long long rows, int t, int tp, int nhead, float scale,
                           DropSite drop) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (int)(row / t / nhead);
  const float* sr = s + row * tp;
  const uint8_t* mk = kpm ? kpm + (size_t)b * t : nullptr;
  float mx = -INFINITY;
  for (int c = lane * 4; c < t; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(sr + c);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < t && !(mk && mk[c + j])) mx = fmaxf(mx, e[j] * scale);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane * 4; c < t; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(sr + c);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < t && !(mk && mk[c + j])) sum += expf(e[j] * scale - mx);
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;  // no unmasked key: 1/0 -> NaN row like the reference
  if (lane == 0) lse[row] = mx + logf(sum);
  uint2* ph = reinterpret_cast<uint2*>(p_hi + row * (tp / 2));
  uint2* pl = p_lo ? reinterpret_cast<uint2*>(p_lo + row * (tp / 2)) : nullptr;
  // attention-probability dropout fused: P (kept for the backward) and P o mask / (1-p) (the operand of P.V)
  uint2* mh = pm_hi ? reinterpret_cast<uint2*>(pm_hi + row * (tp / 2)) : nullptr;
  uint2* ml = pm_lo ? reinterpret_cast<uint2*>(pm_lo + row * (tp / 2)) : nullptr;
  for (int c = lane * 4; c < tp; c += 128) {
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < t) {
      const float4 v = *reinterpret_cast<const float4*>(sr + c);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < t) pv[j] = (mk && mk[c + j]) ? (sum > 0.f ? 0.f : __int_as_float(0x7fc00000)) : expf(e[j] * scale - mx) * inv;
    }
    uint2 h, l;
    split2(pv[0], pv[1], h.x, l.x);
    split2(pv[2], pv[3], h.y, l.y);
    ph[c / 4] = h;
    if (pl) pl[c / 4] = l;
    if (mh) {
      const float4 k = dropout_scale4((size_t)row * (tp / 4) + c / 4, drop.threshold, drop.inv_keep, drop.key, drop.site);
      split2(pv[0] * k.x, pv[1] * k.y, h.x, l.x);
      split2(pv[2] * k.z, pv[3] * k.w, h.y, l.y);
      mh[c / 4] = h;
      if (ml) ml[c / 4] = l;
    }
  }
}

// delta[z, q] = sum_c dO[b, q, h*dh + c] * O[b, q, h*dh + c]; one warp per (b, q, h)
__global__ void attn_delta_mat_kernel(const float4* __restrict__ dctx, const float4* __restrict__ ctx,
                                      float* __restrict__ delta, int batch, int t, int nhead, int dh4) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= batch * t * nhead) return;
  const int h = warp % nhead;
  const int row = warp / nhead;
  const size_t base = ((size_t)row * nhead + h) * dh4;
  float s = 0.f;
  for (int c = lane; c < dh4; c += 32) {
    float4 a = dctx[base + c], o = ctx[base + c];
    s += (a.x * o.x + a.y * o.y) + (a.z * o.z + a.w * o.w);
  }
  s = warp_sum(s);
  if (lane == 0) delta[((size_t)(row / t) * nhead + h) * t + row % t] = s;
}

// dS = scale * P o (dP - delta[row]) as hi/lo planes; one thread per 4 logits
__global__ void __launch_bounds__(256)
attn_ds_planes_kernel(const uint2* __restrict__ p_hi, const uint2* __restrict__ p_lo, const float4* __restrict__ dp,
                      const float* __restrict__ delta, uint2* __restrict__ ds_hi, uint2* __restrict__ ds_lo,
                      long long rows, int tp4, float scale, DropSite drop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * tp4) return;
  const long long row = i / tp4;
  const uint2 h = p_hi[i];
  const uint2 l = p_lo ? p_lo[i] : make_uint2(0u, 0u);
  const float2 p01 = merge2(h.x, l.x), p23 = merge2(h.y, l.y);
  float4 g = dp[i];
  if (drop.threshold) {  // O = (P o M).V  =>  dP_eff = M o (dO.V^T): the forward's mask, regenerated
    const float4 k = dropout_scale4((size_t)i, drop.threshold, drop.inv_keep, drop.key, drop.site);
    g.x *= k.x; g.y *= k.y; g.z *= k.z; g.w *= k.w;
  }
  const float dl = delta[row];
  // P = 0 on PAD keys / pad columns, where dP may hold anything finite
  const float d0 = p01.x != 0.f ? scale * p01.x * (g.x - dl) : 0.f;
  const float d1 = p01.y != 0.f ? scale * p01.y * (g.y - dl) : 0.f;
  const float d2 = p23.x != 0.f ? scale * p23.x * (g.z - dl) : 0.f;
  const float d3 = p23.y != 0.f ? scale * p23.y * (g.w - dl) : 0.f;
  uint2 oh, ol;
  split2(d0, d1, oh.x, ol.x);
  split2(d2, d3, oh.y, ol.y);
  ds_hi[i] = oh;
  if (ds_lo) ds_lo[i] = ol;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_attn_softmax_planes(const float* s, const uint8_t* key_padding_mask, void* p_hi, void* p_lo, float* lse,
                             int batch, int nhead, int t, int tp, float scale, void* stream) {
  return lfs2_attn_softmax_planes_drop(s, key_padding_mask, p_hi, p_lo, nullptr, nullptr, lse, batch, nhead, t, tp, scale,
                                       0.f, 0ull, 0u, stream);
}

int lfs2_attn_softmax_planes_drop(const float* s, const uint8_t* key_padding_mask, void* p_hi, void* p_lo, void* pm_hi,
                                  void* pm_lo, float* lse, int batch, int nhead, int t, int tp, float scale,
                                  float drop_p, unsigned long long drop_seed, unsigned int drop_site, void* stream) {
  LFS2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LFS2_ERR_INVALID_ARG, "attn_softmax_planes: dropout p must be in [0, 1)");
  LFS2_REQUIRE((!pm_hi || aligned16(pm_hi)) && (!pm_lo || aligned16(pm_lo)), LFS2_ERR_INVALID_ARG,
               "attn_softmax_planes: pointers must be 16-byte aligned");
  LFS2_REQUIRE(s && p_hi && lse, LFS2_ERR_INVALID_ARG, "attn_softmax_planes: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && nhead > 0 && t > 0 && tp >= t && tp % 8 == 0, LFS2_ERR_INVALID_ARG,
               "attn_softmax_planes: bad shape (tp must be t rounded up to a multiple of 8)");
  LFS2_REQUIRE(aligned16(s) && aligned16(p_hi) && (!p_lo || aligned16(p_lo)), LFS2_ERR_INVALID_ARG,
               "attn_softmax_planes: pointers must be 16-byte aligned");
  const long long rows = (long long)batch * nhead * t;
  attn_softmax_planes_kernel<<<ceil_div(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      s, key_padding_mask, (uint32_t*)p_hi, (uint32_t*)p_lo, (uint32_t*)pm_hi, (uint32_t*)pm_lo, lse, rows, t, tp, nhead,
      scale, make_drop_site(pm_hi ? drop_p : 0.f, drop_seed, drop_site));
  LFS2_CHECK_LAUNCH("attn_softmax_planes");
  return LFS2_OK;
}

int lfs2_attn_delta(const float* dctx, const float* ctx, float* delta, int batch, int t, int d, int nhead,
                    void* stream) {
  LFS2_REQUIRE(dctx && ctx && delta, LFS2_ERR_INVALID_ARG, "attn_delta: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && nhead > 0 && d % nhead == 0 && (d / nhead) % 4 == 0, LFS2_ERR_INVALID_ARG, "attn_delta: bad shape");
  LFS2_REQUIRE(aligned16(dctx) && aligned16(ctx), LFS2_ERR_INVALID_ARG, "attn_delta: pointers must be 16-byte aligned");
  attn_delta_mat_kernel<<<ceil_div((long long)batch * t * nhead * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)dctx, (const float4*)ctx, delta, batch, t, nhead, d / nhead / 4);
  LFS2_CHECK_LAUNCH("attn_delta");
  return LFS2_OK;
}

int lfs2_attn_ds_planes(const void* p_hi, const void* p_lo, const float* dp, const float* delta, void* ds_hi,
                        void* ds_lo, int batch, int nhead, int t, int tp, float scale, void* stream) {
  return lfs2_attn_ds_planes_drop(p_hi, p_lo, dp, delta, ds_hi, ds_lo, batch, nhead, t, tp, scale, 0.f, 0ull, 0u, stream);
}

int lfs2_attn_ds_planes_drop(const void* p_hi, const void* p_lo, const float* dp, const float* delta, void* ds_hi,
                             void* ds_lo, int batch, int nhead, int t, int tp, float scale, float drop_p,
                             unsigned long long drop_seed, unsigned int drop_site, void* stream) {
  LFS2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LFS2_ERR_INVALID_ARG, "attn_ds_planes: dropout p must be in [0, 1)");
  LFS2_REQUIRE(p_hi && dp && delta && ds_hi, LFS2_ERR_INVALID_ARG, "attn_ds_planes: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && nhead > 0 && t > 0 && tp >= t && tp % 8 == 0, LFS2_ERR_INVALID_ARG, "attn_ds_planes: bad shape");
  LFS2_REQUIRE(aligned16(p_hi) && aligned16(dp) && aligned16(ds_hi) && (!p_lo || aligned16(p_lo)) &&
                   (!ds_lo || aligned16(ds_lo)),
               LFS2_ERR_INVALID_ARG, "attn_ds_planes: pointers must be 16-byte aligned");
  const long long rows = (long long)batch * nhead * t;
  const long long n = rows * (tp / 4);
  attn_ds_planes_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint2*)p_hi, (const uint2*)p_lo, (const float4*)dp, delta, (uint2*)ds_hi, (uint2*)ds_lo, rows, tp / 4,
      scale, make_drop_site(drop_p, drop_seed, drop_site));
  LFS2_CHECK_LAUNCH("attn_ds_planes");
  return LFS2_OK;
}

}  // extern "C"
