// fp32 CUDA-core backward of the multi-head self attention core (flash-style: the (B,h,T,T)
// probabilities are recomputed tile by tile from the saved log-sum-exp, never stored).
//
//   forward (attention_simt.cu):  S = (q*scale).K^T, PAD keys -inf, P = softmax(S) = exp(S - lse), O = P.V
//   backward:  delta_i = sum_c dO_ic O_ic ;  dP = dO.V^T ;  dS = P o (dP - delta)
//              dV = P^T.dO ;  dK = dS^T.(q*scale) ;  dQ = (dS.K) * scale
//
// One kernel, two roles (MODE): a CTA owns TQ "fixed" rows of one (utterance, head) and streams
// tiles of the other side through shared memory:
//   MODE 0: fixed = queries (F1 = q*scale, F2 = dO), streamed = keys (G1 = K, G2 = V)   -> dQ
//   MODE 1: fixed = keys    (F1 = K, F2 = V),        streamed = queries (G1 = q*scale, G2 = dO) -> dK, dV
// so every output row is written by exactly one CTA (deterministic, no atomics).  PAD *queries* are
// real rows (the reference computes them and they feed valid rows through the conv halos); PAD keys
// get P = 0.  This is the exact-fp32 ground truth for the gradient path.
#include <math.h>

#include "common.cuh"

namespace lfs2 {

constexpr int kAbThreads = 256;
constexpr int kAbPad = 4;

__global__ void attn_delta_kernel(const float4* __restrict__ dctx, const float4* __restrict__ ctx,
                                  float* __restrict__ delta, int batch, int t, int nhead, int dh4) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= batch * t * nhead) return;
  const int h = warp % nhead;
  const int row = warp / nhead;  // b * t + q
  const size_t base = ((size_t)row * nhead + h) * dh4;
  float s = 0.f;
  for (int c = lane; c < dh4; c += 32) {
    float4 a = dctx[base + c], o = ctx[base + c];
    s += (a.x * o.x + a.y * o.y) + (a.z * o.z + a.w * o.w);
  }
  s = warp_sum(s);
  if (lane == 0) delta[((size_t)(row / t) * nhead + h) * t + row % t] = s;
}

template <int DH, int TQ, int MODE>
__global__ void __launch_bounds__(kAbThreads)
attention_bwd_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx, const float* __restrict__ lse,
                         const float* __restrict__ delta, const uint8_t* __restrict__ kpm, float* __restrict__ dqkv,
                         int t, int d, float scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = DH + kAbPad;
  constexpr int LP = TQ + kAbPad;
  constexpr int R = TQ / 16;    // micro-tile edge of the TQ x TQ products
  constexpr int NG = DH / 64;   // float4 column groups per thread in the (TQ x DH) accumulators
  float* F1 = smem;
  float* F2 = F1 + TQ * LD;
  float* G1 = F2 + TQ * LD;
  float* G2 = G1 + TQ * LD;
  float* Ps = G2 + TQ * LD;     // [TQ][LP]  P   (MODE 1 only)
  float* Ds = Ps + TQ * LP;     // [TQ][LP]  dS
  float* s_lse = Ds + TQ * LP;  // [TQ] lse of the tile's queries
  float* s_del = s_lse + TQ;    // [TQ] delta of the tile's queries
  int* s_kvalid = reinterpret_cast<int*>(s_del + TQ);  // [TQ] key validity of the tile's keys

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int r0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z, nhead = gridDim.y;
  const size_t ld3 = 3 * (size_t)d;
  const float* qbase = qkv + (size_t)b * t * ld3 + (size_t)h * DH;
  const float* kbase = qbase + d;
  const float* vbase = qbase + 2 * d;
  const float* dobase = dctx + (size_t)b * t * d + (size_t)h * DH;
  const float* lse_b = lse + ((size_t)b * nhead + h) * t;
  const float* del_b = delta + ((size_t)b * nhead + h) * t;
  const uint8_t* kpm_b = kpm ? kpm + (size_t)b * t : nullptr;

  auto load_tile = [&](float* dst, const float* src, size_t ld, int row0, float mul) {
    for (int i = tid; i < TQ * (DH / 4); i += kAbThreads) {
      int r = i / (DH / 4), c4 = i % (DH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < t) v = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + c4 * 4);
      v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
      *reinterpret_cast<float4*>(dst + r * LD + c4 * 4) = v;
    }
  };

  if (MODE == 0) {
    load_tile(F1, qbase, ld3, r0, scale);
    load_tile(F2, dobase, d, r0, 1.f);
    if (tid < TQ) {
      int q = r0 + tid;
      s_lse[tid] = q < t ? lse_b[q] : 0.f;
      s_del[tid] = q < t ? del_b[q] : 0.f;
    }
  } else {
    load_tile(F1, kbase, ld3, r0, 1.f);
    load_tile(F2, vbase, ld3, r0, 1.f);
    int valid = 0;
    if (tid < TQ) {
      int key = r0 + tid;
      valid = key < t && !(kpm_b && kpm_b[key]);
      s_kvalid[tid] = valid;
    }
    if (!__syncthreads_or(valid)) {  // only PAD keys in this tile: their gradients are zero
      for (int i = tid; i < TQ * (DH / 4); i += kAbThreads) {
        int r = i / (DH / 4), c4 = i % (DH / 4);
        if (r0 + r < t) {
          float* o = dqkv + ((size_t)b * t + r0 + r) * ld3 + (size_t)h * DH + c4 * 4;
          *reinterpret_cast<float4*>(o + d) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(o + 2 * d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      return;
    }
  }

  float4 acc1[R][NG], acc2[MODE == 1 ? R : 1][MODE == 1 ? NG : 1];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      acc1[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 1) acc2[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

  for (int c0 = 0; c0 < t; c0 += TQ) {
    __syncthreads();  // previous iteration finished with G1/G2/Ps/Ds and the per-tile vectors
    if (MODE == 0) {
      int valid = 0;
      if (tid < TQ) {
        int key = c0 + tid;
        valid = key < t && !(kpm_b && kpm_b[key]);
        s_kvalid[tid] = valid;
      }
      if (!__syncthreads_or(valid)) continue;  // tile of PAD keys only
      load_tile(G1, kbase, ld3, c0, 1.f);
      load_tile(G2, vbase, ld3, c0, 1.f);
    } else {
      load_tile(G1, qbase, ld3, c0, scale);
      load_tile(G2, dobase, d, c0, 1.f);
      if (tid < TQ) {
        int q = c0 + tid;
        s_lse[tid] = q < t ? lse_b[q] : 0.f;
        s_del[tid] = q < t ? del_b[q] : 0.f;
      }
    }
    __syncthreads();

    float x1[R][R], x2[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) x1[i][j] = x2[i][j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < DH; c += 4) {
      float4 fa[R], ga[R], fb[R], gb[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        fa[i] = *reinterpret_cast<const float4*>(F1 + (ty + 16 * i) * LD + c);
        fb[i] = *reinterpret_cast<const float4*>(F2 + (ty + 16 * i) * LD + c);
        ga[i] = *reinterpret_cast<const float4*>(G1 + (tx + 16 * i) * LD + c);
        gb[i] = *reinterpret_cast<const float4*>(G2 + (tx + 16 * i) * LD + c);
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          x1[i][j] = fmaf(fa[i].x, ga[j].x, x1[i][j]);
          x1[i][j] = fmaf(fa[i].y, ga[j].y, x1[i][j]);
          x1[i][j] = fmaf(fa[i].z, ga[j].z, x1[i][j]);
          x1[i][j] = fmaf(fa[i].w, ga[j].w, x1[i][j]);
          x2[i][j] = fmaf(fb[i].x, gb[j].x, x2[i][j]);
          x2[i][j] = fmaf(fb[i].y, gb[j].y, x2[i][j]);
          x2[i][j] = fmaf(fb[i].z, gb[j].z, x2[i][j]);
          x2[i][j] = fmaf(fb[i].w, gb[j].w, x2[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int rr = ty + 16 * i, cc = tx + 16 * j;
        // (query, key) of this element and their tile-local indices
        const int ql = MODE == 0 ? rr : cc, kl = MODE == 0 ? cc : rr;
        const int q = (MODE == 0 ? r0 : c0) + ql;
        const bool ok = s_kvalid[kl] && q < t;
        const float p = ok ? expf(x1[i][j] - s_lse[ql]) : 0.f;
        const float ds = p * (x2[i][j] - s_del[ql]);
        Ds[rr * LP + cc] = ds;
        if (MODE == 1) Ps[rr * LP + cc] = p;
      }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < TQ; ++j) {
      float dsv[R], pv[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        dsv[i] = Ds[(ty + 16 * i) * LP + j];
        if (MODE == 1) pv[i] = Ps[(ty + 16 * i) * LP + j];
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 v1 = *reinterpret_cast<const float4*>(G1 + j * LD + g * 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          acc1[i][g].x = fmaf(dsv[i], v1.x, acc1[i][g].x);
          acc1[i][g].y = fmaf(dsv[i], v1.y, acc1[i][g].y);
          acc1[i][g].z = fmaf(dsv[i], v1.z, acc1[i][g].z);
          acc1[i][g].w = fmaf(dsv[i], v1.w, acc1[i][g].w);
        }
        if (MODE == 1) {
          const float4 v2 = *reinterpret_cast<const float4*>(G2 + j * LD + g * 64 + tx * 4);
#pragma unroll
          for (int i = 0; i < R; ++i) {
            acc2[i][g].x = fmaf(pv[i], v2.x, acc2[i][g].x);
            acc2[i][g].y = fmaf(pv[i], v2.y, acc2[i][g].y);
            acc2[i][g].z = fmaf(pv[i], v2.z, acc2[i][g].z);
            acc2[i][g].w = fmaf(pv[i], v2.w, acc2[i][g].w);
          }
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int row = r0 + ty + 16 * i;
    if (row >= t) continue;
    float* o = dqkv + ((size_t)b * t + row) * ld3 + (size_t)h * DH;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (MODE == 0) {
        float4 v = acc1[i][g];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *reinterpret_cast<float4*>(o + g * 64 + tx * 4) = v;
      } else {
        *reinterpret_cast<float4*>(o + d + g * 64 + tx * 4) = acc1[i][g];      // dK = dS^T . (q*scale)
        *reinterpret_cast<float4*>(o + 2 * d + g * 64 + tx * 4) = acc2[i][g];  // dV = P^T . dO
      }
    }
  }
}

template <int DH, int TQ>
static int launch_attention_bwd(const float* qkv, const float* dctx, const float* lse, const float* delta,
                                const uint8_t* kpm, float* dqkv, int batch, int t, int d, int nhead, cudaStream_t s) {
  const size_t smem = sizeof(float) * (4 * TQ * (DH + kAbPad) + 2 * TQ * (TQ + kAbPad) + 2 * TQ) + sizeof(int) * TQ;
  auto k0 = attention_bwd_f32_kernel<DH, TQ, 0>;
  auto k1 = attention_bwd_f32_kernel<DH, TQ, 1>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("attention_bwd: cannot reserve %zu bytes of shared memory", smem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(ceil_div(t, TQ), nhead, batch);
  const float scale = 1.0f / sqrtf((float)DH);
  k0<<<grid, kAbThreads, smem, s>>>(qkv, dctx, lse, delta, kpm, dqkv, t, d, scale);
  k1<<<grid, kAbThreads, smem, s>>>(qkv, dctx, lse, delta, kpm, dqkv, t, d, scale);
  LFS2_CHECK_LAUNCH("attention_bwd");
  return LFS2_OK;
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

long long lfs2_attention_bwd_workspace_bytes(int batch, int t, int nhead) {
  return batch > 0 && t > 0 && nhead > 0 ? (long long)batch * t * nhead * (long long)sizeof(float) : 0;
}

int lfs2_attention_bwd(const float* qkv, const float* ctx, const float* dctx, const float* lse,
                       const uint8_t* key_padding_mask, float* dqkv, void* workspace, int batch, int t, int d,
                       int nhead, void* stream) {
  LFS2_REQUIRE(qkv && ctx && dctx && lse && dqkv && workspace, LFS2_ERR_INVALID_ARG, "attention_bwd: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && nhead > 0 && d % nhead == 0, LFS2_ERR_INVALID_ARG,
               "attention_bwd: bad shape");
  LFS2_REQUIRE(batch <= 65535 && nhead <= 65535, LFS2_ERR_UNSUPPORTED, "attention_bwd: batch/nhead > 65535");
  LFS2_REQUIRE(aligned16(qkv) && aligned16(ctx) && aligned16(dctx) && aligned16(dqkv), LFS2_ERR_INVALID_ARG,
               "attention_bwd: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int dh = d / nhead;
  LFS2_REQUIRE(dh % 4 == 0, LFS2_ERR_UNSUPPORTED, "attention_bwd: head_dim %d", dh);
  float* delta = reinterpret_cast<float*>(workspace);
  attn_delta_kernel<<<ceil_div((long long)batch * t * nhead * 32, 256), 256, 0, s>>>(
      (const float4*)dctx, (const float4*)ctx, delta, batch, t, nhead, dh / 4);
  LFS2_CHECK_LAUNCH("attn_delta");
  switch (dh) {
    case 64: return launch_attention_bwd<64, 64>(qkv, dctx, lse, delta, key_padding_mask, dqkv, batch, t, d, nhead, s);
    case 128: return launch_attention_bwd<128, 64>(qkv, dctx, lse, delta, key_padding_mask, dqkv, batch, t, d, nhead, s);
    case 384: return launch_attention_bwd<384, 32>(qkv, dctx, lse, delta, key_padding_mask, dqkv, batch, t, d, nhead, s);
    default:
      set_error("attention_bwd: head_dim %d not supported (64, 128, 384)", dh);
      return LFS2_ERR_UNSUPPORTED;
  }
}

}  // extern "C"
