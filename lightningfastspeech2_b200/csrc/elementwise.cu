// HBM-bound glue kernels of the mel-generation path: front end (embedding + positional
// encoding + speaker term), residual LayerNorm, depthwise Conv1d, predictor head, duration
// rounding + zero-duration guard, bucketize + embedding add.  All are pure streaming
// kernels: 16-byte coalesced accesses, warp-shuffle reductions, no shared-memory staging
// (no reuse to exploit beyond L1/L2).
#include <math.h>
#include <stdarg.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace lfs2 {
namespace tc {
// dwconv_tma.cu: 0 = launched, 1 = not its case (fall back to dwconv1d_k_kernel), < 0 = error
int launch_dwconv_tma(const void* x_hi, const void* x_lo, const float* wt, const float* bias, float* out, void* out_hi,
                      void* out_lo, void* out_f16, int batch, int t, int d, int ksize, const int* row_limit,
                      int limit_extra, cudaStream_t s);
}  // namespace tc
}  // namespace lfs2

namespace lfs2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------
// spk[b, j] = relu(dot(w[j, :], dvec[b, :]) + bias[j]); one warp per output element
__global__ void speaker_proj_kernel(const float* __restrict__ dvec, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ spk, int batch,
                                    int in_dim, int d) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= batch * d) return;
  int b = warp / d, j = warp % d;
  const float* wr = w + (size_t)j * in_dim;
  const float* xr = dvec + (size_t)b * in_dim;
  float acc = 0.f;
  for (int i = lane; i < in_dim; i += 32) acc = fmaf(wr[i], xr[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) spk[warp] = fmaxf(acc + bias[j], 0.f);
}

// x[b,t,:] = emb[phones[b,t],:] + pe[t,:] + spk[b,:]; one thread per float4
__global__ void embed_pe_spk_kernel(const int64_t* __restrict__ phones, const float4* __restrict__ emb,
                                    const float4* __restrict__ pe, const float4* __restrict__ spk,
                                    float4* __restrict__ x, uint8_t* __restrict__ src_mask, int batch, int t,
                                    int d4, int vocab) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)batch * t * d4;
  if (i >= total) return;
  int c = (int)(i % d4);
  size_t row = i / d4;
  int tt = (int)(row % t);
  int b = (int)(row / t);
  long long ph = phones[row];
  if (c == 0) src_mask[row] = (ph == 0);
  ph = ph < 0 ? 0 : (ph >= vocab ? vocab - 1 : ph);  // never read out of the table
  float4 e = emb[(size_t)ph * d4 + c];
  float4 p = pe[(size_t)tt * d4 + c];
  float4 s = spk[(size_t)b * d4 + c];
  // same association as the reference: (emb + pe) + spk
  float4 o;
  o.x = (e.x + p.x) + s.x;
  o.y = (e.y + p.y) + s.y;
  o.z = (e.z + p.z) + s.z;
  o.w = (e.w + p.w) + s.w;
  x[i] = o;
}

__global__ void add_pe_spk_kernel(float4* __restrict__ x, const float4* __restrict__ pe,
                                  const float4* __restrict__ spk, int batch, int t, int d4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)batch * t * d4;
  if (i >= total) return;
  int c = (int)(i % d4);
  size_t row = i / d4;
  int tt = (int)(row % t);
  int b = (int)(row / t);
  float4 v = x[i];
  float4 p = pe[(size_t)tt * d4 + c];
  float4 s = spk[(size_t)b * d4 + c];
  v.x = (v.x + p.x) + s.x;
  v.y = (v.y + p.y) + s.y;
  v.z = (v.z + p.z) + s.z;
  v.w = (v.w + p.w) + s.w;
  x[i] = v;
}

// ---------------------------------------------------------------------------------------
// out[m,:] = LN(x[m,:] + y[m,:]); one warp per row, row kept in registers (d <= 1024),
// two-pass mean / centred variance like torch's CPU kernel.
constexpr int kLnMaxVec = 8;  // float4 per lane -> d <= 1024
__global__ void add_layernorm_kernel(const float4* __restrict__ x, const float4* __restrict__ y,
                                     const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                     float4* __restrict__ out, float4* __restrict__ z_out,
                                     float2* __restrict__ stats, int m, int d4, float eps, DropSite drop) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= m) return;
  const float4* xr = x + (size_t)row * d4;
  const float4* yr = y ? y + (size_t)row * d4 : nullptr;
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    int c = lane + 32 * i;
    if (c < d4) {
      float4 a = xr[c];
      if (yr) {
        float4 b = yr[c];
        if (drop.threshold) {  // fused dropout of the branch (dropout1 / dropout2, model.py:114-115): z = x + drop(y)
          const float4 k = dropout_scale4((size_t)row * d4 + c, drop.threshold, drop.inv_keep, drop.key, drop.site);
          b.x *= k.x; b.y *= k.y; b.z *= k.z; b.w *= k.w;
        }
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[i] = a;
      if (z_out) z_out[(size_t)row * d4 + c] = a;  // pre-norm sum, saved for the backward pass
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  float inv_d = 1.f / (float)(d4 * 4);
  float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    int c = lane + 32 * i;
    if (c < d4) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + dd * dd);
    }
  }
  float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
  float4* orow = out + (size_t)row * d4;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    int c = lane + 32 * i;
    if (c < d4) {
      float4 g = gamma[c], bt = beta[c], o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x;
      o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z;
      o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      orow[c] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------
// depthwise conv, channels-last: out[b,t,c] = bias[c] + sum_j wt[j,c] * x[b,t+j-h,c]
// thread = 4 channels x kDwT consecutive frames (sliding window held in registers)
constexpr int kDwT = 8;
__device__ __forceinline__ float4 planes_to_f4(uint2 h, uint2 l) {
  float4 r;
  r.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  r.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  r.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  r.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return r;
}

template <bool IN_PLANES>
__global__ void dwconv1d_kernel(const float4* __restrict__ x, const uint2* __restrict__ x_hi,
                                const uint2* __restrict__ x_lo, const float4* __restrict__ wt,
                                const float4* __restrict__ bias, float4* __restrict__ out, uint2* __restrict__ out_hi,
                                uint2* __restrict__ out_lo, int batch, int t, int d4, int ksize, int nchunk) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)batch * nchunk * d4;
  if (i >= total) return;
  int c = (int)(i % d4);
  int chunk = (int)((i / d4) % nchunk);
  int b = (int)(i / ((size_t)d4 * nchunk));
  int t0 = chunk * kDwT;
  int h = (ksize - 1) / 2;
  const size_t xoff = (size_t)b * t * d4 + c;
  float4 acc[kDwT];
  float4 bz = bias[c];
#pragma unroll
  for (int o = 0; o < kDwT; ++o) acc[o] = bz;
  // input frame t0 - h + j contributes to output o = j - tap for tap in [0, ksize)
  for (int j = 0; j < kDwT + ksize - 1; ++j) {
    int ti = t0 - h + j;
    if (ti < 0 || ti >= t) continue;
    float4 xv;
    if (IN_PLANES) xv = planes_to_f4(x_hi[xoff + (size_t)ti * d4], x_lo[xoff + (size_t)ti * d4]);
    else xv = x[xoff + (size_t)ti * d4];
#pragma unroll
    for (int o = 0; o < kDwT; ++o) {
      int tap = j - o;
      if (tap >= 0 && tap < ksize) {
        const float4 w = wt[(size_t)tap * d4 + c];
        fma4(acc[o], w, xv);
      }
    }
  }
  const size_t obase = (size_t)b * t * d4 + c;
#pragma unroll
  for (int o = 0; o < kDwT; ++o)
    if (t0 + o < t) {
      const size_t oi = obase + (size_t)(t0 + o) * d4;
      if (out) out[oi] = acc[o];
      if (out_hi) {  // bf16 hi/lo planes (x = hi + lo), the operand format of the tcgen05 GEMMs
        __nv_bfloat16 h[4], l[4];
        const float v[4] = {acc[o].x, acc[o].y, acc[o].z, acc[o].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          h[q] = __float2bfloat16_rn(v[q]);
          l[q] = __float2bfloat16_rn(v[q] - __bfloat162float(h[q]));
        }
        out_hi[oi] = make_uint2((uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
                                (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
        out_lo[oi] = make_uint2((uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
                                (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
      }
    }
}

// ---------------------------------------------------------------------------------------
// out[m] = mask[m] ? 0 : dot(z[m,:], w) + bias ; one warp per row
__global__ void rowdot_mask_kernel(const float4* __restrict__ z, const float4* __restrict__ w,
                                   const float* __restrict__ bias, const uint8_t* __restrict__ mask,
                                   float* __restrict__ out, int m, int f4) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= m) return;
  const float4* zr = z + (size_t)row * f4;
  float acc = 0.f;
  for (int c = lane; c < f4; c += 32) {
    float4 a = zr[c], b = w[c];
    acc += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = (mask && mask[row]) ? 0.f : acc + bias[0];
}

// ---------------------------------------------------------------------------------------
// inference durations + zero-duration guard; one CTA per utterance
__global__ void duration_round_guard_kernel(const float* __restrict__ log_dur, const uint8_t* __restrict__ src_mask,
                                            int32_t* __restrict__ dur, int tp) {
  int b = blockIdx.x;
  const float* p = log_dur + (size_t)b * tp;
  const uint8_t* mk = src_mask + (size_t)b * tp;
  int32_t* o = dur + (size_t)b * tp;
  long long total = 0;
  int nvalid = 0;
  for (int i = threadIdx.x; i < tp; i += blockDim.x) {
    float v = rintf(expf(p[i]) - 1.f);  // torch.round = round-half-even
    v = fmaxf(v, 0.f);                   // clamp(min=0); NaN -> 0 differs from torch only for NaN inputs
    int32_t di = v >= 2147483520.f ? 2147483647 : (int32_t)v;
    o[i] = di;
    if (!mk[i]) {
      total += di;
      ++nvalid;
    }
  }
  __shared__ long long s_total[32];
  __shared__ int s_nvalid[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, off);
    nvalid += __shfl_xor_sync(0xffffffffu, nvalid, off);
  }
  if (lane == 0) {
    s_total[w] = total;
    s_nvalid[w] = nvalid;
  }
  __syncthreads();
  total = 0;
  nvalid = 0;
  for (int i = 0; i < (blockDim.x >> 5); ++i) {
    total += s_total[i];
    nvalid += s_nvalid[i];
  }
  if (total <= (long long)(nvalid / 2)) {
    for (int i = threadIdx.x; i < tp; i += blockDim.x)
      if (!mk[i]) o[i] = 1;
  }
}

// ---------------------------------------------------------------------------------------
// bucketize (right=False, torch's lower-bound loop incl. its NaN behaviour) + embedding add
__global__ void bucket_embed_add_kernel(const float4* x_in, float4* x, const float* __restrict__ val, float stdv,
                                        float meanv, const float* __restrict__ bins, int nb,
                                        const float4* __restrict__ emb, const int64_t* __restrict__ idx_forced,
                                        int64_t* __restrict__ idx_out, float4* __restrict__ acc, int acc_mode,
                                        int m, int d4) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= m) return;
  int idx;
  if (idx_forced) {
    idx = (int)idx_forced[row];
  } else {
    // two roundings (mul then add), exactly like `prediction * std + mean` in torch
    float v = __fadd_rn(__fmul_rn(val[row], stdv), meanv);
    int lo = 0, hi = nb;
    while (lo < hi) {
      int mid = lo + ((hi - lo) >> 1);
      if (!(bins[mid] >= v)) lo = mid + 1;
      else hi = mid;
    }
    idx = lo;
  }
  if (lane == 0 && idx_out) idx_out[row] = idx;
  const float4* e = emb + (size_t)idx * d4;
  float4* xr = x + (size_t)row * d4;
  const float4* xi = x_in + (size_t)row * d4;  // == xr for the in-place form
  float4* ar = acc ? acc + (size_t)row * d4 : nullptr;
  for (int c = lane; c < d4; c += 32) {
    float4 ev = e[c], xv = xi[c];
    xv.x += ev.x; xv.y += ev.y; xv.z += ev.z; xv.w += ev.w;
    xr[c] = xv;
    if (acc_mode == 1) {
      ar[c] = ev;
    } else if (acc_mode == 2) {
      float4 av = ar[c];
      av.x += ev.x; av.y += ev.y; av.z += ev.z; av.w += ev.w;
      ar[c] = av;
    }
  }
}

__device__ __forceinline__ void split_pack2_ew(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}

// Decoder input in ONE pass: y = ((x + emb[bucket(val)]) + pe[t]) + spk[b], written only as the operand planes of the
// first decoder block's tensor-core GEMMs (bf16 hi/lo, optionally the fp16 plane of the 2-pass recipe) -- the last
// frame-level variance encoder's embedding add (model.py:263-266), the positional / speaker add (fastspeech2.py:716-
// 721) and lfs2_split_bf16 fused: 4 + 6 bytes per element instead of 26.  Same operation order as the three kernels it
// replaces, so the planes are bit-identical.  emb == NULL: no bucket term.  One warp per row.
__global__ void decoder_input_planes_kernel(const float4* __restrict__ x, const float* __restrict__ val, float stdv,
                                            float meanv, const float* __restrict__ bins, int nb,
                                            const float4* __restrict__ emb, const int64_t* __restrict__ idx_forced,
                                            int64_t* __restrict__ idx_out, float4* __restrict__ acc, int acc_mode,
                                            const float4* __restrict__ pe, const float4* __restrict__ spk, int t,
                                            uint2* __restrict__ out_hi, uint2* __restrict__ out_lo,
                                            uint2* __restrict__ out_f16, int m, int d4) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  const float4* e = nullptr;
  if (emb) {
    int idx;
    if (idx_forced) {
      idx = (int)idx_forced[row];
    } else {
      float v = __fadd_rn(__fmul_rn(val[row], stdv), meanv);
      int lo = 0, hi = nb;
      while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (!(bins[mid] >= v)) lo = mid + 1;
        else hi = mid;
      }
      idx = lo;
    }
    if (lane == 0 && idx_out) idx_out[row] = idx;
    e = emb + (size_t)idx * d4;
  }
  const float4* xr = x + (size_t)row * d4;
  const float4* pr = pe + (size_t)(row % t) * d4;
  const float4* sr = spk + (size_t)(row / t) * d4;
  float4* ar = acc ? acc + (size_t)row * d4 : nullptr;
  for (int c = lane; c < d4; c += 32) {
    float4 xv = xr[c];
    if (e) {
      const float4 ev = e[c];
      xv.x += ev.x; xv.y += ev.y; xv.z += ev.z; xv.w += ev.w;
      if (acc_mode == 1) {
        ar[c] = ev;
      } else if (acc_mode == 2) {
        float4 av = ar[c];
        av.x += ev.x; av.y += ev.y; av.z += ev.z; av.w += ev.w;
        ar[c] = av;
      }
    }
    const float4 p = pr[c], sp = sr[c];
    xv.x = (xv.x + p.x) + sp.x;
    xv.y = (xv.y + p.y) + sp.y;
    xv.z = (xv.z + p.z) + sp.z;
    xv.w = (xv.w + p.w) + sp.w;
    const size_t o = (size_t)row * d4 + c;
    uint2 h, l;
    split_pack2_ew(xv.x, xv.y, h.x, l.x);
    split_pack2_ew(xv.z, xv.w, h.y, l.y);
    out_hi[o] = h;
    out_lo[o] = l;
    if (out_f16) {
      uint2 f;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(f.x) : "f"(xv.y), "f"(xv.x));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(f.y) : "f"(xv.w), "f"(xv.z));
      out_f16[o] = f;
    }
  }
}

// PriorEmbedding (reference model.py:146-164): out[b,:] = relu(emb[bucketize(prior[b], bins), :]); one warp per utterance
__global__ void prior_embed_kernel(const float* __restrict__ prior, const float* __restrict__ bins, int nb,
                                   const float4* __restrict__ emb, float4* __restrict__ out, int64_t* __restrict__ idx_out,
                                   int batch, int d4) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= batch) return;
  const float v = prior[b];
  int lo = 0, hi = nb;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if (!(bins[mid] >= v)) lo = mid + 1;
    else hi = mid;
  }
  if (lane == 0 && idx_out) idx_out[b] = lo;
  for (int c = lane; c < d4; c += 32) {
    float4 e = emb[(size_t)lo * d4 + c];
    e.x = fmaxf(e.x, 0.f); e.y = fmaxf(e.y, 0.f); e.z = fmaxf(e.z, 0.f); e.w = fmaxf(e.w, 0.f);
    out[(size_t)b * d4 + c] = e;
  }
}

// Depthwise conv specialised on the kernel size, shared-memory tiled: one CTA = 64 (K <= 9) or
// 32 output frames x 256 channels of one utterance, two CTAs per SM so one CTA's load phase
// overlaps the other's arithmetic.  The (tile + K - 1) input rows are staged once in
// shared memory as fp32 (coalesced 8/16-byte loads, all independent -> deep memory-level
// parallelism; rows outside [0, T) are zeros = Conv1d's "same" padding), then each thread
// slides NT consecutive frames of its 4 channels through registers: the K tap weights live in
// registers and the (input row, output row) -> tap mapping is resolved at compile time, so the
// inner loop is LDS.128 + FFMA only.  Results leave as fp32 and/or bf16 hi/lo planes.

constexpr int kDwCh4 = 64;    // float4 channel groups per CTA (256 channels)

template <int K, int NT, int kDwTile, bool IN_PLANES>
__global__ void __launch_bounds__(kDwCh4 * (kDwTile / NT), 2)
dwconv1d_k_kernel(const float4* __restrict__ x, const uint2* __restrict__ x_hi, const uint2* __restrict__ x_lo,
                  const float4* __restrict__ wt, const float4* __restrict__ bias, float4* __restrict__ out,
                  uint2* __restrict__ out_hi, uint2* __restrict__ out_lo, uint2* __restrict__ out_f16, int t, int d4,
                  const int* __restrict__ row_limit, int limit_extra) {
  extern __shared__ float4 xs[];  // [kDwTile + K - 1][kDwCh4]
  constexpr int H = (K - 1) / 2;
  constexpr int kRows = kDwTile + K - 1;
  constexpr int kThreads = kDwCh4 * (kDwTile / NT);
  const int t0 = blockIdx.x * kDwTile;
  const int cb = blockIdx.y * kDwCh4;  // first float4 channel group of this CTA
  const int b = blockIdx.z;
  // rows the caller does not need (same 128-row granularity as the tensor-core GEMM that consumes the result);
  // the rows past the last kept 128-row group are nobody's output, so they are read as zeros, not from memory
  int t_in = t;
  if (row_limit) {
    const int lim = __ldg(row_limit + b) + limit_extra;
    if ((t0 & ~127) >= lim) return;
    t_in = min(t, (lim + 127) & ~127);
  }
  const int nch = min(kDwCh4, d4 - cb);
  const size_t base = (size_t)b * t * d4 + cb;

  // stage the input rows: all global loads of a thread are issued before the first use
  constexpr int kIters = (kRows * kDwCh4 + kThreads - 1) / kThreads;
  {
    const int c = threadIdx.x % kDwCh4;            // kThreads is a multiple of kDwCh4
    const int row0 = threadIdx.x / kDwCh4;
    constexpr int kRowStep = kThreads / kDwCh4;
    uint2 rh[kIters], rl[kIters];
    float4 rf[IN_PLANES ? 1 : kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int row = row0 + it * kRowStep;
      const int ti = t0 - H + row;
      const bool ok = row < kRows && ti >= 0 && ti < t_in && c < nch;
      const size_t gi = base + (size_t)(ok ? ti : 0) * d4 + (ok ? c : 0);
      if (IN_PLANES) {
        rh[it] = ok ? x_hi[gi] : make_uint2(0u, 0u);
        rl[it] = ok ? x_lo[gi] : make_uint2(0u, 0u);
      } else {
        rf[it] = ok ? x[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int row = row0 + it * kRowStep;
      if (row < kRows) xs[row * kDwCh4 + c] = IN_PLANES ? planes_to_f4(rh[it], rl[it]) : rf[it];
    }
  }
  __syncthreads();

  const int c = threadIdx.x % kDwCh4, tg = threadIdx.x / kDwCh4;
  if (c >= nch) return;
  float4 w[K];
#pragma unroll
  for (int j = 0; j < K; ++j) w[j] = wt[(size_t)j * d4 + cb + c];
  const float4 bz = bias[cb + c];
  float4 acc[NT];
#pragma unroll
  for (int o = 0; o < NT; ++o) acc[o] = bz;
  const float4* xr = xs + (size_t)(tg * NT) * kDwCh4 + c;
#pragma unroll
  for (int j = 0; j < NT + K - 1; ++j) {
    const float4 xv = xr[j * kDwCh4];
#pragma unroll
    for (int o = 0; o < NT; ++o) {
      const int tap = j - o;  // compile-time after unrolling
      if (tap >= 0 && tap < K) {
        fma4(acc[o], w[tap], xv);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NT; ++o) {
    const int to = t0 + tg * NT + o;
    if (to < t) {
      const size_t oi = base + (size_t)to * d4 + c;
      if (out) out[oi] = acc[o];
      if (out_hi) {
        uint2 h, l;
        split_pack2_ew(acc[o].x, acc[o].y, h.x, l.x);
        split_pack2_ew(acc[o].z, acc[o].w, h.y, l.y);
        out_hi[oi] = h;
        out_lo[oi] = l;
      }
      if (out_f16) {  // ONE fp16 plane (saturating): the activation operand of a 2-pass GEMM
        uint2 h;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(acc[o].y), "f"(acc[o].x));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(acc[o].w), "f"(acc[o].z));
        out_f16[oi] = h;
      }
    }
  }
}

template <int K, int NT, int kDwTile>
static int launch_dwconv_k(const float* x, const void* x_hi, const void* x_lo, const float* wt, const float* bias,
                           float* out, void* out_hi, void* out_lo, void* out_f16, int batch, int t, int d,
                           const int* row_limit, int limit_extra, cudaStream_t s) {
  constexpr int kThreads = kDwCh4 * (kDwTile / NT);
  constexpr int kSmem = (kDwTile + K - 1) * kDwCh4 * 16;
  dim3 grid(ceil_div(t, kDwTile), ceil_div(d / 4, kDwCh4), batch);
  static bool configured = false;
  auto kf = dwconv1d_k_kernel<K, NT, kDwTile, false>;
  auto kp = dwconv1d_k_kernel<K, NT, kDwTile, true>;
  if (!configured) {
    if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) {
      set_error("dwconv1d: cannot reserve %d bytes of shared memory", kSmem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  if (x)
    kf<<<grid, kThreads, kSmem, s>>>((const float4*)x, nullptr, nullptr, (const float4*)wt, (const float4*)bias,
                                     (float4*)out, (uint2*)out_hi, (uint2*)out_lo, (uint2*)out_f16, t, d / 4, row_limit,
                                     limit_extra);
  else
    kp<<<grid, kThreads, kSmem, s>>>(nullptr, (const uint2*)x_hi, (const uint2*)x_lo, (const float4*)wt,
                                     (const float4*)bias, (float4*)out, (uint2*)out_hi, (uint2*)out_lo, (uint2*)out_f16, t,
                                     d / 4, row_limit, limit_extra);
  return LFS2_OK;
}

// out = LayerNorm(x + y) with the residual stream x and the result as bf16 hi/lo planes, y fp32 (a GEMM's fp32 output):
// the d != 256 FFTBlock (no LayerNorm epilogue in the GEMM) stays in plane form from block to block -- no merge / split
// passes around the LayerNorm.  One warp per row, statistics in fp32 exactly like add_layernorm_kernel.
__global__ void add_layernorm_planes_kernel(const uint2* __restrict__ x_hi, const uint2* __restrict__ x_lo,
                                            const float4* __restrict__ y, const float4* __restrict__ gamma,
                                            const float4* __restrict__ beta, uint2* __restrict__ out_hi,
                                            uint2* __restrict__ out_lo, int m, int d4, float eps, int t,
                                            const int* __restrict__ row_limit, int limit_extra) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  // rows of the 128-row groups that start at or after row_limit[b] + extra are nobody's input (PAD-row skipping)
  if (row_limit && ((row % t) & ~127) >= __ldg(row_limit + row / t) + limit_extra) return;
  const size_t base = (size_t)row * d4;
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < d4) {
      float4 a = planes_to_f4(x_hi[base + c], x_lo[base + c]);
      if (y) {
        const float4 b = y[base + c];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[i] = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float inv_d = 1.f / (float)(d4 * 4);
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < d4) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + dd * dd);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < d4) {
      const float4 g = gamma[c], bt = beta[c];
      const float ox = (v[i].x - mean) * rstd * g.x + bt.x, oy = (v[i].y - mean) * rstd * g.y + bt.y;
      const float oz = (v[i].z - mean) * rstd * g.z + bt.z, ow = (v[i].w - mean) * rstd * g.w + bt.w;
      uint2 h, l;
      split_pack2_ew(ox, oy, h.x, l.x);
      split_pack2_ew(oz, ow, h.y, l.y);
      out_hi[base + c] = h;
      out_lo[base + c] = l;
    }
  }
}

// x = hi + lo (fp32) from bf16 planes; one thread per 4 elements
__global__ void merge_planes_kernel(const uint2* __restrict__ hi, const uint2* __restrict__ lo, float4* __restrict__ out,
                                    size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) out[i] = planes_to_f4(hi[i], lo[i]);
}

// x[r, :] = 0 where mask[r] != 0; one warp per row, float4 stores
__global__ void zero_masked_rows_kernel(float4* __restrict__ x, const uint8_t* __restrict__ mask, long long rows, int w4) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows || !mask[r]) return;
  for (int c = threadIdx.x & 31; c < w4; c += 32) x[r * w4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_zero_masked_rows(float* x, const uint8_t* mask, long long rows, int width, void* stream) {
  LFS2_REQUIRE(x && mask, LFS2_ERR_INVALID_ARG, "zero_masked_rows: null pointer");
  if (rows == 0) return LFS2_OK;
  LFS2_REQUIRE(rows > 0 && width > 0 && width % 4 == 0, LFS2_ERR_UNSUPPORTED,
               "zero_masked_rows: width must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(x), LFS2_ERR_INVALID_ARG, "zero_masked_rows: x must be 16-byte aligned");
  zero_masked_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>((float4*)x, mask, rows, width / 4);
  LFS2_CHECK_LAUNCH("zero_masked_rows");
  return LFS2_OK;
}

int lfs2_merge_planes(const void* hi, const void* lo, float* out, long long n, void* stream) {
  LFS2_REQUIRE(hi && lo && out, LFS2_ERR_INVALID_ARG, "merge_planes: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "merge_planes: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(hi) && aligned16(lo) && aligned16(out), LFS2_ERR_INVALID_ARG,
               "merge_planes: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  merge_planes_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const uint2*)hi, (const uint2*)lo,
                                                                          (float4*)out, n4);
  LFS2_CHECK_LAUNCH("merge_planes");
  return LFS2_OK;
}

int lfs2_version(void) { return 100; }
const char* lfs2_last_error(void) { return lfs2::g_err; }

int lfs2_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return LFS2_ERR_CUDA;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

int lfs2_speaker_proj(const float* dvec, const float* w, const float* bias, float* spk, int batch, int in_dim,
                      int d, void* stream) {
  LFS2_REQUIRE(dvec && w && bias && spk, LFS2_ERR_INVALID_ARG, "speaker_proj: null pointer");
  LFS2_REQUIRE(batch > 0 && in_dim > 0 && d > 0, LFS2_ERR_INVALID_ARG, "speaker_proj: bad shape");
  int threads = 256;
  int blocks = ceil_div((long long)batch * d * 32, threads);
  speaker_proj_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(dvec, w, bias, spk, batch, in_dim, d);
  LFS2_CHECK_LAUNCH("speaker_proj");
  return LFS2_OK;
}

int lfs2_embed_pe_spk(const int64_t* phones, const float* emb, const float* pe, const float* spk, float* x,
                      uint8_t* src_mask, int batch, int t, int d, int vocab, void* stream) {
  LFS2_REQUIRE(phones && emb && pe && spk && x && src_mask, LFS2_ERR_INVALID_ARG, "embed_pe_spk: null pointer");
  LFS2_REQUIRE(batch > 0 && t > 0 && vocab > 0, LFS2_ERR_INVALID_ARG, "embed_pe_spk: bad shape");
  LFS2_REQUIRE(d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "embed_pe_spk: d=%d must be a multiple of 4", d);
  LFS2_REQUIRE(aligned16(emb) && aligned16(pe) && aligned16(spk) && aligned16(x), LFS2_ERR_INVALID_ARG,
               "embed_pe_spk: pointers must be 16-byte aligned");
  size_t total = (size_t)batch * t * (d / 4);
  embed_pe_spk_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      phones, (const float4*)emb, (const float4*)pe, (const float4*)spk, (float4*)x, src_mask, batch, t, d / 4,
      vocab);
  LFS2_CHECK_LAUNCH("embed_pe_spk");
  return LFS2_OK;
}

int lfs2_add_pe_spk(float* x, const float* pe, const float* spk, int batch, int t, int d, void* stream) {
  LFS2_REQUIRE(x && pe && spk, LFS2_ERR_INVALID_ARG, "add_pe_spk: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "add_pe_spk: d=%d must be a multiple of 4", d);
  LFS2_REQUIRE(aligned16(x) && aligned16(pe) && aligned16(spk), LFS2_ERR_INVALID_ARG,
               "add_pe_spk: pointers must be 16-byte aligned");
  size_t total = (size_t)batch * t * (d / 4);
  add_pe_spk_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((float4*)x, (const float4*)pe,
                                                                          (const float4*)spk, batch, t, d / 4);
  LFS2_CHECK_LAUNCH("add_pe_spk");
  return LFS2_OK;
}

int lfs2_add_layernorm(const float* x, const float* y, const float* gamma, const float* beta, float* out, int m,
                       int d, float eps, void* stream) {
  return lfs2_add_layernorm_train(x, y, gamma, beta, out, nullptr, nullptr, m, d, eps, 0.f, 0ull, 0u, stream);
}

int lfs2_add_layernorm_train(const float* x, const float* y, const float* gamma, const float* beta, float* out,
                             float* z_out, float* stats, int m, int d, float eps, float drop_p,
                             unsigned long long drop_seed, unsigned int drop_site, void* stream) {
  LFS2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LFS2_ERR_INVALID_ARG, "add_layernorm: dropout p must be in [0, 1)");
  LFS2_REQUIRE(x && gamma && beta && out, LFS2_ERR_INVALID_ARG, "add_layernorm: null pointer");
  LFS2_REQUIRE((!z_out || aligned16(z_out)) && (reinterpret_cast<uintptr_t>(stats) & 7u) == 0, LFS2_ERR_INVALID_ARG,
               "add_layernorm: z_out / stats misaligned");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0 && d <= 128 * kLnMaxVec, LFS2_ERR_UNSUPPORTED,
               "add_layernorm: d=%d must be a multiple of 4 and <= %d", d, 128 * kLnMaxVec);
  LFS2_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && aligned16(out) && (!y || aligned16(y)),
               LFS2_ERR_INVALID_ARG, "add_layernorm: pointers must be 16-byte aligned");
  int threads = 256;
  add_layernorm_kernel<<<ceil_div((long long)m * 32, threads), threads, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (const float4*)y, (const float4*)gamma, (const float4*)beta, (float4*)out, (float4*)z_out,
      (float2*)stats, m, d / 4, eps, make_drop_site(y ? drop_p : 0.f, drop_seed, drop_site));
  LFS2_CHECK_LAUNCH("add_layernorm");
  return LFS2_OK;
}

int lfs2_add_layernorm_planes(const void* x_hi, const void* x_lo, const float* y, const float* gamma, const float* beta,
                              void* out_hi, void* out_lo, int m, int d, float eps, void* stream) {
  return lfs2_add_layernorm_planes_limited(x_hi, x_lo, y, gamma, beta, out_hi, out_lo, 1, m, d, eps, nullptr, 0, stream);
}

int lfs2_add_layernorm_planes_limited(const void* x_hi, const void* x_lo, const float* y, const float* gamma,
                                      const float* beta, void* out_hi, void* out_lo, int batch, int t, int d, float eps,
                                      const int* row_limit, int limit_extra, void* stream) {
  LFS2_REQUIRE(x_hi && x_lo && gamma && beta && out_hi && out_lo, LFS2_ERR_INVALID_ARG, "add_layernorm_planes: null pointer");
  LFS2_REQUIRE(batch >= 0 && t >= 0 && (long long)batch * t <= 2147483647LL, LFS2_ERR_INVALID_ARG,
               "add_layernorm_planes: bad shape");
  const int m = batch * t;
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0 && d <= 128 * kLnMaxVec, LFS2_ERR_UNSUPPORTED,
               "add_layernorm_planes: d=%d must be a multiple of 4 and <= %d", d, 128 * kLnMaxVec);
  LFS2_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(gamma) && aligned16(beta) && aligned16(out_hi) &&
                   aligned16(out_lo) && (!y || aligned16(y)),
               LFS2_ERR_INVALID_ARG, "add_layernorm_planes: pointers must be 16-byte aligned");
  add_layernorm_planes_kernel<<<ceil_div((long long)m * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint2*)x_hi, (const uint2*)x_lo, (const float4*)y, (const float4*)gamma, (const float4*)beta,
      (uint2*)out_hi, (uint2*)out_lo, m, d / 4, eps, t, row_limit, limit_extra);
  LFS2_CHECK_LAUNCH("add_layernorm_planes");
  return LFS2_OK;
}

int lfs2_dwconv1d(const float* x, const float* wt, const float* bias, float* out, int batch, int t, int d,
                  int ksize, void* stream) {
  return lfs2_dwconv1d_planes(x, nullptr, nullptr, wt, bias, out, nullptr, nullptr, batch, t, d, ksize, stream);
}

int lfs2_dwconv1d_planes(const float* x, const void* x_hi, const void* x_lo, const float* wt, const float* bias,
                         float* out, void* out_hi, void* out_lo, int batch, int t, int d, int ksize, void* stream) {
  return lfs2_dwconv1d_planes_limited(x, x_hi, x_lo, wt, bias, out, out_hi, out_lo, batch, t, d, ksize, nullptr, 0, stream);
}

int lfs2_dwconv1d_planes_limited(const float* x, const void* x_hi, const void* x_lo, const float* wt, const float* bias,
                                 float* out, void* out_hi, void* out_lo, int batch, int t, int d, int ksize,
                                 const int* row_limit, int limit_extra, void* stream) {
  return lfs2_dwconv1d_planes_ex(x, x_hi, x_lo, wt, bias, out, out_hi, out_lo, nullptr, batch, t, d, ksize, row_limit,
                                 limit_extra, stream);
}

int lfs2_dwconv1d_planes_ex(const float* x, const void* x_hi, const void* x_lo, const float* wt, const float* bias,
                            float* out, void* out_hi, void* out_lo, void* out_f16, int batch, int t, int d, int ksize,
                            const int* row_limit, int limit_extra, void* stream) {
  LFS2_REQUIRE((x || (x_hi && x_lo)) && wt && bias && (out || out_hi || out_f16), LFS2_ERR_INVALID_ARG,
               "dwconv1d: null pointer");
  LFS2_REQUIRE(!out_f16 || (ksize <= 25 && aligned16(out_f16)), LFS2_ERR_UNSUPPORTED,
               "dwconv1d: the fp16 output plane needs an odd kernel size <= 25 and a 16-byte aligned pointer");
  LFS2_REQUIRE(!x || !x_hi, LFS2_ERR_INVALID_ARG, "dwconv1d: give the input as fp32 OR as planes");
  LFS2_REQUIRE(!out_hi == !out_lo, LFS2_ERR_INVALID_ARG, "dwconv1d: out_hi and out_lo go together");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "dwconv1d: d=%d must be a multiple of 4", d);
  LFS2_REQUIRE(ksize > 0 && ksize % 2 == 1, LFS2_ERR_UNSUPPORTED,
               "dwconv1d: kernel size %d must be odd ('same' padding is asymmetric otherwise)", ksize);
  LFS2_REQUIRE((!x || aligned16(x)) && (!x_hi || (aligned16(x_hi) && aligned16(x_lo))) && aligned16(wt) &&
                   aligned16(bias) && (!out || aligned16(out)) &&
                   (!out_hi || (aligned16(out_hi) && aligned16(out_lo))),
               LFS2_ERR_INVALID_ARG, "dwconv1d: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  LFS2_REQUIRE(batch <= 65535, LFS2_ERR_UNSUPPORTED, "dwconv1d: batch exceeds the grid limit");
  if (x_hi) {  // long kernels on plane-form input: tiles staged by TMA through a two-stage ring (dwconv_tma.cu)
    const int rc = tc::launch_dwconv_tma(x_hi, x_lo, wt, bias, out, out_hi, out_lo, out_f16, batch, t, d, ksize, row_limit,
                                         limit_extra, s);
    if (rc < 0) return rc;
    if (rc == 0) {
      LFS2_CHECK_LAUNCH("dwconv1d");
      return LFS2_OK;
    }
  }
#define LFS2_DW_CASE(K, TT)                                                                          \
  case K: {                                                                                          \
    int rc = launch_dwconv_k<K, TT, (K <= 9 ? 64 : 32)>(x, x_hi, x_lo, wt, bias, out, out_hi, out_lo, out_f16, batch, t, \
                                                        d, row_limit, limit_extra, s);                                \
    if (rc != LFS2_OK) return rc;                                                                    \
  } break;
  switch (ksize) {
    LFS2_DW_CASE(1, 16) LFS2_DW_CASE(3, 16) LFS2_DW_CASE(5, 16) LFS2_DW_CASE(7, 16) LFS2_DW_CASE(9, 16)
    LFS2_DW_CASE(11, 8) LFS2_DW_CASE(13, 8) LFS2_DW_CASE(15, 8) LFS2_DW_CASE(17, 8) LFS2_DW_CASE(19, 8)
    LFS2_DW_CASE(21, 8) LFS2_DW_CASE(23, 8) LFS2_DW_CASE(25, 8)
    default: {
      LFS2_REQUIRE(!row_limit, LFS2_ERR_UNSUPPORTED, "dwconv1d: row limits need an odd kernel size <= 25");
      int nchunk = ceil_div(t, kDwT);
      size_t total = (size_t)batch * nchunk * (d / 4);
      if (x)
        dwconv1d_kernel<false><<<ceil_div(total, 128), 128, 0, s>>>(
            (const float4*)x, nullptr, nullptr, (const float4*)wt, (const float4*)bias, (float4*)out, (uint2*)out_hi,
            (uint2*)out_lo, batch, t, d / 4, ksize, nchunk);
      else
        dwconv1d_kernel<true><<<ceil_div(total, 128), 128, 0, s>>>(
            nullptr, (const uint2*)x_hi, (const uint2*)x_lo, (const float4*)wt, (const float4*)bias, (float4*)out,
            (uint2*)out_hi, (uint2*)out_lo, batch, t, d / 4, ksize, nchunk);
    }
  }
#undef LFS2_DW_CASE
  LFS2_CHECK_LAUNCH("dwconv1d");
  return LFS2_OK;
}

int lfs2_rowdot_mask(const float* z, const float* w, const float* bias, const uint8_t* mask, float* out, int m,
                     int f, void* stream) {
  LFS2_REQUIRE(z && w && bias && out, LFS2_ERR_INVALID_ARG, "rowdot_mask: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(f > 0 && f % 4 == 0, LFS2_ERR_UNSUPPORTED, "rowdot_mask: f=%d must be a multiple of 4", f);
  LFS2_REQUIRE(aligned16(z) && aligned16(w), LFS2_ERR_INVALID_ARG, "rowdot_mask: pointers must be 16-byte aligned");
  int threads = 256;
  rowdot_mask_kernel<<<ceil_div((long long)m * 32, threads), threads, 0, (cudaStream_t)stream>>>(
      (const float4*)z, (const float4*)w, bias, mask, out, m, f / 4);
  LFS2_CHECK_LAUNCH("rowdot_mask");
  return LFS2_OK;
}

int lfs2_duration_round_guard(const float* log_dur, const uint8_t* src_mask, int32_t* dur, int batch, int tp,
                              void* stream) {
  LFS2_REQUIRE(log_dur && src_mask && dur, LFS2_ERR_INVALID_ARG, "duration_round_guard: null pointer");
  if (batch == 0 || tp == 0) return LFS2_OK;
  duration_round_guard_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(log_dur, src_mask, dur, tp);
  LFS2_CHECK_LAUNCH("duration_round_guard");
  return LFS2_OK;
}

int lfs2_prior_embed(const float* prior, const float* bins, int nbins, const float* emb, float* out, int64_t* idx_out,
                     int batch, int d, void* stream) {
  LFS2_REQUIRE(prior && bins && emb && out, LFS2_ERR_INVALID_ARG, "prior_embed: null pointer");
  if (batch == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && d > 0 && d % 4 == 0 && nbins >= 1, LFS2_ERR_UNSUPPORTED, "prior_embed: bad shape");
  LFS2_REQUIRE(aligned16(emb) && aligned16(out), LFS2_ERR_INVALID_ARG, "prior_embed: pointers must be 16-byte aligned");
  prior_embed_kernel<<<ceil_div((long long)batch * 32, 128), 128, 0, (cudaStream_t)stream>>>(
      prior, bins, nbins - 1, (const float4*)emb, (float4*)out, idx_out, batch, d / 4);
  LFS2_CHECK_LAUNCH("prior_embed");
  return LFS2_OK;
}

int lfs2_bucket_embed_add(float* x, const float* val, float stdv, float meanv, const float* bins, int nbins,
                          const float* emb, const int64_t* idx_forced, int64_t* idx_out, float* acc, int acc_mode,
                          int m, int d, void* stream) {
  return lfs2_bucket_embed_add_oop(x, x, val, stdv, meanv, bins, nbins, emb, idx_forced, idx_out, acc, acc_mode, m, d,
                                   stream);
}

int lfs2_bucket_embed_add_oop(const float* x_in, float* x, const float* val, float stdv, float meanv,
                              const float* bins, int nbins, const float* emb, const int64_t* idx_forced,
                              int64_t* idx_out, float* acc, int acc_mode, int m, int d, void* stream) {
  LFS2_REQUIRE(x_in && x && emb && (idx_forced || (val && bins)), LFS2_ERR_INVALID_ARG,
               "bucket_embed_add: null pointer");
  LFS2_REQUIRE(aligned16(x_in), LFS2_ERR_INVALID_ARG, "bucket_embed_add: x_in must be 16-byte aligned");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0 && nbins >= 1, LFS2_ERR_UNSUPPORTED, "bucket_embed_add: bad d/nbins");
  LFS2_REQUIRE(acc_mode == 0 || acc, LFS2_ERR_INVALID_ARG, "bucket_embed_add: acc_mode without acc");
  LFS2_REQUIRE(aligned16(x) && aligned16(emb) && (!acc || aligned16(acc)), LFS2_ERR_INVALID_ARG,
               "bucket_embed_add: pointers must be 16-byte aligned");
  int threads = 256;
  bucket_embed_add_kernel<<<ceil_div((long long)m * 32, threads), threads, 0, (cudaStream_t)stream>>>(
      (const float4*)x_in, (float4*)x, val, stdv, meanv, bins, nbins - 1, (const float4*)emb, idx_forced, idx_out,
      (float4*)acc,
      acc ? acc_mode : 0, m, d / 4);
  LFS2_CHECK_LAUNCH("bucket_embed_add");
  return LFS2_OK;
}

int lfs2_decoder_input_planes(const float* x, const float* val, float stdv, float meanv, const float* bins, int nbins,
                              const float* emb, const int64_t* idx_forced, int64_t* idx_out, float* acc, int acc_mode,
                              const float* pe, const float* spk, int batch, int t, int d, void* out_hi, void* out_lo,
                              void* out_f16, void* stream) {
  LFS2_REQUIRE(x && pe && spk && out_hi && out_lo, LFS2_ERR_INVALID_ARG, "decoder_input_planes: null pointer");
  LFS2_REQUIRE(!emb || idx_forced || (val && bins), LFS2_ERR_INVALID_ARG, "decoder_input_planes: the bucket term needs val + bins or forced indices");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && d % 4 == 0 && (!emb || nbins >= 1), LFS2_ERR_UNSUPPORTED, "decoder_input_planes: bad shape");
  LFS2_REQUIRE((long long)batch * t <= 0x7fffffffLL / 32, LFS2_ERR_UNSUPPORTED, "decoder_input_planes: too many rows");
  LFS2_REQUIRE(acc_mode == 0 || (acc && emb), LFS2_ERR_INVALID_ARG, "decoder_input_planes: acc_mode without acc / emb");
  LFS2_REQUIRE(aligned16(x) && aligned16(emb) && aligned16(acc) && aligned16(pe) && aligned16(spk) && aligned16(out_hi) &&
                   aligned16(out_lo) && aligned16(out_f16),
               LFS2_ERR_INVALID_ARG, "decoder_input_planes: pointers must be 16-byte aligned");
  const int m = batch * t, threads = 256;
  decoder_input_planes_kernel<<<ceil_div((long long)m * 32, threads), threads, 0, (cudaStream_t)stream>>>(
      (const float4*)x, val, stdv, meanv, bins, nbins - 1, (const float4*)emb, idx_forced, idx_out, (float4*)acc,
      acc ? acc_mode : 0, (const float4*)pe, (const float4*)spk, t, (uint2*)out_hi, (uint2*)out_lo, (uint2*)out_f16, m,
      d / 4);
  LFS2_CHECK_LAUNCH("decoder_input_planes");
  return LFS2_OK;
}

}  // extern "C"
