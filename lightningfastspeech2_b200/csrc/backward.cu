// Backward kernels of the mel-generation path (train-step config, SURVEY 8a "Backward notes"),
// exact fp32 on CUDA cores.  They are the gradient ground truth on the device (the tensor-core
// weight-gradient GEMM in gemm_tc2.cu is checked against gemm_tn here) and serve every
// shape.  Parameter gradients ACCUMULATE (+=) into caller-owned buffers -- the caller zeroes
// them once per step (the fused AdamW kernel does) -- so a parameter used twice (speaker
// projection at Tp and Tm) needs no extra pass; reductions over rows use fp32 atomics.
//
//   gemm_tn            dW (n,k)  += dY (m,n)^T . X (m,k)      [Conv1d taps: X rows shifted per utterance]
//   colsum             db (n)    += sum_m dY (m,n)
//   relu_bwd           dx = y > 0 ? dy : 0
//   layernorm_bwd      dz, dgamma +=, dbeta +=   from dy, pre-norm z and saved (mean, rstd)
//   dwconv1d_bwd_w     dwt (k,d) +=, dbias (d) +=  (input gradient = lfs2_dwconv1d with flipped taps)
//   length_regulate_bwd  dx[b,p] = sum of d_out rows of the frames phone p was repeated into
//   embedding_bwd      demb[idx[m]] += dx[m]   (runs of equal indices are pre-summed in registers)
//   rowdot_mask_bwd    predictor head Linear(f,1) + masked_fill
//   sum_over_time      dspk[b,:] += sum_t dx[b,t,:]
//   fold_pw            W_eff = W21 . blockdiag(W20) and its backward (grouped 1x1 conv folded into conv2.1)
#include <math.h>

#include "common.cuh"

namespace lfs2 {

// ---------------------------------------------------------------------------------------
// c[n0+i, k0+j] += sum_{r in split} a[r, n0+i] * b[r + shift, k0+j]
// 128x128 output tile per CTA, 8x8 register micro-tiles; the contraction index is the ROW of
// both operands, so both tiles are loaded row-major with coalesced float4 loads (no transpose).
constexpr int TN_B = 128, TN_R = 16, TN_PAD = 4;
constexpr int kTnThreads = 256;

__global__ void __launch_bounds__(kTnThreads)
gemm_tn_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, int m, int n, int k,
               int lda, int ldb, int ldc, int rows_per_split, int t, int shift) {
  __shared__ __align__(16) float As[2][TN_R][TN_B + TN_PAD];
  __shared__ __align__(16) float Bs[2][TN_R][TN_B + TN_PAD];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * TN_B, n0 = blockIdx.y * TN_B;
  const int r_begin = blockIdx.z * rows_per_split;
  const int r_end = min(m, r_begin + rows_per_split);
  if (r_begin >= r_end) return;
  const int lr = tid >> 5, lc = (tid & 31) * 4;  // tile row 0..7 (+8), column group
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int r0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = r0 + lr + h * 8;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (row < r_end) {
        if (n0 + lc < n) va = *reinterpret_cast<const float4*>(a + (size_t)row * lda + n0 + lc);
        bool ok = k0 + lc < k;
        if (t > 0) {
          const int tt = row % t + shift;
          ok = ok && tt >= 0 && tt < t;
        }
        if (ok) vb = *reinterpret_cast<const float4*>(b + (size_t)(row + shift) * ldb + k0 + lc);
      }
      ra[h] = va;
      rb[h] = vb;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&As[buf][lr + h * 8][lc]) = ra[h];
      *reinterpret_cast<float4*>(&Bs[buf][lr + h * 8][lc]) = rb[h];
    }
  };

  load_tiles(r_begin);
  store_tiles(0);
  __syncthreads();
  const int nsteps = (r_end - r_begin + TN_R - 1) / TN_R;
  for (int st = 0; st < nsteps; ++st) {
    const int buf = st & 1;
    if (st + 1 < nsteps) load_tiles(r_begin + (st + 1) * TN_R);
#pragma unroll
    for (int kk = 0; kk < TN_R; ++kk) {
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
    if (st + 1 < nsteps) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col < k) atomicAdd(c + (size_t)row * ldc + col, acc[i][j]);
    }
  }
}

// out[col] += sum_m a[m, col]; CTA = 32 float4 column groups x 8 row lanes over a row range
__global__ void colsum_kernel(const float4* __restrict__ a, float* __restrict__ out, int m, int n4, int rows_per_cta) {
  __shared__ float4 part[8][32];
  const int cg = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < n4)
    for (int r = r0 + rl; r < r1; r += 8) {
      float4 v = a[(size_t)r * n4 + cg];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  part[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && cg < n4) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      float4 v = part[i][threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    atomicAdd(out + cg * 4 + 0, s.x);
    atomicAdd(out + cg * 4 + 1, s.y);
    atomicAdd(out + cg * 4 + 2, s.z);
    atomicAdd(out + cg * 4 + 3, s.w);
  }
}

// dx = y > 0 ? dy * scale : 0.  scale != 1 folds the backward of a dropout that followed the ReLU: the saved y is the
// dropped activation (0 where dropped), so "y > 0" already is relu-mask AND keep-mask; only 1/(1-p) remains
__global__ void relu_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ dx,
                                size_t n4, float scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = dy[i], v = y[i];
  g.x = v.x > 0.f ? g.x * scale : 0.f;
  g.y = v.y > 0.f ? g.y * scale : 0.f;
  g.z = v.z > 0.f ? g.z * scale : 0.f;
  g.w = v.w > 0.f ? g.w * scale : 0.f;
  dx[i] = g;
}

__global__ void add_inplace_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = dst[i], b = src[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  dst[i] = a;
}

// (rows, cols) -> (cols, rows), 32x32 tiles through shared memory
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)(r0 + i) * cols + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < cols && r < rows) out[(size_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

// ---------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (grid-stride), d <= 1024:
//   xhat = (z - mean) * rstd ; g = dy * gamma ; dz = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ add)
//   dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy     (register partials, one atomic per lane at the end)
constexpr int kLnbMaxVec = 8;
// RELOAD (wide rows, NV >= 4): the second pass re-reads dy / z (L1 / L2 hits: the warp has just streamed the row) instead
// of keeping g and xhat in 8 * NV registers; with the dgamma / dbeta partials that is what held the d = 768 kernel at
// 139 registers = ONE 256-thread CTA per SM (2.3 TB/s).  Same arithmetic in the same order either way.
template <int NV, bool RELOAD>  // float4 column groups per lane: d <= 128 * NV
__global__ void __launch_bounds__(256, RELOAD ? (NV > 6 ? 2 : 3) : 1)
layernorm_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ dy2, const float4* __restrict__ z,
                     const float2* __restrict__ stats,
                     const float4* __restrict__ gamma, const float4* __restrict__ add, float4* __restrict__ dz,
                     float4* __restrict__ dz_drop, float* __restrict__ dgamma, float* __restrict__ dbeta, int m, int d4,
                     DropSite drop) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 gacc[NV], bacc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    bacc[i] = gacc[i];
  }
  const float inv_d = 1.f / (float)(d4 * 4);
  for (int row = warp; row < m; row += nwarps) {
    const float2 st = stats[row];
    const float mean = st.x, rstd = st.y;
    float4 g[RELOAD ? 1 : NV], xh[RELOAD ? 1 : NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < d4) {
        float4 dv = dy[(size_t)row * d4 + c];
        if (dy2) {  // the incoming gradient as two summands (a residual join upstream): added on load
          const float4 e = dy2[(size_t)row * d4 + c];
          dv.x += e.x; dv.y += e.y; dv.z += e.z; dv.w += e.w;
        }
        const float4 zv = z[(size_t)row * d4 + c], gm = __ldg(gamma + c);
        float4 x;
        x.x = (zv.x - mean) * rstd; x.y = (zv.y - mean) * rstd; x.z = (zv.z - mean) * rstd; x.w = (zv.w - mean) * rstd;
        gacc[i].x = fmaf(dv.x, x.x, gacc[i].x); gacc[i].y = fmaf(dv.y, x.y, gacc[i].y);
        gacc[i].z = fmaf(dv.z, x.z, gacc[i].z); gacc[i].w = fmaf(dv.w, x.w, gacc[i].w);
        bacc[i].x += dv.x; bacc[i].y += dv.y; bacc[i].z += dv.z; bacc[i].w += dv.w;
        float4 gg;
        gg.x = dv.x * gm.x; gg.y = dv.y * gm.y; gg.z = dv.z * gm.z; gg.w = dv.w * gm.w;
        s1 += (gg.x + gg.y) + (gg.z + gg.w);
        s2 += (gg.x * x.x + gg.y * x.y) + (gg.z * x.z + gg.w * x.w);
        if (!RELOAD) {
          g[i] = gg;
          xh[i] = x;
        }
      }
    }
    const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < d4) {
        float4 gi, xi;
        if (RELOAD) {
          float4 dv = dy[(size_t)row * d4 + c];
          if (dy2) {
            const float4 e = dy2[(size_t)row * d4 + c];
            dv.x += e.x; dv.y += e.y; dv.z += e.z; dv.w += e.w;
          }
          const float4 zv = z[(size_t)row * d4 + c], gm = __ldg(gamma + c);
          xi.x = (zv.x - mean) * rstd; xi.y = (zv.y - mean) * rstd; xi.z = (zv.z - mean) * rstd; xi.w = (zv.w - mean) * rstd;
          gi.x = dv.x * gm.x; gi.y = dv.y * gm.y; gi.z = dv.z * gm.z; gi.w = dv.w * gm.w;
        } else {
          gi = g[i];
          xi = xh[i];
        }
        float4 o;
        o.x = rstd * (gi.x - c1 - xi.x * c2);
        o.y = rstd * (gi.y - c1 - xi.y * c2);
        o.z = rstd * (gi.z - c1 - xi.z * c2);
        o.w = rstd * (gi.w - c1 - xi.w * c2);
        if (dz_drop) {  // gradient of the dropped branch: z = x + drop(y)  =>  dy = drop(dz) with the forward's mask
          const float4 k = dropout_scale4((size_t)row * d4 + c, drop.threshold, drop.inv_keep, drop.key, drop.site);
          dz_drop[(size_t)row * d4 + c] = make_float4(o.x * k.x, o.y * k.y, o.z * k.z, o.w * k.w);
        }
        if (add) {
          const float4 e = add[(size_t)row * d4 + c];
          o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
        }
        dz[(size_t)row * d4 + c] = o;
      }
    }
  }
  // CTA-level pre-reduction of the 8 warps' partials through shared memory, then one atomic per column
  __shared__ float4 red[8][32];
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (32 * i < d4) {  // CTA-uniform
        const int c = lane + 32 * i;
        red[w][lane] = pass == 0 ? gacc[i] : bacc[i];
        __syncthreads();
        if (w == 0 && c < d4) {
          float4 s = red[0][lane];
          for (int q = 1; q < 8; ++q) {
            float4 v = red[q][lane];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
          }
          float* dst = (pass == 0 ? dgamma : dbeta) + c * 4;
          atomicAdd(dst + 0, s.x);
          atomicAdd(dst + 1, s.y);
          atomicAdd(dst + 2, s.z);
          atomicAdd(dst + 3, s.w);
        }
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// depthwise conv weight/bias gradient.  thread = one channel, walks kDwbRows consecutive frames of one
// utterance with the K-frame input window in registers:
//   dwt[j,c] += sum_t dy[b,t,c] * x[b,t+j-h,c] ; dbias[c] += sum_t dy[b,t,c]
constexpr int kDwbRows = 64;
template <int K>
__global__ void __launch_bounds__(128)
dwconv1d_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dwt,
                      float* __restrict__ dbias, int t, int d) {
  constexpr int H = (K - 1) / 2;
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= d) return;
  const int b = blockIdx.z, t0 = blockIdx.x * kDwbRows;
  const float* xb = x + (size_t)b * t * d + c;
  const float* gb = dy + (size_t)b * t * d + c;
  float win[K], acc[K];
  float accb = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    acc[j] = 0.f;
    const int ti = t0 - H + j;
    win[j] = (j < K - 1 && ti >= 0 && ti < t) ? xb[(size_t)ti * d] : 0.f;  // window for frame t0 minus its last entry
  }
  const int t1 = min(t, t0 + kDwbRows);
  for (int tt = t0; tt < t1; ++tt) {
    const int ti = tt + H;
    win[K - 1] = ti < t ? xb[(size_t)ti * d] : 0.f;
    const float g = gb[(size_t)tt * d];
    accb += g;
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fmaf(g, win[j], acc[j]);
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
  }
#pragma unroll
  for (int j = 0; j < K; ++j) atomicAdd(dwt + (size_t)j * d + c, acc[j]);
  if (dbias) atomicAdd(dbias + c, accb);
}

// ---------------------------------------------------------------------------------------
// LengthRegulator backward: warp per (utterance, phone); frames [cum[p-1], min(cum[p], l)) are summed
// in ascending order (deterministic).
__global__ void lr_bwd_kernel(const float4* __restrict__ dout, const int64_t* __restrict__ cum, float4* __restrict__ dx,
                              int batch, int tp, int l, int d4) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= batch * tp) return;
  const int b = warp / tp, p = warp % tp;
  const long long start = p ? cum[(size_t)b * tp + p - 1] : 0;
  long long end = cum[(size_t)b * tp + p];
  if (end > l) end = l;
  for (int c = lane; c < d4; c += 32) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long f = start; f < end; ++f) {
      float4 v = dout[((size_t)b * l + f) * d4 + c];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dx[(size_t)warp * d4 + c] = s;
  }
}

// ---------------------------------------------------------------------------------------
// demb[idx[m], :] += dx[m, :].  A warp walks kEmbRows consecutive rows and pre-sums runs of equal
// indices in registers (PAD frames of an utterance all hit one bucket / the padding row), so the hot
// row sees one atomic per run instead of one per frame.  Rows with idx == skip_idx get no gradient
// (nn.Embedding padding_idx).
constexpr int kEmbRows = 32;
__global__ void embedding_bwd_kernel(const float4* __restrict__ dx, const int64_t* __restrict__ idx,
                                     float* __restrict__ demb, int m, int d4, int nrows_emb, long long skip_idx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * kEmbRows;
  if (r0 >= m) return;
  const int r1 = min(m, r0 + kEmbRows);
  for (int c0 = 0; c0 < d4; c0 += 32) {
    const int c = c0 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long cur = -1;
    for (int r = r0; r <= r1; ++r) {
      long long id = r < r1 ? idx[r] : -2;
      if (id != cur) {
        if (cur >= 0 && cur < nrows_emb && cur != skip_idx && c < d4) {
          float* dst = demb + ((size_t)cur * d4 + c) * 4;
          atomicAdd(dst + 0, acc.x);
          atomicAdd(dst + 1, acc.y);
          atomicAdd(dst + 2, acc.z);
          atomicAdd(dst + 3, acc.w);
        }
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cur = id;
      }
      if (r < r1 && c < d4) {
        float4 v = dx[(size_t)r * d4 + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// predictor head backward: out[m] = mask ? 0 : z[m,:].w + b
//   dz[m,:] = g * w ; dw += sum_m g * z[m,:] ; db += sum_m g     with g = mask[m] ? 0 : dout[m]
__global__ void __launch_bounds__(256)
rowdot_mask_bwd_kernel(const float* __restrict__ dout, const float4* __restrict__ z, const float4* __restrict__ w,
                       const uint8_t* __restrict__ mask, float4* __restrict__ dz, float* __restrict__ dw,
                       float* __restrict__ db, int m, int f4) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 wacc[kLnbMaxVec], wv[kLnbMaxVec];
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    wacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = lane + 32 * i;
    wv[i] = c < f4 ? w[c] : wacc[i];
  }
  float bacc = 0.f;
  for (int row = warp; row < m; row += nwarps) {
    const float g = (mask && mask[row]) ? 0.f : dout[row];
    bacc += g;
#pragma unroll
    for (int i = 0; i < kLnbMaxVec; ++i) {
      int c = lane + 32 * i;
      if (c < f4) {
        float4 zv = z[(size_t)row * f4 + c];
        wacc[i].x = fmaf(g, zv.x, wacc[i].x); wacc[i].y = fmaf(g, zv.y, wacc[i].y);
        wacc[i].z = fmaf(g, zv.z, wacc[i].z); wacc[i].w = fmaf(g, zv.w, wacc[i].w);
        dz[(size_t)row * f4 + c] = make_float4(g * wv[i].x, g * wv[i].y, g * wv[i].z, g * wv[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    int c = lane + 32 * i;
    if (c < f4) {
      atomicAdd(dw + c * 4 + 0, wacc[i].x);
      atomicAdd(dw + c * 4 + 1, wacc[i].y);
      atomicAdd(dw + c * 4 + 2, wacc[i].z);
      atomicAdd(dw + c * 4 + 3, wacc[i].w);
    }
  }
  if (lane == 0) atomicAdd(db, bacc);
}

// out[b, c] += sum_t dx[b, t, c]; CTA = (utterance, 32 float4 column groups, frame chunk)
__global__ void sum_over_time_kernel(const float4* __restrict__ dx, float* __restrict__ out, int t, int d4,
                                     int rows_per_cta) {
  __shared__ float4 part[8][32];
  const int cg = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5, b = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(t, r0 + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < d4)
    for (int r = r0 + rl; r < r1; r += 8) {
      float4 v = dx[((size_t)b * t + r) * d4 + cg];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  part[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && cg < d4) {
    for (int i = 1; i < 8; ++i) {
      float4 v = part[i][threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* dst = out + ((size_t)b * d4 + cg) * 4;
    atomicAdd(dst + 0, s.x);
    atomicAdd(dst + 1, s.y);
    atomicAdd(dst + 2, s.z);
    atomicAdd(dst + 3, s.w);
  }
}

// ---------------------------------------------------------------------------------------
// conv2.0 (F->F, groups=d, 1x1, weight (F, g, 1)) followed by conv2.1 (F->d, 1x1, weight (d, F, 1)) is one
// linear map.  With channel c = G*g + o (group G, member o):
//   W_eff[n, G*g+i] = sum_o W21[n, G*g+o] * W20[G*g+o, i] ;  b_eff[n] = b21[n] + sum_c W21[n,c] * b20[c]
// thread = one (n, G): g x g block (g <= 8)
constexpr int kFoldMaxG = 8;
__global__ void fold_pw_fwd_kernel(const float* __restrict__ w21, const float* __restrict__ w20,
                                   const float* __restrict__ b20, const float* __restrict__ b21,
                                   float* __restrict__ w_eff, float* __restrict__ b_eff, int d_out, int groups, int g) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= d_out * groups) return;
  const int n = i0 / groups, G = i0 % groups, f = groups * g;
  float o[kFoldMaxG];
  for (int i = 0; i < g; ++i) o[i] = 0.f;
  for (int oo = 0; oo < g; ++oo) {
    const float w = w21[(size_t)n * f + G * g + oo];
    for (int i = 0; i < g; ++i) o[i] = fmaf(w, w20[(size_t)(G * g + oo) * g + i], o[i]);
  }
  for (int i = 0; i < g; ++i) w_eff[(size_t)n * f + G * g + i] = o[i];
}

// b_eff[n] = W21[n, :] . b20 + b21[n]; one warp per output channel, fixed summation order (two model instances built
// from the same weights must produce the same bits: checkpoints, multi-GPU replicas)
__global__ void fold_pw_bias_kernel(const float* __restrict__ w21, const float* __restrict__ b20,
                                    const float* __restrict__ b21, float* __restrict__ b_eff, int d_out, int f) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= d_out) return;
  float acc = 0.f;
  for (int c = lane; c < f; c += 32) acc = fmaf(w21[(size_t)n * f + c], b20[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) b_eff[n] = acc + b21[n];
}

// dW21[n, G*g+o] += sum_i dWe[n, G*g+i] * W20[G*g+o, i] + dbe[n] * b20[G*g+o]
// dW20[G*g+o, i] += sum_n dWe[n, G*g+i] * W21[n, G*g+o] ; db20[c] += sum_n dbe[n] * W21[n,c] ; db21 += dbe
// thread = one group G of conv2.0, CTA = 128 groups x a chunk of kFoldRows output rows n: the group's g x g weight
// gradient and its g bias gradients are summed over the chunk in registers and reach memory as ONE atomic per entry and
// CTA (one thread per (n, G) used to add straight into dw20 / db20: d_out-way contention on every address, 141 us at
// d = 768, F = 3072); dw21[n, c] is owned by exactly one thread of the launch (plain accumulate).
constexpr int kFoldRows = 32;
__global__ void __launch_bounds__(128)
fold_pw_bwd_kernel(const float* __restrict__ dwe, const float* __restrict__ dbe, const float* __restrict__ w21,
                   const float* __restrict__ w20, const float* __restrict__ b20, float* __restrict__ dw21,
                   float* __restrict__ dw20, float* __restrict__ db20, float* __restrict__ db21, int d_out, int groups,
                   int g) {
  const int G = blockIdx.x * blockDim.x + threadIdx.x;
  if (G >= groups) return;
  const int f = groups * g;
  const int n0 = blockIdx.y * kFoldRows, n1 = min(d_out, n0 + kFoldRows);
  float wg[kFoldMaxG][kFoldMaxG], bg[kFoldMaxG], acc[kFoldMaxG][kFoldMaxG], accb[kFoldMaxG];
#pragma unroll
  for (int oo = 0; oo < kFoldMaxG; ++oo) {
    bg[oo] = oo < g ? b20[G * g + oo] : 0.f;
    accb[oo] = 0.f;
#pragma unroll
    for (int i = 0; i < kFoldMaxG; ++i) {
      wg[oo][i] = (oo < g && i < g) ? w20[(size_t)(G * g + oo) * g + i] : 0.f;
      acc[oo][i] = 0.f;
    }
  }
  for (int n = n0; n < n1; ++n) {
    const float dbn = dbe[n];
    float de[kFoldMaxG], w[kFoldMaxG];
#pragma unroll
    for (int i = 0; i < kFoldMaxG; ++i) {
      de[i] = i < g ? dwe[(size_t)n * f + G * g + i] : 0.f;
      w[i] = i < g ? w21[(size_t)n * f + G * g + i] : 0.f;
    }
#pragma unroll
    for (int oo = 0; oo < kFoldMaxG; ++oo) {
      if (oo < g) {
        float sum = dbn * bg[oo];
#pragma unroll
        for (int i = 0; i < kFoldMaxG; ++i) {
          sum = fmaf(de[i], wg[oo][i], sum);
          acc[oo][i] = fmaf(de[i], w[oo], acc[oo][i]);
        }
        dw21[(size_t)n * f + G * g + oo] += sum;
        accb[oo] = fmaf(dbn, w[oo], accb[oo]);
      }
    }
    if (G == 0) atomicAdd(db21 + n, dbn);
  }
#pragma unroll
  for (int oo = 0; oo < kFoldMaxG; ++oo) {
    if (oo < g) {
#pragma unroll
      for (int i = 0; i < kFoldMaxG; ++i)
        if (i < g) atomicAdd(dw20 + (size_t)(G * g + oo) * g + i, acc[oo][i]);
      atomicAdd(db20 + G * g + oo, accb[oo]);
    }
  }
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_gemm_tn(const float* a, const float* b, float* c, int m, int n, int k, int lda, int ldb, int ldc, int t,
                 int shift, void* stream) {
  LFS2_REQUIRE(a && b && c, LFS2_ERR_INVALID_ARG, "gemm_tn: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && n > 0 && k > 0 && t >= 0, LFS2_ERR_INVALID_ARG, "gemm_tn: bad shape");
  LFS2_REQUIRE(n % 4 == 0 && k % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, LFS2_ERR_UNSUPPORTED,
               "gemm_tn: n, k and the leading dimensions must be multiples of 4 (n=%d k=%d)", n, k);
  LFS2_REQUIRE(t > 0 || shift == 0, LFS2_ERR_INVALID_ARG, "gemm_tn: a row shift needs the utterance length t");
  LFS2_REQUIRE(aligned16(a) && aligned16(b), LFS2_ERR_INVALID_ARG, "gemm_tn: operands must be 16-byte aligned");
  const int tiles = ceil_div(k, TN_B) * ceil_div(n, TN_B);
  int splits = ceil_div(4 * num_sms(), tiles);
  int rows_per = ceil_div(m, splits);
  rows_per = ceil_div(rows_per, TN_R) * TN_R;
  if (rows_per < 4 * TN_R) rows_per = 4 * TN_R;
  splits = ceil_div(m, rows_per);
  dim3 grid(ceil_div(k, TN_B), ceil_div(n, TN_B), splits);
  LFS2_REQUIRE(grid.y <= 65535 && grid.z <= 65535, LFS2_ERR_UNSUPPORTED, "gemm_tn: grid too large");
  gemm_tn_kernel<<<grid, kTnThreads, 0, (cudaStream_t)stream>>>(a, b, c, m, n, k, lda, ldb, ldc, rows_per, t, shift);
  LFS2_CHECK_LAUNCH("gemm_tn");
  return LFS2_OK;
}

int lfs2_colsum(const float* a, float* out, int m, int n, void* stream) {
  LFS2_REQUIRE(a && out, LFS2_ERR_INVALID_ARG, "colsum: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "colsum: n=%d must be a positive multiple of 4", n);
  LFS2_REQUIRE(aligned16(a), LFS2_ERR_INVALID_ARG, "colsum: input must be 16-byte aligned");
  const int cgs = ceil_div(n / 4, 32);
  int chunks = ceil_div(4 * num_sms(), cgs);
  int rows_per = ceil_div(m, chunks);
  if (rows_per < 64) rows_per = 64;
  dim3 grid(cgs, ceil_div(m, rows_per));
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)a, out, m, n / 4, rows_per);
  LFS2_CHECK_LAUNCH("colsum");
  return LFS2_OK;
}

int lfs2_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream) {
  return lfs2_relu_bwd_scaled(dy, y, dx, n, 1.f, stream);
}

int lfs2_relu_bwd_scaled(const float* dy, const float* y, float* dx, long long n, float scale, void* stream) {
  LFS2_REQUIRE(dy && y && dx, LFS2_ERR_INVALID_ARG, "relu_bwd: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "relu_bwd: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(dy) && aligned16(y) && aligned16(dx), LFS2_ERR_INVALID_ARG, "relu_bwd: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  relu_bwd_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)dy, (const float4*)y, (float4*)dx, n4,
                                                                       scale);
  LFS2_CHECK_LAUNCH("relu_bwd");
  return LFS2_OK;
}

int lfs2_add_inplace(float* dst, const float* src, long long n, void* stream) {
  LFS2_REQUIRE(dst && src, LFS2_ERR_INVALID_ARG, "add_inplace: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "add_inplace: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(dst) && aligned16(src), LFS2_ERR_INVALID_ARG, "add_inplace: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  add_inplace_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((float4*)dst, (const float4*)src, n4);
  LFS2_CHECK_LAUNCH("add_inplace");
  return LFS2_OK;
}

int lfs2_transpose(const float* in, float* out, int rows, int cols, void* stream) {
  LFS2_REQUIRE(in && out, LFS2_ERR_INVALID_ARG, "transpose: null pointer");
  if (rows == 0 || cols == 0) return LFS2_OK;
  LFS2_REQUIRE(rows > 0 && cols > 0, LFS2_ERR_INVALID_ARG, "transpose: bad shape");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  LFS2_REQUIRE(grid.y <= 65535, LFS2_ERR_UNSUPPORTED, "transpose: too many rows");
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, rows, cols);
  LFS2_CHECK_LAUNCH("transpose");
  return LFS2_OK;
}

int lfs2_layernorm_bwd(const float* dy, const float* z, const float* stats, const float* gamma, const float* add,
                       float* dz, float* dgamma, float* dbeta, int m, int d, void* stream) {
  return lfs2_layernorm_bwd_drop(dy, z, stats, gamma, add, dz, nullptr, dgamma, dbeta, m, d, 0.f, 0ull, 0u, stream);
}

int lfs2_layernorm_bwd_drop(const float* dy, const float* z, const float* stats, const float* gamma, const float* add,
                            float* dz, float* dz_drop, float* dgamma, float* dbeta, int m, int d, float drop_p,
                            unsigned long long drop_seed, unsigned int drop_site, void* stream) {
  return lfs2_layernorm_bwd_ex(dy, nullptr, z, stats, gamma, add, dz, dz_drop, dgamma, dbeta, m, d, drop_p, drop_seed,
                               drop_site, stream);
}

int lfs2_layernorm_bwd_ex(const float* dy, const float* dy2, const float* z, const float* stats, const float* gamma,
                          const float* add, float* dz, float* dz_drop, float* dgamma, float* dbeta, int m, int d,
                          float drop_p, unsigned long long drop_seed, unsigned int drop_site, void* stream) {
  LFS2_REQUIRE(dy && z && stats && gamma && dz && dgamma && dbeta, LFS2_ERR_INVALID_ARG, "layernorm_bwd: null pointer");
  LFS2_REQUIRE(aligned16(dy2), LFS2_ERR_INVALID_ARG, "layernorm_bwd: dy2 must be 16-byte aligned");
  LFS2_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (!dz_drop || aligned16(dz_drop)), LFS2_ERR_INVALID_ARG,
               "layernorm_bwd: bad dropout arguments");
  const DropSite drop = make_drop_site(drop_p, drop_seed, drop_site);
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0 && d <= 128 * kLnbMaxVec, LFS2_ERR_UNSUPPORTED,
               "layernorm_bwd: d=%d must be a multiple of 4 and <= %d", d, 128 * kLnbMaxVec);
  LFS2_REQUIRE(aligned16(dy) && aligned16(z) && aligned16(gamma) && aligned16(dz) && (!add || aligned16(add)) &&
                   ((reinterpret_cast<uintptr_t>(stats) & 7u) == 0),
               LFS2_ERR_INVALID_ARG, "layernorm_bwd: pointers must be 16-byte aligned");
  // 8 rows per warp on large launches (fewer dgamma / dbeta atomics), down to 2 when that would leave CTA slots empty
  // (three 256-thread CTAs fit per SM: a data-parallel rank's 10 k rows used 160 of 444 slots)
  int rows_per_warp = ceil_div(m, 3 * num_sms() * 8);
  rows_per_warp = rows_per_warp < 2 ? 2 : (rows_per_warp > 8 ? 8 : rows_per_warp);
  int blocks = ceil_div(m, 8 * rows_per_warp);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  if (blocks < 1) blocks = 1;
  const int nv = ceil_div(d / 4, 32);
#define LFS2_LNB(NV)                                                                                            \
  layernorm_bwd_kernel<NV, (NV >= 4)><<<blocks, 256, 0, (cudaStream_t)stream>>>(                                \
      (const float4*)dy, (const float4*)dy2, (const float4*)z, (const float2*)stats, (const float4*)gamma,      \
      (const float4*)add,                                                                                       \
      (float4*)dz, (float4*)dz_drop, dgamma, dbeta, m, d / 4, drop)
  if (nv <= 1) LFS2_LNB(1);
  else if (nv <= 2) LFS2_LNB(2);
  else if (nv <= 4) LFS2_LNB(4);
  else if (nv <= 6) LFS2_LNB(6);
  else LFS2_LNB(8);
#undef LFS2_LNB
  LFS2_CHECK_LAUNCH("layernorm_bwd");
  return LFS2_OK;
}

int lfs2_dwconv1d_bwd_w(const float* dy, const float* x, float* dwt, float* dbias, int batch, int t, int d, int ksize,
                        void* stream) {
  LFS2_REQUIRE(dy && x && dwt, LFS2_ERR_INVALID_ARG, "dwconv1d_bwd_w: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0, LFS2_ERR_INVALID_ARG, "dwconv1d_bwd_w: bad shape");
  LFS2_REQUIRE(batch <= 65535, LFS2_ERR_UNSUPPORTED, "dwconv1d_bwd_w: batch exceeds the grid limit");
  dim3 grid(ceil_div(t, kDwbRows), ceil_div(d, 128), batch);
  cudaStream_t s = (cudaStream_t)stream;
#define LFS2_DWB_CASE(K) \
  case K: dwconv1d_bwd_w_kernel<K><<<grid, 128, 0, s>>>(dy, x, dwt, dbias, t, d); break;
  switch (ksize) {
    LFS2_DWB_CASE(1) LFS2_DWB_CASE(3) LFS2_DWB_CASE(5) LFS2_DWB_CASE(7) LFS2_DWB_CASE(9) LFS2_DWB_CASE(11)
    LFS2_DWB_CASE(13) LFS2_DWB_CASE(15) LFS2_DWB_CASE(17) LFS2_DWB_CASE(19) LFS2_DWB_CASE(21) LFS2_DWB_CASE(23)
    LFS2_DWB_CASE(25)
    default:
      set_error("dwconv1d_bwd_w: kernel size %d not supported (odd, 1..25)", ksize);
      return LFS2_ERR_UNSUPPORTED;
  }
#undef LFS2_DWB_CASE
  LFS2_CHECK_LAUNCH("dwconv1d_bwd_w");
  return LFS2_OK;
}

int lfs2_length_regulate_bwd(const float* dout, const int64_t* cum, float* dx, int batch, int tp, int l, int d,
                             void* stream) {
  LFS2_REQUIRE(cum && dx && (dout || l == 0), LFS2_ERR_INVALID_ARG, "length_regulate_bwd: null pointer");
  LFS2_REQUIRE(batch > 0 && tp > 0 && l >= 0, LFS2_ERR_INVALID_ARG, "length_regulate_bwd: bad shape");
  LFS2_REQUIRE(d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "length_regulate_bwd: d=%d must be a multiple of 4", d);
  LFS2_REQUIRE(aligned16(dx) && (!dout || aligned16(dout)), LFS2_ERR_INVALID_ARG,
               "length_regulate_bwd: pointers must be 16-byte aligned");
  lr_bwd_kernel<<<ceil_div((long long)batch * tp * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)dout, cum, (float4*)dx, batch, tp, l, d / 4);
  LFS2_CHECK_LAUNCH("length_regulate_bwd");
  return LFS2_OK;
}

int lfs2_embedding_bwd(const float* dx, const int64_t* idx, float* demb, int m, int d, int nrows_emb,
                       long long skip_idx, void* stream) {
  LFS2_REQUIRE(dx && idx && demb, LFS2_ERR_INVALID_ARG, "embedding_bwd: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && d > 0 && d % 4 == 0 && nrows_emb > 0, LFS2_ERR_UNSUPPORTED, "embedding_bwd: bad shape");
  LFS2_REQUIRE(aligned16(dx), LFS2_ERR_INVALID_ARG, "embedding_bwd: dx must be 16-byte aligned");
  const int warps = ceil_div(m, kEmbRows);
  embedding_bwd_kernel<<<ceil_div((long long)warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)dx, idx, demb, m, d / 4, nrows_emb, skip_idx);
  LFS2_CHECK_LAUNCH("embedding_bwd");
  return LFS2_OK;
}

int lfs2_rowdot_mask_bwd(const float* dout, const float* z, const float* w, const uint8_t* mask, float* dz, float* dw,
                         float* db, int m, int f, void* stream) {
  LFS2_REQUIRE(dout && z && w && dz && dw && db, LFS2_ERR_INVALID_ARG, "rowdot_mask_bwd: null pointer");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(f > 0 && f % 4 == 0 && f <= 128 * kLnbMaxVec, LFS2_ERR_UNSUPPORTED,
               "rowdot_mask_bwd: f=%d must be a multiple of 4 and <= %d", f, 128 * kLnbMaxVec);
  LFS2_REQUIRE(aligned16(z) && aligned16(w) && aligned16(dz), LFS2_ERR_INVALID_ARG,
               "rowdot_mask_bwd: pointers must be 16-byte aligned");
  int blocks = ceil_div(m, 8 * 16);
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  if (blocks < 1) blocks = 1;
  rowdot_mask_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dout, (const float4*)z, (const float4*)w, mask,
                                                                 (float4*)dz, dw, db, m, f / 4);
  LFS2_CHECK_LAUNCH("rowdot_mask_bwd");
  return LFS2_OK;
}

int lfs2_sum_over_time(const float* dx, float* out, int batch, int t, int d, void* stream) {
  LFS2_REQUIRE(dx && out, LFS2_ERR_INVALID_ARG, "sum_over_time: null pointer");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(d > 0 && d % 4 == 0, LFS2_ERR_UNSUPPORTED, "sum_over_time: d=%d must be a multiple of 4", d);
  LFS2_REQUIRE(aligned16(dx), LFS2_ERR_INVALID_ARG, "sum_over_time: dx must be 16-byte aligned");
  LFS2_REQUIRE(batch <= 65535, LFS2_ERR_UNSUPPORTED, "sum_over_time: batch exceeds the grid limit");
  const int cgs = ceil_div(d / 4, 32);
  int rows_per = 128;
  dim3 grid(cgs, ceil_div(t, rows_per), batch);
  sum_over_time_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)dx, out, t, d / 4, rows_per);
  LFS2_CHECK_LAUNCH("sum_over_time");
  return LFS2_OK;
}

int lfs2_fold_pw_fwd(const float* w21, const float* w20, const float* b20, const float* b21, float* w_eff,
                     float* b_eff, int d_out, int groups, int g, void* stream) {
  LFS2_REQUIRE(w21 && w20 && b20 && b21 && w_eff && b_eff, LFS2_ERR_INVALID_ARG, "fold_pw_fwd: null pointer");
  LFS2_REQUIRE(d_out > 0 && groups > 0 && g > 0 && g <= kFoldMaxG, LFS2_ERR_UNSUPPORTED,
               "fold_pw: group size %d not supported (1..%d)", g, kFoldMaxG);
  cudaStream_t s = (cudaStream_t)stream;
  fold_pw_fwd_kernel<<<ceil_div((long long)d_out * groups, 256), 256, 0, s>>>(w21, w20, b20, b21, w_eff, b_eff, d_out,
                                                                             groups, g);
  LFS2_CHECK_LAUNCH("fold_pw_fwd");
  fold_pw_bias_kernel<<<ceil_div((long long)d_out * 32, 256), 256, 0, s>>>(w21, b20, b21, b_eff, d_out, groups * g);
  LFS2_CHECK_LAUNCH("fold_pw_bias");
  return LFS2_OK;
}

int lfs2_fold_pw_bwd(const float* dw_eff, const float* db_eff, const float* w21, const float* w20, const float* b20,
                     float* dw21, float* dw20, float* db20, float* db21, int d_out, int groups, int g, void* stream) {
  LFS2_REQUIRE(dw_eff && db_eff && w21 && w20 && b20 && dw21 && dw20 && db20 && db21, LFS2_ERR_INVALID_ARG,
               "fold_pw_bwd: null pointer");
  LFS2_REQUIRE(d_out > 0 && groups > 0 && g > 0 && g <= kFoldMaxG, LFS2_ERR_UNSUPPORTED,
               "fold_pw: group size %d not supported (1..%d)", g, kFoldMaxG);
  dim3 grid(ceil_div(groups, 128), ceil_div(d_out, kFoldRows));
  fold_pw_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dw_eff, db_eff, w21, w20, b20, dw21, dw20, db20, db21, d_out,
                                                             groups, g);
  LFS2_CHECK_LAUNCH("fold_pw_bwd");
  return LFS2_OK;
}

}  // extern "C"
