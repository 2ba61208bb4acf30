// Depthwise conv of the LightSpeech FFN blocks for the 11 .. 19-tap kernels on plane-form activations, with the input
// tiles staged by TMA (north_star: "TMA staging of (B, T, d_model) tiles into shared memory").
//
// dwconv1d_k_kernel (elementwise.cu) loads a tile with per-thread global loads, converts it to fp32 in shared memory
// and only then starts its arithmetic: with 32-row tiles and 20 halo rows the load phase (~2 us of latency) is longer
// than the arithmetic, and two CTAs per SM overlap it only partly (k = 21: 98 us for the 43 M outputs of a decoder
// layer, 2.6 TB/s, against 40 us of HBM time).  Here a CTA walks kTilesPerCta consecutive 32-row tiles of one utterance
// through a two-stage ring: one thread issues the NEXT tile's two bulk-tensor copies (hi and lo plane, (32 + k - 1) rows
// x 256 channels each, rows outside [0, T) zero-filled by TMA = Conv1d's "same" padding) before the CTA starts on the
// current one, so the copy engine, not the warps, waits for memory.  The threads read the bf16 planes straight from the
// ring (a warp's lanes cover 256 contiguous bytes per row and plane), rebuild fp32 as hi + lo and slide NT frames of
// their 4 channels through registers exactly like dwconv1d_k_kernel: same operations in the same order, so the two
// kernels agree bit for bit (tests/test_gpu_ops.py).  Measured at the C2 decoder's launch size (profiles/r4l_dwconv_ab.txt,
// fp16 plane out): k = 13 88 -> 73 us, k = 17 92 -> 87 us; at k >= 21 the ring version is 14 % SLOWER (112 vs 98 us: the
// kernel is issue-bound there -- 4 k tap registers under the 128-register cap of two CTAs per SM, conversions on every
// read -- and the copy latency it hides was already covered by the second CTA), so those stay on dwconv1d_k_kernel.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kDtCh4 = 64;          // float4 channel groups per CTA: 256 channels
constexpr int kDtTile = 32;         // output frames per tile
constexpr int kDtNT = 8;            // frames per thread
constexpr int kDtThreads = kDtCh4 * (kDtTile / kDtNT);  // 256
// ring depth S = 2 with two CTAs per SM.  (S = 4 with one CTA per SM -- three tiles in flight instead of two -- was
// 25-40 % slower at k >= 17: eight warps do not cover the arithmetic's latencies; profiles/r4l_dwconv_ab.txt.)
__host__ __device__ constexpr int dt_tiles_per_cta(int stages) { return stages == 2 ? 4 : 8; }
constexpr int kDtRowBytes = kDtCh4 * 4 * 2;             // one plane row of the CTA's channels: 512 B

__device__ __forceinline__ float4 dt_planes_to_f4(uint2 h, uint2 l) {
  float4 r;
  r.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  r.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  r.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  r.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return r;
}

template <int K, int S>
__global__ void __launch_bounds__(kDtThreads, S == 2 ? 2 : 1)
dwconv1d_tma_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const float4* __restrict__ wt, const float4* __restrict__ bias, float4* __restrict__ out,
                    uint2* __restrict__ out_hi, uint2* __restrict__ out_lo, uint2* __restrict__ out_f16, int t, int d4,
                    const int* __restrict__ row_limit, int limit_extra) {
  constexpr int H = (K - 1) / 2;
  constexpr int kRows = kDtTile + K - 1;
  constexpr int kPlane = kRows * kDtRowBytes;  // one plane of one stage
  constexpr int kStage = 2 * kPlane;           // hi | lo
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t full_bar[S];
  constexpr int kDtTilesPerCta = dt_tiles_per_cta(S);

  const int cb = blockIdx.y * kDtCh4;  // first float4 channel group of this CTA
  const int b = blockIdx.z;
  const int tile0 = blockIdx.x * kDtTilesPerCta;
  // rows the caller does not need (128-row groups at or after row_limit[b] + extra): neither computed nor, as inputs,
  // read from memory -- they count as zeros (same contract as dwconv1d_k_kernel)
  int t_in = t, t_end = t;
  if (row_limit) {
    const int lim = __ldg(row_limit + b) + limit_extra;
    t_in = min(t, (lim + 127) & ~127);
    t_end = t_in;  // tiles starting at or after the last kept 128-row group are skipped
  }
  int ntiles = 0;
  for (int i = 0; i < kDtTilesPerCta; ++i)
    if ((tile0 + i) * kDtTile < t_end) ntiles = i + 1;
  if (ntiles == 0) return;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_hi);
    prefetch_tmap(&map_lo);
    for (int q = 0; q < S; ++q) mbar_init(&full_bar[q], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int i) {  // tile tile0 + i -> stage i % S
    uint8_t* st = smem + (i % S) * kStage;
    uint64_t* bar = &full_bar[i % S];
    const int row0 = (tile0 + i) * kDtTile - H;
    mbar_expect_tx(bar, kStage);
    tma_load_3d(st, &map_hi, bar, cb * 4, row0, b);
    tma_load_3d(st + kPlane, &map_lo, bar, cb * 4, row0, b);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < S - 1 && i < ntiles; ++i) issue(i);

  const int c = threadIdx.x % kDtCh4, tg = threadIdx.x / kDtCh4;
  const size_t base = (size_t)b * t * d4 + cb;

#pragma unroll 1
  for (int i = 0; i < ntiles; ++i) {
    // the stage tile i + S - 1 lands in was read in iteration i - 1; every thread has passed that iteration's barrier
    if (threadIdx.x == 0 && i + S - 1 < ntiles) {
      fence_proxy_async_smem();
      issue(i + S - 1);
    }
    // (the taps are re-read per tile -- L1 hits -- instead of living in 4 K registers across the loop: at k >= 21 that
    //  is the difference between 118 registers and spills under the two-CTAs-per-SM cap)
    float4 w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) w[j] = __ldg(wt + (size_t)j * d4 + cb + c);
    const float4 bz = __ldg(bias + cb + c);
    mbar_wait(&full_bar[i % S], (i / S) & 1);
    const uint8_t* st = smem + (i % S) * kStage;
    const int t0 = (tile0 + i) * kDtTile;
    float4 acc[kDtNT];
#pragma unroll
    for (int o = 0; o < kDtNT; ++o) acc[o] = bz;
    const uint8_t* xr = st + (size_t)(tg * kDtNT) * kDtRowBytes + c * 8;
#pragma unroll
    for (int j = 0; j < kDtNT + K - 1; ++j) {
      const uint2 h = *reinterpret_cast<const uint2*>(xr + j * kDtRowBytes);
      const uint2 l = *reinterpret_cast<const uint2*>(xr + kPlane + j * kDtRowBytes);
      float4 xv = dt_planes_to_f4(h, l);
      if (t0 - H + tg * kDtNT + j >= t_in) xv = make_float4(0.f, 0.f, 0.f, 0.f);  // (only ever true with a row limit)
#pragma unroll
      for (int o = 0; o < kDtNT; ++o) {
        const int tap = j - o;  // compile-time after unrolling
        if (tap >= 0 && tap < K) {
          fma4(acc[o], w[tap], xv);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < kDtNT; ++o) {
      const int to = t0 + tg * kDtNT + o;
      if (to < t) {
        const size_t oi = base + (size_t)to * d4 + c;
        if (out) out[oi] = acc[o];
        if (out_hi) {
          uint2 hh, ll;
          split_pack2(acc[o].x, acc[o].y, hh.x, ll.x);
          split_pack2(acc[o].z, acc[o].w, hh.y, ll.y);
          out_hi[oi] = hh;
          out_lo[oi] = ll;
        }
        if (out_f16) {  // ONE fp16 plane (saturating): the activation operand of a 2-pass GEMM
          uint2 f;
          f.x = pack_f16_sat(acc[o].x, acc[o].y);
          f.y = pack_f16_sat(acc[o].z, acc[o].w);
          out_f16[oi] = f;
        }
      }
    }
    __syncthreads();  // all reads of stage i % S are done before tile i + S is copied into it
  }
}

template <int K, int S>
static int launch_dwconv_tma_k(const CUtensorMap& mh, const CUtensorMap& ml, const float* wt, const float* bias, float* out,
                               void* out_hi, void* out_lo, void* out_f16, int batch, int t, int d, const int* row_limit,
                               int limit_extra, cudaStream_t s) {
  constexpr int kSmem = S * 2 * (kDtTile + K - 1) * kDtRowBytes + 128;
  auto kern = dwconv1d_tma_kernel<K, S>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) {
      set_error("dwconv1d (TMA): cannot reserve %d bytes of shared memory", kSmem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles = ceil_div(t, kDtTile);
  dim3 grid(ceil_div(tiles, dt_tiles_per_cta(S)), d / 4 / kDtCh4, batch);
  kern<<<grid, kDtThreads, kSmem, s>>>(mh, ml, (const float4*)wt, (const float4*)bias, (float4*)out, (uint2*)out_hi,
                                       (uint2*)out_lo, (uint2*)out_f16, t, d / 4, row_limit, limit_extra);
  return LFS2_OK;
}

// -> LFS2_OK, an error, or 1 = "not my case" (the caller falls back to dwconv1d_k_kernel)
int launch_dwconv_tma(const void* x_hi, const void* x_lo, const float* wt, const float* bias, float* out, void* out_hi,
                      void* out_lo, void* out_f16, int batch, int t, int d, int ksize, const int* row_limit,
                      int limit_extra, cudaStream_t s) {
  static const int enabled = [] {
    const char* e = getenv("LFS2_DWCONV_TMA");  // A/B knob (tools): 0 keeps the per-thread-load kernel
    return e ? atoi(e) : 1;
  }();
  if (!enabled || ksize < 11 || ksize > (enabled == 3 ? 23 : 19) || d % (4 * kDtCh4) != 0 || t < 64) return 1;  // (3: A/B)
  CUtensorMap mh, ml;
  const uint32_t rows = kDtTile + ksize - 1;
  if (!make_tmap_3d_ex(&mh, x_hi, 2, d, t, batch, 4 * kDtCh4, rows, 0) ||
      !make_tmap_3d_ex(&ml, x_lo, 2, d, t, batch, 4 * kDtCh4, rows, 0))
    return 1;
#define LFS2_DT_CASE(K) \
  case K: return launch_dwconv_tma_k<K, 2>(mh, ml, wt, bias, out, out_hi, out_lo, out_f16, batch, t, d, row_limit, limit_extra, s);
  switch (ksize) {
    LFS2_DT_CASE(11) LFS2_DT_CASE(13) LFS2_DT_CASE(15) LFS2_DT_CASE(17) LFS2_DT_CASE(19) LFS2_DT_CASE(21) LFS2_DT_CASE(23)
  }
#undef LFS2_DT_CASE
  return 1;
}

}  // namespace tc
}  // namespace lfs2
