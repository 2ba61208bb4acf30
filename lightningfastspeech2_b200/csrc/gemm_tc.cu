// tcgen05 GEMM / implicit-GEMM Conv1d of the FFTBlock and predictor stacks:
//
//     acc[128 x N_TILE] (TMEM, fp32) = sum over (tap, k-slab)  A[rows + tap - half, slab] . W[n-tile, tap*d + slab]^T
//
// * operands are bf16 "hi/lo" planes (x = hi + lo): NPASS = 3 issues hi.hi + lo.hi + hi.lo
//   per k-step, which reproduces an fp32 product to ~2^-16 (fp32-parity mode); NPASS = 1
//   issues hi.hi only (bf16 mode).  Accumulation is fp32 in TMEM.
// * A is a (B, T, d) row-major tensor fetched by rank-3 TMA boxes (1 x 128 rows x 32 cols,
//   64-byte swizzle); rows outside [0, T) of an utterance are zero-filled by TMA, which is
//   exactly Conv1d's zero "same" padding, so Conv1d(d -> n, k) is k shifted GEMMs into one
//   accumulator with no im2col.  W is an (n, taps*d) row-major tensor.
// * persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA
//   issuer (one elected thread) + TMEM allocator, warps 2..5 = epilogue (thread = output row).
//   4-stage smem ring (full/empty mbarriers), double-buffered TMEM accumulator so the
//   epilogue of tile i overlaps the MMAs of tile i+1.
// * epilogues (fused, thread-per-row, fp32): + bias, ReLU, + residual, LayerNorm over the
//   full row (two-pass, statistics in fp32, staged through TMEM), outputs as fp32 and/or
//   as bf16 hi/lo planes for the next GEMM.
#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kBM = 128;     // rows per tile (UMMA M)
constexpr int kBK = 32;      // k-slab: 32 bf16 = 64 bytes = one SWIZZLE_64B row
constexpr int kStages = 4;
constexpr int kGemmTcThreads = 192;

struct GemmTcParams {
  int batch, t, d, taps, half;  // A is (batch, t, d); K = taps * d
  int n;                        // output columns
  int m_tiles_per_batch, n_tiles, total_tiles;
  const float* bias;            // (n) or null
  int relu;
  const float* residual;        // (batch*t, n) or null   [LN mode]
  const float* gamma;           // non-null => LayerNorm epilogue (requires n == N_TILE)
  const float* beta;
  float eps;
  float* out_f32;               // (batch*t, n) or null
  __nv_bfloat16* out_hi;        // (batch*t, n) or null
  __nv_bfloat16* out_lo;
};

template <int N_TILE, int NPASS>
struct SmemLayout {
  static constexpr int kAPlane = kBM * kBK * 2;        // 8 KB
  static constexpr int kWPlane = N_TILE * kBK * 2;
  static constexpr int kStage = (NPASS == 3 ? 2 : 1) * (kAPlane + kWPlane);
  static constexpr int kTotal = kStages * kStage + 1024;  // + alignment slack
};

template <int N_TILE, int NPASS, bool LN>
__global__ void __launch_bounds__(kGemmTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const GemmTcParams p) {
  using L = SmemLayout<N_TILE, NPASS>;
  constexpr int kAccCols = (N_TILE <= 32) ? 32 : (N_TILE <= 64) ? 64 : (N_TILE <= 128) ? 128 : 256;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  constexpr int kChunks = (N_TILE + 31) / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_slabs = p.taps * (p.d / kBK);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_w_hi);
    if (NPASS == 3) {
      prefetch_tmap(&map_a_lo);
      prefetch_tmap(&map_w_lo);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        int b = m_tile / p.m_tiles_per_batch, t0 = (m_tile % p.m_tiles_per_batch) * kBM;
        int n0 = n_tile * N_TILE;
        for (int ks = 0; ks < k_slabs; ++ks) {
          int tap = ks / (p.d / kBK), c0 = (ks % (p.d / kBK)) * kBK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStage;
          mbar_expect_tx(&full_bar[stage], L::kStage);
          tma_load_3d(st, &map_a_hi, &full_bar[stage], c0, t0 + tap - p.half, b);
          tma_load_3d(st + L::kAPlane, &map_w_hi, &full_bar[stage], tap * p.d + c0, n0, 0);
          if (NPASS == 3) {
            tma_load_3d(st + L::kAPlane + L::kWPlane, &map_a_lo, &full_bar[stage], c0, t0 + tap - p.half, b);
            tma_load_3d(st + 2 * L::kAPlane + L::kWPlane, &map_w_lo, &full_bar[stage], tap * p.d + c0, n0, 0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, kBM, N_TILE, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        int acc = it & 1;
        uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int ks = 0; ks < k_slabs; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint32_t sa_hi = smem_u32(smem + stage * L::kStage);
          uint32_t sw_hi = sa_hi + L::kAPlane;
          uint32_t sa_lo = sw_hi + L::kWPlane;
          uint32_t sw_lo = sa_lo + L::kAPlane;
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16) {
            uint32_t off = k16 * 32;  // 16 bf16 = 32 bytes along K inside the 64-byte swizzled row
            uint64_t a_hi = make_smem_desc(sa_hi + off, 16, 512, kSwizzle64);
            uint64_t w_hi = make_smem_desc(sw_hi + off, 16, 512, kSwizzle64);
            umma_f16(d_tmem, a_hi, w_hi, idesc, (ks | k16) ? 1u : 0u);
            if (NPASS == 3) {
              uint64_t a_lo = make_smem_desc(sa_lo + off, 16, 512, kSwizzle64);
              uint64_t w_lo = make_smem_desc(sw_lo + off, 16, 512, kSwizzle64);
              umma_f16(d_tmem, a_lo, w_hi, idesc, 1u);
              umma_f16(d_tmem, a_hi, w_lo, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: warps 2..5, thread = one output row =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      int acc = it & 1;
      uint32_t acc_phase = (it >> 1) & 1;
      int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
      int b = m_tile / p.m_tiles_per_batch, t0 = (m_tile % p.m_tiles_per_batch) * kBM;
      int n0 = n_tile * N_TILE;
      int tt = t0 + row_in_tile;
      bool row_ok = tt < p.t;
      size_t row = (size_t)b * p.t + tt;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      uint32_t taddr = tmem_base + acc * kAccCols + ((uint32_t)(quad * 32) << 16);
      float v[32];

      float mean = 0.f, rstd = 1.f;
      if (LN) {
        // pass 1: v = acc + bias (relu) + residual, written back to TMEM; row sum
        float s = 0.f;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
          int col0 = n0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j] + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
            if (p.relu) x = fmaxf(x, 0.f);
            v[j] = x;
          }
          if (p.residual && row_ok) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + row * p.n + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r = r4[j];
              v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) s += v[j];
          tmem_st32(taddr + c * 32, v);
        }
        tmem_wait_st();
        mean = s * (1.f / N_TILE);
        float q = 0.f;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float dlt = v[j] - mean;
            q = fmaf(dlt, dlt, q);
          }
        }
        rstd = rsqrtf(q * (1.f / N_TILE) + p.eps);
      }

#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        tmem_ld32(taddr + c * 32, v);
        int col0 = n0 + c * 32;
        if (LN) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = (v[j] - mean) * rstd * __ldg(p.gamma + col0 + j) + __ldg(p.beta + col0 + j);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int col = col0 + j;
            float x = v[j] + ((p.bias && col < p.n) ? __ldg(p.bias + col) : 0.f);
            if (p.relu) x = fmaxf(x, 0.f);
            v[j] = x;
          }
        }
        if (row_ok) {
          // n is a multiple of 16: a 32-column chunk is either fully or half inside
          int ncols = min(32, p.n - col0);
          if (ncols > 0) {
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(p.out_f32 + row * p.n + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (4 * j < ncols) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.out_hi) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v[2 * j], h0, l0);
                split_bf16(v[2 * j + 1], h1, l1);
                hi[j] = pack_bf16(h0, h1);
                lo[j] = pack_bf16(l0, l1);
              }
              uint4* oh = reinterpret_cast<uint4*>(p.out_hi + row * p.n + col0);
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo + row * p.n + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (8 * j < ncols) {
                  oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                  ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------- host
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0,
                  uint32_t box1, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int N_TILE, int NPASS, bool LN>
static int launch_gemm_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                          const GemmTcParams& p, cudaStream_t s) {
  using L = SmemLayout<N_TILE, NPASS>;
  auto kern = gemm_tc_kernel<N_TILE, NPASS, LN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) {
      set_error("gemm_tc: cannot reserve %d bytes of shared memory", L::kTotal);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  kern<<<grid, kGemmTcThreads, L::kTotal, s>>>(ah, al, wh, wl, p);
  LFS2_CHECK_LAUNCH("gemm_tc");
  return LFS2_OK;
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" {

int lfs2_gemm_tc(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps, const void* w_hi,
                 const void* w_lo, int n, const float* bias, int relu, const float* residual, const float* gamma,
                 const float* beta, float eps, float* out_f32, void* out_hi, void* out_lo, int npass, void* stream) {
  LFS2_REQUIRE(a_hi && w_hi, LFS2_ERR_INVALID_ARG, "gemm_tc: null operand");
  LFS2_REQUIRE(npass == 1 || npass == 3, LFS2_ERR_INVALID_ARG, "gemm_tc: npass must be 1 or 3");
  LFS2_REQUIRE(npass == 1 || (a_lo && w_lo), LFS2_ERR_INVALID_ARG, "gemm_tc: npass=3 needs the lo planes");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && n > 0 && taps > 0, LFS2_ERR_INVALID_ARG, "gemm_tc: bad shape");
  LFS2_REQUIRE(taps % 2 == 1, LFS2_ERR_UNSUPPORTED, "gemm_tc: kernel size %d must be odd", taps);
  LFS2_REQUIRE(d % kBK == 0, LFS2_ERR_UNSUPPORTED, "gemm_tc: d=%d must be a multiple of %d", d, kBK);
  LFS2_REQUIRE(n % 16 == 0, LFS2_ERR_UNSUPPORTED, "gemm_tc: n=%d must be a multiple of 16", n);
  LFS2_REQUIRE(out_f32 || out_hi, LFS2_ERR_INVALID_ARG, "gemm_tc: no output");
  LFS2_REQUIRE(!out_hi || out_lo, LFS2_ERR_INVALID_ARG, "gemm_tc: out_hi without out_lo");
  LFS2_REQUIRE(aligned16(a_hi) && aligned16(w_hi) && (!a_lo || aligned16(a_lo)) && (!w_lo || aligned16(w_lo)) &&
                   (!out_f32 || aligned16(out_f32)) && (!out_hi || (aligned16(out_hi) && aligned16(out_lo))) &&
                   (!residual || aligned16(residual)),
               LFS2_ERR_INVALID_ARG, "gemm_tc: pointers must be 16-byte aligned");
  const bool ln = gamma != nullptr;
  LFS2_REQUIRE(!ln || beta, LFS2_ERR_INVALID_ARG, "gemm_tc: gamma without beta");
  LFS2_REQUIRE(ln || !residual, LFS2_ERR_UNSUPPORTED, "gemm_tc: residual is only fused with LayerNorm");
  int n_tile = ln ? n : (n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : (n <= 128 ? 128 : 0)));
  if (ln) LFS2_REQUIRE(n == 256, LFS2_ERR_UNSUPPORTED, "gemm_tc: LayerNorm epilogue needs n == 256 (got %d)", n);
  LFS2_REQUIRE(n_tile != 0, LFS2_ERR_UNSUPPORTED, "gemm_tc: n=%d not tileable", n);

  CUtensorMap ah, al, wh, wl;
  const uint64_t ktot = (uint64_t)taps * d;
  bool ok = make_tmap_3d(&ah, a_hi, d, t, batch, kBK, kBM, 64) && make_tmap_3d(&wh, w_hi, ktot, n, 1, kBK, n_tile, 64);
  if (npass == 3)
    ok = ok && make_tmap_3d(&al, a_lo, d, t, batch, kBK, kBM, 64) && make_tmap_3d(&wl, w_lo, ktot, n, 1, kBK, n_tile, 64);
  else {
    al = ah;
    wl = wh;
  }
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed (driver entry point %s)",
               get_encode_tiled() ? "found" : "missing");

  GemmTcParams p;
  p.batch = batch; p.t = t; p.d = d; p.taps = taps; p.half = (taps - 1) / 2;
  p.n = n;
  p.m_tiles_per_batch = (t + kBM - 1) / kBM;
  p.n_tiles = (n + n_tile - 1) / n_tile;
  p.total_tiles = batch * p.m_tiles_per_batch * p.n_tiles;
  p.bias = bias; p.relu = relu; p.residual = residual; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.out_f32 = out_f32; p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
  cudaStream_t s = (cudaStream_t)stream;
  if (ln) {
    return npass == 3 ? launch_gemm_tc<256, 3, true>(ah, al, wh, wl, p, s)
                      : launch_gemm_tc<256, 1, true>(ah, al, wh, wl, p, s);
  }
  if (n_tile == 256)
    return npass == 3 ? launch_gemm_tc<256, 3, false>(ah, al, wh, wl, p, s)
                      : launch_gemm_tc<256, 1, false>(ah, al, wh, wl, p, s);
  return npass == 3 ? launch_gemm_tc<128, 3, false>(ah, al, wh, wl, p, s)
                    : launch_gemm_tc<128, 1, false>(ah, al, wh, wl, p, s);
}

// x (n) fp32 -> hi/lo bf16 planes
__global__ void split_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = x[i];
  __nv_bfloat16 h[4], l[4];
  split_bf16(v.x, h[0], l[0]);
  split_bf16(v.y, h[1], l[1]);
  split_bf16(v.z, h[2], l[2]);
  split_bf16(v.w, h[3], l[3]);
  hi[i] = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
  lo[i] = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
}

int lfs2_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream) {
  LFS2_REQUIRE(x && hi && lo, LFS2_ERR_INVALID_ARG, "split_bf16: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "split_bf16: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(x) && aligned16(hi) && aligned16(lo), LFS2_ERR_INVALID_ARG,
               "split_bf16: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  split_bf16_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n4);
  LFS2_CHECK_LAUNCH("split_bf16");
  return LFS2_OK;
}

}  // extern "C"
