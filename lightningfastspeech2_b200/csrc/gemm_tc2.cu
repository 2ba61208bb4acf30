// General batched tcgen05 GEMM with per-operand majorness -- the tensor-core kernel behind the
// weight gradients and the attention products of the train step:
//
//     C[z] (m x n, fp32)  (+)=  A[z] (m x k) . B[z] (n x k)^T          z = b * nhead + h
//
// Each operand is a window of a row-major bf16 tensor described by a rank-3 TMA map and is either
//   K-major  : the tensor's rows are the operand's m (or n) index, its columns the contraction index k
//              (activations x weights, Q.K^T, dS.K^T ...), one [32 k x rows] box per stage, or
//   MN-major : the tensor's ROWS are the contraction index and its columns the m (or n) index
//              (dY^T.X weight gradients, P.V, P^T.dO, dS^T.Q -- no transposed copy is ever made),
//              fetched as [32 k-rows x 32 columns] boxes, one per 32-wide column block.
// Both forms use the 64-byte swizzle; the UMMA shared-memory descriptors differ only in
// (LBO, SBO, k16 advance) = (16, 512, +32 B) for K-major and (box bytes, 512, +1024 B) for
// MN-major (conventions verified on hardware by tools/probe/probe_umma.cu).
// Heads of a packed (B, T, 3d) qkv tensor are column windows: column offset = col0 + h * hstride.
//
// Operands are bf16 hi/lo planes; NPASS = 3 issues hi.hi + lo.hi + hi.lo (fp32-parity), NPASS = 1
// hi.hi.  Accumulation is fp32 in TMEM.  The contraction can be split across CTAs (grid.y): the
// epilogue then reduces into C with fp32 vector atomics (weight gradients accumulate into p.grad).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue (thread = output row).
// MH = 2 (m >= 256): a CTA owns TWO 128-row accumulators of the same column tile (2 x 256 tensor-memory columns) and
// issues every B slab against both A halves.  With hi/lo planes on both operands a 128 x 256 tile moves 48 KB from L2
// per 768 MMA cycles = 62 B/clk/SM, above what the L2 delivers per SM (~43 B/clk with all 148 SMs pulling: the weight
// gradients ran at 53 % of the 3-pass issue rate); 256 x 256 moves 64 KB per 1536 cycles = 42 B/clk.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kG2M = 128;         // UMMA M
constexpr int kG2K = 32;          // contraction elements per stage
constexpr int kG2Threads = 192;
constexpr int kG2BoxBytes = kG2K * 64;  // one MN-major box: 32 k-rows x 64 B

struct G2Operand {
  int col0, hstride, per_z;  // column offset of head h: col0 + h * hstride; batch coordinate: z (per_z) or b
};

struct G2Params {
  int m, n, k, nhead;
  G2Operand a, b;
  float* c;
  long long c_bstride, c_hstride;
  int ldc;
  int k_per_split;  // multiple of kG2K
  int n_tiles;
  int atomic;
};

template <int N_TILE, int NPASS, int MH, int OCC>
struct G2Smem {
  static constexpr int kAHalf = kG2M * kG2K * 2;     // 8 KB: one 128-row half of the A slab
  static constexpr int kAPlane = MH * kAHalf;
  static constexpr int kBPlane = N_TILE * kG2K * 2;  // 8 / 16 KB
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kStage = kPlanes * (kAPlane + kBPlane);
  static constexpr int kOffBHi = kAPlane;
  static constexpr int kOffALo = kAPlane + kBPlane;
  static constexpr int kOffBLo = 2 * kAPlane + kBPlane;
  static constexpr int kBudget = (OCC == 2 ? 100 : 200) * 1024;  // OCC = 2: two CTAs share the SM (short contractions)
  static constexpr int kStages = (kBudget / kStage) > 8 ? 8 : (kBudget / kStage);
  static constexpr int kTotal = kStages * kStage + 1024;
  static_assert(kStages >= 2, "gemm_tc2: the pipeline needs two stages");
};

template <bool MN>
__device__ __forceinline__ void g2_load(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col_off, int row0,
                                        int rows, int k0, int c2) {
  if (MN) {
    for (int j = 0; j < rows / 32; ++j) tma_load_3d(dst + j * kG2BoxBytes, map, bar, col_off + row0 + 32 * j, k0, c2);
  } else {
    tma_load_3d(dst, map, bar, col_off + k0, row0, c2);
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool A_MN, bool B_MN, int N_TILE, int NPASS, int MH, int OCC>
__global__ void __launch_bounds__(kG2Threads, OCC)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const G2Params p) {
  using L = G2Smem<N_TILE, NPASS, MH, OCC>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = MH * N_TILE <= 128 ? 128 : (MH * N_TILE <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, z = blockIdx.z;
  const int m0 = (tile / p.n_tiles) * (kG2M * MH), n0 = (tile % p.n_tiles) * N_TILE;
  const int h = z % p.nhead, b = z / p.nhead;
  const int k_begin = split * p.k_per_split;
  const int k_end = min(p.k, k_begin + p.k_per_split);
  const int nstages = (k_end - k_begin + kG2K - 1) / kG2K;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_b_hi);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0 && nstages > 0) {
      const int ca = p.a.col0 + h * p.a.hstride, cb = p.b.col0 + h * p.b.hstride;
      const int za = p.a.per_z ? z : b, zb = p.b.per_z ? z : b;
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < nstages; ++ks) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * L::kStage;
        const int k0 = k_begin + ks * kG2K;
        mbar_expect_tx(&full_bar[stage], L::kStage);
        g2_load<A_MN>(st, &map_a_hi, &full_bar[stage], ca, m0, kG2M * MH, k0, za);
        g2_load<B_MN>(st + L::kOffBHi, &map_b_hi, &full_bar[stage], cb, n0, N_TILE, k0, zb);
        if (NPASS == 3) {
          g2_load<A_MN>(st + L::kOffALo, &map_a_lo, &full_bar[stage], ca, m0, kG2M * MH, k0, za);
          g2_load<B_MN>(st + L::kOffBLo, &map_b_lo, &full_bar[stage], cb, n0, N_TILE, k0, zb);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kFmtBF16, kG2M, N_TILE, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t kStepA = A_MN ? 1024 : 32, kStepB = B_MN ? 1024 : 32;
    const uint64_t da0 = make_smem_desc(smem_u32(smem), A_MN ? kG2BoxBytes : 16, 512, kSwizzle64);
    const uint64_t db0 = make_smem_desc(smem_u32(smem), B_MN ? kG2BoxBytes : 16, 512, kSwizzle64);
    int stage = 0;
    uint32_t phase = 0;
    for (int ks = 0; ks < nstages; ++ks) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_hi = desc_advance(da0, stage * L::kStage);
        const uint64_t b_hi = desc_advance(db0, stage * L::kStage + L::kOffBHi);
        const uint64_t a_lo = desc_advance(da0, stage * L::kStage + L::kOffALo);
        const uint64_t b_lo = desc_advance(db0, stage * L::kStage + L::kOffBLo);
#pragma unroll
        for (int s = 0; s < kG2K / 16; ++s) {
#pragma unroll
          for (int mh = 0; mh < MH; ++mh) {  // both row halves against the same B slab
            const uint32_t acc = tmem_base + mh * N_TILE;
            const uint32_t oa = mh * L::kAHalf + s * kStepA;
            if (ks == 0 && s == 0) umma_f16_c<false>(acc, desc_advance(a_hi, oa), b_hi, idesc);
            else umma_f16_c<true>(acc, desc_advance(a_hi, oa), desc_advance(b_hi, s * kStepB), idesc);
            if (NPASS == 3) {
              umma_f16_c<true>(acc, desc_advance(a_lo, oa), desc_advance(b_hi, s * kStepB), idesc);
              umma_f16_c<true>(acc, desc_advance(a_hi, oa), desc_advance(b_lo, s * kStepB), idesc);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (ks + 1 == nstages) umma_commit(&acc_full);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (nstages > 0) {
    // ===================== epilogue: warps 2..5, thread = output row =====================
    // A lane holds 32 consecutive columns of ITS row after tcgen05.ld; written out like that, every store instruction
    // touches 32 rows x 16 bytes (half-filled sectors: the T x T logits of the attention products left at 1.5 TB/s).
    // Each warp transposes its 32 x 32 chunk through a padded shared-memory tile instead -- the pipeline stages are idle
    // once acc_full has fired (every TMA write landed, every MMA read its operands) -- so that 8 lanes cover 128
    // contiguous bytes of one row.
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + quad * (32 * 33);
    float* cbase = p.c + (long long)b * p.c_bstride + (long long)h * p.c_hstride;
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    float v[32];
#pragma unroll 1
    for (int cc = 0; cc < MH * (N_TILE / 32); ++cc) {
      const int mh = cc / (N_TILE / 32), c = cc % (N_TILE / 32);
      const int rbase = m0 + mh * kG2M + quad * 32;
      if (m0 + mh * kG2M >= p.m) break;  // warp-uniform: the second half of the last row tile may be empty
      const uint32_t taddr = tmem_base + mh * N_TILE + ((uint32_t)(quad * 32) << 16);
      const int col0 = n0 + c * 32;
      if (col0 >= p.n) continue;
      tmem_ld32(taddr + c * 32, v);
      if (col0 + 32 <= p.n && (p.ldc & 3) == 0) {  // warp-uniform
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + sub;
          if (rbase + rr < p.m) {
            const float* sp = stg + rr * 33 + c4;
            float* dst = cbase + (long long)(rbase + rr) * p.ldc + col0 + c4;
            if (p.atomic) red_add_v4(dst, sp[0], sp[1], sp[2], sp[3]);
            else *reinterpret_cast<float4*>(dst) = make_float4(sp[0], sp[1], sp[2], sp[3]);
          }
        }
      } else if (rbase + lane < p.m) {
        float* crow = cbase + (long long)(m0 + mh * kG2M + r) * p.ldc;
        for (int j = 0; j < 32 && col0 + j < p.n; ++j) {
          if (p.atomic) atomicAdd(crow + col0 + j, v[j]);
          else crow[col0 + j] = v[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool A_MN, bool B_MN, int N_TILE, int NPASS, int MH, int OCC>
static int launch_g2(const CUtensorMap* maps, const G2Params& p, dim3 grid, cudaStream_t s) {
  using L = G2Smem<N_TILE, NPASS, MH, OCC>;
  auto kern = gemm_tc2_kernel<A_MN, B_MN, N_TILE, NPASS, MH, OCC>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) {
      set_error("gemm_tc2: cannot reserve %d bytes of shared memory", L::kTotal);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  kern<<<grid, kG2Threads, L::kTotal, s>>>(maps[0], maps[1], maps[2], maps[3], p);
  LFS2_CHECK_LAUNCH("gemm_tc2");
  return LFS2_OK;
}

template <bool A_MN, bool B_MN, int MH, int OCC>
static int dispatch_g2_mh(const CUtensorMap* maps, const G2Params& p, int n_tile, int npass, dim3 grid, cudaStream_t s) {
  if (n_tile == 256)
    return npass == 3 ? launch_g2<A_MN, B_MN, 256, 3, MH, OCC>(maps, p, grid, s)
                      : launch_g2<A_MN, B_MN, 256, 1, MH, OCC>(maps, p, grid, s);
  return npass == 3 ? launch_g2<A_MN, B_MN, 128, 3, MH, OCC>(maps, p, grid, s)
                    : launch_g2<A_MN, B_MN, 128, 1, MH, OCC>(maps, p, grid, s);
}

// mh = 2: two row halves per CTA, one CTA per SM (long contractions: weight gradients);
// mh = 1, occ = 2: single tiles, two CTAs per SM so that one's epilogue runs under the other's MMAs (attention products)
template <bool A_MN, bool B_MN>
static int dispatch_g2(const CUtensorMap* maps, const G2Params& p, int n_tile, int npass, int mh, int occ, dim3 grid,
                       cudaStream_t s) {
  if (mh == 2) return dispatch_g2_mh<A_MN, B_MN, 2, 1>(maps, p, n_tile, npass, grid, s);
  return occ == 2 ? dispatch_g2_mh<A_MN, B_MN, 1, 2>(maps, p, n_tile, npass, grid, s)
                  : dispatch_g2_mh<A_MN, B_MN, 1, 1>(maps, p, n_tile, npass, grid, s);
}

// Tile shape per launch.  Long contractions (>= 64 stages per CTA: the weight gradients) take two 128-row halves per
// CTA; short ones (attention products: k = head_dim or T) keep single tiles but let two CTAs share the SM -- there the
// epilogue (a 128 x 256 fp32 tile leaves through per-row stores) costs as much as the MMAs and has to overlap with them.
// LFS2_G2_MH = 1|2 / LFS2_G2_OCC = 1|2 in the environment force a shape (A/B runs).
static int g2_env(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}
static void g2_shape(int m, int k, int* mh, int* occ) {
  static const int forced_mh = g2_env("LFS2_G2_MH"), forced_occ = g2_env("LFS2_G2_OCC");
  *mh = (m >= 2 * kG2M && k >= 64 * kG2K) ? 2 : 1;
  if (forced_mh == 1 || forced_mh == 2) *mh = forced_mh;
  *occ = *mh == 1 ? 2 : 1;
  if (*mh == 1 && (forced_occ == 1 || forced_occ == 2)) *occ = forced_occ;
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" {

int lfs2_gemm_tc2(const void* a_hi, const void* a_lo, const lfs2_operand* a, const void* b_hi, const void* b_lo,
                  const lfs2_operand* b, float* c, int ldc, long long c_bstride, long long c_hstride, int m, int n,
                  int k, int nbatch, int nhead, int npass, int accumulate, void* stream) {
  LFS2_REQUIRE(a_hi && b_hi && a && b && c, LFS2_ERR_INVALID_ARG, "gemm_tc2: null pointer");
  LFS2_REQUIRE(npass == 1 || npass == 3, LFS2_ERR_INVALID_ARG, "gemm_tc2: npass must be 1 or 3");
  LFS2_REQUIRE(npass == 1 || (a_lo && b_lo), LFS2_ERR_INVALID_ARG, "gemm_tc2: npass=3 needs the lo planes");
  if (m == 0 || n == 0 || nbatch == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && n > 0 && k > 0 && nbatch > 0 && nhead > 0, LFS2_ERR_INVALID_ARG, "gemm_tc2: bad shape");
  LFS2_REQUIRE(a->d0 % 8 == 0 && b->d0 % 8 == 0, LFS2_ERR_UNSUPPORTED,
               "gemm_tc2: operand row pitch must be a multiple of 8 elements (TMA strides are 16-byte multiples)");
  LFS2_REQUIRE(aligned16(a_hi) && aligned16(b_hi) && (!a_lo || aligned16(a_lo)) && (!b_lo || aligned16(b_lo)) &&
                   (reinterpret_cast<uintptr_t>(c) & 3u) == 0,
               LFS2_ERR_INVALID_ARG, "gemm_tc2: pointers must be 16-byte aligned");
  LFS2_REQUIRE((long long)nbatch * nhead <= 65535, LFS2_ERR_UNSUPPORTED, "gemm_tc2: batch * heads exceeds the grid limit");
  const int n_tile = n > 128 ? 256 : 128;
  int mh, occ;
  g2_shape(m, k, &mh, &occ);
  // K-major operands must not run past their k window into a neighbouring head: k is a multiple of the stage
  LFS2_REQUIRE((a->mn_major || a->hstride == 0 || k % kG2K == 0) && (b->mn_major || b->hstride == 0 || k % kG2K == 0),
               LFS2_ERR_UNSUPPORTED, "gemm_tc2: k=%d must be a multiple of %d for head-windowed K-major operands", k, kG2K);

  CUtensorMap maps[4];
  auto mk = [&](CUtensorMap* out, const void* base, const lfs2_operand* o, int rows_tile) {
    if (o->mn_major) return make_tmap_3d(out, base, o->d0, o->d1, o->d2, 32, kG2K, 64);
    return make_tmap_3d(out, base, o->d0, o->d1, o->d2, kG2K, rows_tile, 64);
  };
  bool ok = mk(&maps[0], a_hi, a, kG2M * mh) && mk(&maps[2], b_hi, b, n_tile);
  if (npass == 3) ok = ok && mk(&maps[1], a_lo, a, kG2M * mh) && mk(&maps[3], b_lo, b, n_tile);
  else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "gemm_tc2: cuTensorMapEncodeTiled failed");

  G2Params p;
  p.m = m; p.n = n; p.k = k; p.nhead = nhead;
  p.a = {a->col0, a->hstride, a->per_z};
  p.b = {b->col0, b->hstride, b->per_z};
  p.c = c; p.c_bstride = c_bstride; p.c_hstride = c_hstride; p.ldc = ldc;
  const int m_tiles = ceil_div(m, kG2M * mh);
  p.n_tiles = ceil_div(n, n_tile);
  const int z = nbatch * nhead;
  const long long ctas = (long long)m_tiles * p.n_tiles * z;
  int splits = 1;
  if (accumulate && ctas < 2 * num_sms()) {
    // split the contraction so that the grid fills (at most) two full waves of the 148 SMs: rounding the
    // split count DOWN keeps the CTA count <= 2 * 148 -- rounding up would leave a third, mostly empty wave
    splits = (int)((2 * num_sms()) / ctas);
    const int max_splits = ceil_div(k, 8 * kG2K);  // at least 8 stages per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  p.k_per_split = ceil_div(ceil_div(k, splits), kG2K) * kG2K;
  splits = ceil_div(k, p.k_per_split);
  p.atomic = accumulate ? 1 : 0;
  dim3 grid(m_tiles * p.n_tiles, splits, z);
  cudaStream_t s = (cudaStream_t)stream;
  if (a->mn_major)
    return b->mn_major ? dispatch_g2<true, true>(maps, p, n_tile, npass, mh, occ, grid, s)
                       : dispatch_g2<true, false>(maps, p, n_tile, npass, mh, occ, grid, s);
  return b->mn_major ? dispatch_g2<false, true>(maps, p, n_tile, npass, mh, occ, grid, s)
                     : dispatch_g2<false, false>(maps, p, n_tile, npass, mh, occ, grid, s);
}

}  // extern "C"
