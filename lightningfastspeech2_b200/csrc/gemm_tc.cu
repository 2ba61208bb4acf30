// tcgen05 GEMM / implicit-GEMM Conv1d of the FFTBlock and predictor stacks:
//
//     acc[128 x N_TILE] (TMEM, fp32) = sum over (tap, k-slab)  A[rows + tap - half, slab] . W[n-tile, tap*d + slab]^T
//                                      (+ residual[rows, n-tile] . I   -- the residual add rides the tensor pipe)
//
// * operands are bf16 "hi/lo" planes (x = hi + lo): NPASS = 3 issues hi.hi + lo.hi + hi.lo
//   per k-step, which reproduces an fp32 product to ~2^-16 (fp32-parity mode); NPASS = 1
//   issues hi.hi only (bf16 mode).  Accumulation is fp32 in TMEM.
// * A is a (B, T, d) row-major tensor fetched by rank-3 TMA boxes (1 x 128 rows x 32 cols,
//   64-byte swizzle); rows outside [0, T) of an utterance are zero-filled by TMA, which is
//   exactly Conv1d's zero "same" padding, so Conv1d(d -> n, k) is k shifted GEMMs into one
//   accumulator with no im2col.  W is an (n, taps*d) row-major tensor.
// * the residual of the post-norm FFTBlock (x + f(x)) is added by the tensor core: its hi/lo
//   planes are streamed as extra k-slabs against an identity weight (hi.I + lo.I, exact in the
//   fp32 accumulator), so the epilogue never issues a strided global load.
// * persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA
//   issuer (one elected thread) + TMEM allocator, warps 2..5 = epilogue (thread = output row).
//   smem ring (full/empty mbarriers), double-buffered TMEM accumulator so the epilogue of
//   tile i overlaps the MMAs of tile i+1.
// * epilogue (fp32, thread-per-row): + bias, ReLU, LayerNorm over the full row (statistics in
//   one TMEM pass, normalisation in a second), then 32-column chunks are written to a
//   swizzled shared-memory staging buffer and leave as coalesced TMA tensor stores -- either
//   fp32 or bf16 hi/lo planes for the next GEMM.  TMA clips rows/columns outside the tensor.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kBM = 128;     // rows per tile (UMMA M)
constexpr int kBK = 32;      // k-slab: 32 bf16 = 64 bytes = one SWIZZLE_64B row
constexpr int kGemmTcThreads = 320;    // TMA + MMA warps, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kGemmTcThreadsWide = 576;  // 256-column tiles: 16 epilogue warps (four per quadrant), see the epilogue below
__host__ __device__ constexpr int gemm_tc_threads(int n_tile) { return n_tile == 256 ? kGemmTcThreadsWide : kGemmTcThreads; }
constexpr int kStageChunk = kBM * 32 * 4;  // one 128 x 32 staging chunk: 16 KB (fp32) or 2 x 8 KB (hi | lo)

// epilogue output: bf16 hi/lo planes | fp32 | ONE fp16 plane | ONE bf16 plane
enum { kOutPlanes = 0, kOutF32 = 1, kOutF16 = 2, kOutBF16 = 3 };

struct GemmTcParams {
  int batch, t, d, taps, half;  // A is (batch, t, d); K = taps * d
  int dil;                      // tap spacing in rows (dilated Conv1d); 1 otherwise
  int n;                        // output columns
  int m_tiles_per_batch, n_tiles, total_tiles;
  int has_residual;             // residual planes (batch, t, n) ride as extra k-slabs against I (n x n)
  const float* bias;            // (n) or null
  int relu;                     // 0 none, 1 ReLU, 2 leaky ReLU with `slope`
  float slope;
  const uint8_t* row_mask;      // null, or (batch, t): rows with a non-zero byte are written as zeros (PAD frames of a
                                //   ragged batch whose convolutions must see zeros there, like a per-utterance call would)
  const float* gamma;           // LN only (n == N_TILE)
  const float* beta;
  float eps;
  const int* tile_list;         // null, or [0] = number of active m-tiles, [1..] = their indices (b * m_tiles_per_batch + mt):
                                //   only those row tiles are processed (lfs2_gemm_tc_limited), dealt round-robin to the CTAs
  int m_step, t_shift;          // row tile i of an utterance starts at i * m_step + t_shift (128 / 0; 126 / -1 for kEpiStencil)
  // fused predictor epilogues (lfs2_predictor_layer_tc): see EPI below
  const float* st_w;            // kEpiStencil: (3, n) taps of the NEXT layer's depthwise conv, tap-major; st_b its bias (n)
  const float* st_b;
  const float* dot_w;           // kEpiDot: (n) head weight, dot_b (1) its bias, dot_mask (batch, t) or null, dot_out (batch, t)
  const float* dot_b;
  const uint8_t* dot_mask;
  float* dot_out;
};
// EPI: what the LayerNorm epilogue does with the normalised row z
//   kEpiNone     store it (every other use of the kernel)
//   kEpiStencil  store u = depthwise3(z) -- the k = 3 depthwise conv of the NEXT predictor layer, rows z[r-1], z[r], z[r+1]
//                taken from the neighbouring epilogue threads (warp shuffles; a warp's first and last row through shared
//                memory).  A tile's first and last row have no neighbour inside the tile, so tiles advance by 126 rows
//                and store rows 1..126; z rows outside the utterance are zeros (Conv1d's padding).
//   kEpiDot      store only out[row] = z . dot_w + dot_b (masked): the predictor head (model.py:512-518)
enum { kEpiNone = 0, kEpiStencil = 1, kEpiDot = 2 };

// active m-tiles of a row-limited launch: utterance b needs the tiles that start before row_limit[b] + extra
__global__ void gemm_tile_list_kernel(const int* __restrict__ row_limit, int extra, int batch, int m_tiles_per_batch,
                                      int* __restrict__ list, int m_step) {
  for (int i = threadIdx.x; i < batch * m_tiles_per_batch; i += blockDim.x) {
    const int b = i / m_tiles_per_batch, mt = i % m_tiles_per_batch;
    if (mt * m_step < row_limit[b] + extra) list[1 + atomicAdd(list, 1)] = i;
  }
}

template <int N_TILE, int NPASS, bool LN, int EPI = 0, bool MC = false>
struct SmemLayout {
  // kEpiStencil: taps w0 | w1 | w2 | bias, then the warp-edge rows of the neighbour exchange: [4 groups][2 chunk
  // parities][4 quadrants][first | last row][32 floats] = 8 KB; kEpiDot: the head weight
  static constexpr int kEdgeRows = EPI == 1 ? 4 * 2 * 4 * 2 * 32 * 4 : 0;
  static constexpr int kEdge = EPI != 0 ? 4 * N_TILE * 4 + kEdgeRows : 0;
  static constexpr bool kHasLo = NPASS == 3 || LN;  // LN variants stream residual lo planes even in bf16 mode
  static constexpr int kAPlane = kBM * kBK * 2;     // 8 KB
  static constexpr int kWPlane = N_TILE * kBK * 2 / (MC ? 2 : 1);  // CTA pair: each CTA holds half of the weight slab
  static constexpr int kStage = (kHasLo ? 2 : 1) * kAPlane + (NPASS >= 2 ? 2 : 1) * kWPlane;
  static constexpr int kOffWHi = kAPlane;
  static constexpr int kOffALo = kAPlane + kWPlane;
  static constexpr int kOffWLo = (kHasLo ? 2 : 1) * kAPlane + kWPlane;
  static constexpr int kI32 = kHasLo ? 32 * kBK * 2 : 0;               // 32 x 32 identity block of the residual products (pair: 16 rows each)
  static constexpr int kFixed = 4 * kStageChunk + kI32 + (LN ? 3 * N_TILE * 4 + 2 * 4 * kBM * 8 : 0) + kEdge + 1024;
  static constexpr int kStages = (226 * 1024 - kFixed) / kStage > 6 ? 6 : (226 * 1024 - kFixed) / kStage;
  static constexpr int kOffStaging = kStages * kStage;                 // 2 halves x 2 chunks, 1024-aligned
  static constexpr int kOffI32 = kOffStaging + 4 * kStageChunk;        // 1024-aligned (swizzled operand tile)
  static constexpr int kOffVec = kOffI32 + kI32;                       // bias | gamma | beta for LN: 3 * N_TILE floats
  static constexpr int kOffStats = kOffVec + 3 * N_TILE * 4;           // LN partial (sum, sumsq): [2 tiles][4 groups][128]
  static constexpr int kOffEdge = kOffStats + 2 * 4 * kBM * 8;         // fused-epilogue vectors (kEdge bytes)
  static constexpr int kTotal = kStages * kStage + kFixed;
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
  static_assert(kStage % 1024 == 0, "stage must keep 1024-byte alignment");
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- CTA-pair variant (MC): the two CTAs of a cluster work on two row tiles of the SAME column tile as ONE
// tcgen05.mma.cta_group::2 of M = 256 (tc_common.cuh): each CTA fetches its own A rows and HALF of every weight slab
// (its N_TILE/2 rows), the leader issues the MMAs, each CTA's accumulator rows land in its own tensor memory.  A K = 256
// GEMM re-reads its whole weight tile for every 128-row tile (262 KB of the 521 KB a tile moves L2 -> SM in fp32 mode),
// which is what bounds these kernels; the pair halves that part per SM (a multicast of the halves into both CTAs, the
// round-1 scheme, moved the same bytes into every SM and gained 1 %).
// work items of one CTA: item i -> (utterance b, first row t0, first column n0)
template <bool MC>
struct TileWalk {
  int first, stride, count, rank, m_count;
  __device__ __forceinline__ explicit TileWalk(const GemmTcParams& p) {
    rank = MC ? (int)cluster_ctarank() : 0;
    m_count = p.tile_list ? __ldg(p.tile_list) : p.batch * p.m_tiles_per_batch;
    if (MC) {  // a cluster takes a PAIR of row tiles of one column tile; an odd tail pairs with an out-of-range tile
      first = blockIdx.x >> 1;
      stride = gridDim.x >> 1;
      count = ((m_count + 1) >> 1) * p.n_tiles;
    } else {
      first = blockIdx.x;
      stride = gridDim.x;
      count = m_count * p.n_tiles;
    }
  }
  __device__ __forceinline__ void coords(const GemmTcParams& p, int i, int n_tile_cols, int& b, int& t0, int& n0) const {
    n0 = (i % p.n_tiles) * n_tile_cols;
    const int mi = MC ? 2 * (i / p.n_tiles) + rank : i / p.n_tiles;
    if (mi < m_count) {
      const int m_tile = p.tile_list ? __ldg(p.tile_list + 1 + mi) : mi;
      b = m_tile / p.m_tiles_per_batch;
      t0 = (m_tile % p.m_tiles_per_batch) * p.m_step + p.t_shift;
    } else {  // no such tile: TMA zero-fills loads and drops stores outside the tensor
      b = p.batch;
      t0 = 0;
    }
  }
};

__device__ __forceinline__ float activate(float x, int kind, float slope) {
  return kind == 1 ? fmaxf(x, 0.f) : (kind == 2 ? (x > 0.f ? x : x * slope) : x);
}
// v[0..32) <- LayerNorm row values of one 32-column chunk: ((act(v + bias) - mean) * rstd) * gamma + beta, the three
// vectors read from shared memory as broadcast 16-byte loads (vec = bias | gamma | beta, n floats each)
__device__ __forceinline__ void ln_chunk(float (&v)[32], const float* vec, int n, int col0, float mean, float rstd, int act,
                                         float slope) {
  const float4* b4 = reinterpret_cast<const float4*>(vec + col0);
  const float4* g4 = reinterpret_cast<const float4*>(vec + n + col0);
  const float4* e4 = reinterpret_cast<const float4*>(vec + 2 * n + col0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bb = b4[j], g = g4[j], bt = e4[j];
    v[4 * j] = (activate(v[4 * j] + bb.x, act, slope) - mean) * rstd * g.x + bt.x;
    v[4 * j + 1] = (activate(v[4 * j + 1] + bb.y, act, slope) - mean) * rstd * g.y + bt.y;
    v[4 * j + 2] = (activate(v[4 * j + 2] + bb.z, act, slope) - mean) * rstd * g.z + bt.z;
    v[4 * j + 3] = (activate(v[4 * j + 3] + bb.w, act, slope) - mean) * rstd * g.w + bt.w;
  }
}

template <int N_TILE, int NPASS, bool LN, int OUT, bool MC, int EPI = 0>
__global__ void __launch_bounds__(gemm_tc_threads(N_TILE), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
               const __grid_constant__ CUtensorMap map_ident, const __grid_constant__ CUtensorMap map_o0,
               const __grid_constant__ CUtensorMap map_o1, const GemmTcParams p) {
  static_assert(EPI == kEpiNone || (LN && !MC && OUT == kOutPlanes), "fused predictor epilogues ride the LayerNorm variant");
  static_assert(NPASS != 2 || !LN, "the 2-pass recipe (fp16 activation plane) has no LayerNorm / residual build");
  using L = SmemLayout<N_TILE, NPASS, LN, EPI, MC>;
  constexpr int kStages = L::kStages;
  constexpr int kAccCols = (N_TILE <= 32) ? 32 : (N_TILE <= 64) ? 64 : (N_TILE <= 128) ? 128 : 256;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  constexpr int kChunks = N_TILE / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2], ident_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_slabs = p.taps * (p.d / kBK);
  const int r_slabs = p.has_residual ? N_TILE / kBK : 0;
  const int k_slabs = a_slabs + r_slabs;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_w_hi);
    prefetch_tmap(&map_o0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);  // (pair: the leader's commit arrives here in both CTAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (N_TILE == 256 ? 16 : 8) * (MC ? 2 : 1));  // one arrive per epilogue warp (pair: of both CTAs, on the leader's barrier)
    }
    mbar_init(&ident_bar, 1);
    fence_barrier_init();
  }
  if (MC) cluster_sync_all();  // both CTAs are running and their barriers exist before the pair-wide allocation
  if (warp == 1) {
    if (MC) tmem_alloc_pair(&tmem_base_smem, kTmemCols);
    else tmem_alloc(&tmem_base_smem, kTmemCols);
  }
  if (LN && warp >= 2) {  // bias | gamma | beta -> shared (broadcast reads in the epilogue)
    float* vec = reinterpret_cast<float*>(smem + L::kOffVec);
    for (int i = threadIdx.x - 64; i < N_TILE; i += (int)blockDim.x - 64) {
      vec[i] = p.bias ? p.bias[i] : 0.f;
      vec[N_TILE + i] = p.gamma[i];
      vec[2 * N_TILE + i] = p.beta[i];
    }
    if (EPI != kEpiNone) {
      float* ev = reinterpret_cast<float*>(smem + L::kOffEdge);
      for (int i = threadIdx.x - 64; i < N_TILE; i += (int)blockDim.x - 64) {
        if (EPI == kEpiStencil) {
          ev[i] = p.st_w[i];
          ev[N_TILE + i] = p.st_w[N_TILE + i];
          ev[2 * N_TILE + i] = p.st_w[2 * N_TILE + i];
          ev[3 * N_TILE + i] = p.st_b[i];
        } else {
          ev[i] = p.dot_w[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's tensor memory is allocated before the leader's MMAs write into it
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const TileWalk<MC> walk(p);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // MC (CTA pair): the weight maps have half-height boxes; this CTA fetches rows [rank * N_TILE/2, +N_TILE/2) of a
      // slab into ITS shared memory; every load of both CTAs completes on the leader's barrier, which expects both shares
      const int w_row = MC ? walk.rank * (N_TILE / 2) : 0;
      const bool leader = !MC || walk.rank == 0;
      auto load = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
        if (MC) tma_load_3d_pair(dst, map, pair_leader_bar(bar), c0, c1, c2);
        else tma_load_3d(dst, map, bar, c0, c1, c2);
      };
      auto load_w = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
        load(dst, map, bar, c0, c1 + w_row, 0);
      };
      if (L::kHasLo && r_slabs > 0 && walk.first < walk.count) {  // the identity block of the residual products: once
        if (leader) mbar_expect_tx(&ident_bar, L::kI32);           // (pair: 16 of its 32 rows in each CTA)
        load(smem + L::kOffI32, &map_ident, &ident_bar, 0, MC ? walk.rank * 16 : 0, 0);
      }
      for (int tile = walk.first; tile < walk.count; tile += walk.stride) {
        int b, t0, n0;
        walk.coords(p, tile, N_TILE, b, t0, n0);
        for (int ks = 0; ks < k_slabs; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStage;
          if (ks < a_slabs) {
            int tap = ks / (p.d / kBK), c0 = (ks % (p.d / kBK)) * kBK;
#ifdef LFS2_DIAG_NO_LO_LOADS  // timing diagnostics only (tools/gemm_ab.py): wrong results
            if (leader) mbar_expect_tx(&full_bar[stage], (MC ? 2 : 1) * (L::kAPlane + L::kWPlane));
#else
            if (leader)
              mbar_expect_tx(&full_bar[stage],
                             (MC ? 2 : 1) * ((NPASS == 3 ? 2 : 1) * L::kAPlane + (NPASS >= 2 ? 2 : 1) * L::kWPlane));
#endif
            load(st, &map_a_hi, &full_bar[stage], c0, t0 + (tap - p.half) * p.dil, b);
            load_w(st + L::kOffWHi, &map_w_hi, &full_bar[stage], tap * p.d + c0, n0);
#ifdef LFS2_DIAG_NO_LO_LOADS
            if (false) {
#else
            if (NPASS >= 2) {
#endif
              if (NPASS == 3) load(st + L::kOffALo, &map_a_lo, &full_bar[stage], c0, t0 + (tap - p.half) * p.dil, b);
              load_w(st + L::kOffWLo, &map_w_lo, &full_bar[stage], tap * p.d + c0, n0);
            }
          } else if (L::kHasLo) {  // residual slab: R_hi, R_lo (against the resident 32 x 32 identity block)
            int c0 = n0 + (ks - a_slabs) * kBK;
            if (leader) mbar_expect_tx(&full_bar[stage], (MC ? 2 : 1) * 2 * L::kAPlane);
            load(st, &map_r_hi, &full_bar[stage], c0, t0, b);
            load(st + L::kOffALo, &map_r_lo, &full_bar[stage], c0, t0, b);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && (!MC || walk.rank == 0)) {
    // ===================== MMA issuer (pair: the leader CTA's, for both) =====================
    // whole warp runs the uniform control flow (descriptors stay in uniform registers), one
    // elected lane issues; per instruction the descriptor is a 64-bit add on a precomputed base
    // NPASS = 2: the activation operand is ONE fp16 plane against fp16 hi/lo weight planes (a.w_hi + a.w_lo)
    constexpr int kM = MC ? 2 * kBM : kBM;  // pair: one instruction covers both CTAs' row tiles
    constexpr uint32_t idesc = make_idesc(NPASS == 2 ? kFmtF16 : kFmtBF16, kM, N_TILE, 0, 0);
    constexpr uint32_t idesc_r = make_idesc(kFmtBF16, kM, 32, 0, 0);  // residual: one 32-column block per slab
    auto mma = [&](bool accumulate, uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
      if (MC) {
        if (accumulate) umma_f16_pair<true>(d, a, b, id);
        else umma_f16_pair<false>(d, a, b, id);
      } else {
        if (accumulate) umma_f16_c<true>(d, a, b, id);
        else umma_f16_c<false>(d, a, b, id);
      }
    };
    const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 512, kSwizzle64);
    const uint64_t d_i32 = make_smem_desc(smem_u32(smem + L::kOffI32), 16, 512, kSwizzle64);
    if (L::kHasLo && r_slabs > 0 && walk.first < walk.count) mbar_wait(&ident_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = walk.first; tile < walk.count; tile += walk.stride, ++it) {
      int acc = it & 1;
      uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      for (int ks = 0; ks < k_slabs; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_hi = desc_advance(d0, stage * L::kStage);
          const uint64_t w_hi = desc_advance(a_hi, L::kOffWHi);
          const uint64_t a_lo = desc_advance(a_hi, L::kOffALo);
          const uint64_t w_lo = desc_advance(a_hi, L::kOffWLo);
          const bool res = ks >= a_slabs;
          if (L::kHasLo && res) {
            // residual: acc[:, 32 j .. 32 j + 32) += R_hi[j] . I32 + R_lo[j] . I32 -- N = 32 instructions, an eighth of
            // the tensor work of a full-width identity slab, and the identity never travels again
            const uint32_t acc_r = d_tmem + 32 * (ks - a_slabs);
            mma(true, acc_r, a_hi, d_i32, idesc_r);
            mma(true, acc_r, desc_advance(a_hi, 32), desc_advance(d_i32, 32), idesc_r);
            mma(true, acc_r, a_lo, d_i32, idesc_r);
            mma(true, acc_r, desc_advance(a_lo, 32), desc_advance(d_i32, 32), idesc_r);
          } else {
            // 16 bf16 = 32 bytes along K inside the 64-byte swizzled row per k16 step
            mma(ks != 0, d_tmem, a_hi, w_hi, idesc);
            mma(true, d_tmem, desc_advance(a_hi, 32), desc_advance(w_hi, 32), idesc);
#ifdef LFS2_DIAG_NO_LO_MMAS
            if (false) {
#else
            if (NPASS == 3) {
#endif
              mma(true, d_tmem, a_lo, w_hi, idesc);
              mma(true, d_tmem, desc_advance(a_lo, 32), desc_advance(w_hi, 32), idesc);
            }
#ifdef LFS2_DIAG_NO_LO_MMAS
            if (false) {
#else
            if (NPASS >= 2) {
#endif
              mma(true, d_tmem, a_hi, w_lo, idesc);
              mma(true, d_tmem, desc_advance(a_hi, 32), desc_advance(w_lo, 32), idesc);
            }
          }
          if (MC) {  // pair: the slot is free / the accumulator complete in BOTH CTAs
            umma_commit_pair(&empty_bar[stage], (uint16_t)3);
            if (ks + 1 == k_slabs) umma_commit_pair(&tmem_full[acc], (uint16_t)3);
          } else {
            umma_commit(&empty_bar[stage]);                        // smem slot reusable once these MMAs retire
            if (ks + 1 == k_slabs) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (N_TILE == 256 && warp >= 2) {
    // ===================== epilogue of the 256-column tiles (every LayerNorm build): warps 2..17 =====================
    // SIXTEEN warps, four per TMEM lane quadrant (quad = warp & 3): group g = (warp - 2) >> 2 owns the tile's
    // 32-column chunks [2g, 2g + 2).  The epilogue is what paces these builds (knock-out builds, tools/gemm_ab.py:
    // releasing the accumulator unread takes 23 % off the QKV GEMM, 30 % off the out-projection and 60 % off the
    // predictor layers) and its chains -- tmem load -> normalise -> stage -> barrier -> store -- hide poorly behind
    // two warps per scheduler.  Each group has ONE 16 KB staging buffer, its own store-issuing thread and named
    // barrier; LayerNorm row statistics meet in shared memory.
    static_assert(!LN || N_TILE == 256, "LayerNorm epilogue: one 256-column tile");
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;
    const int bar_id = 1 + grp;       // 128 threads
    constexpr int kAllBar = 5;        // all 512 epilogue threads
    constexpr int kCPG = kChunks / 4;
    const int c_begin = grp * kCPG, c_end = c_begin + kCPG;
    auto release_acc = [&](uint64_t* bar) {
      if (MC) mbar_arrive_cluster(bar, 0);
      else mbar_arrive(bar);
    };
    const float* vec = reinterpret_cast<const float*>(smem + L::kOffVec);
    float2* stats = reinterpret_cast<float2*>(smem + L::kOffStats);
    uint8_t* sbuf = smem + L::kOffStaging + grp * kStageChunk;
    const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0;
    int it = 0;
    for (int tile = walk.first; tile < walk.count; tile += walk.stride, ++it) {
      int b, t0, n0;
      walk.coords(p, tile, N_TILE, b, t0, n0);
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
#ifdef LFS2_DIAG_NO_EPILOGUE
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty[acc]);
      continue;
#endif
      const uint32_t taddr = tmem_base + acc * kAccCols + ((uint32_t)(quad * 32) << 16);
      float v[32];
      const bool row_masked = p.row_mask && b < p.batch && t0 + r < p.t && p.row_mask[(size_t)b * p.t + t0 + r] != 0;

      // ---- row statistics: every group sums its chunks, the four partial sums meet in shared memory ----
      float s = 0.f, q = 0.f;
#ifdef LFS2_DIAG_LN_NO_PASS1  // timing diagnostics only: no statistics pass over tensor memory
      if (false)
#endif
      if (LN)
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        tmem_ld32(taddr + c * 32, v);
        const float4* b4 = reinterpret_cast<const float4*>(vec + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          const float x0 = activate(v[4 * j] + bb.x, p.relu, p.slope), x1 = activate(v[4 * j + 1] + bb.y, p.relu, p.slope);
          const float x2 = activate(v[4 * j + 2] + bb.z, p.relu, p.slope), x3 = activate(v[4 * j + 3] + bb.w, p.relu, p.slope);
          s += x0; q = fmaf(x0, x0, q);
          s += x1; q = fmaf(x1, x1, q);
          s += x2; q = fmaf(x2, x2, q);
          s += x3; q = fmaf(x3, x3, q);
        }
      }
      float2* st = stats + (it & 1) * 4 * kBM;
      float mean = 0.f, rstd = 1.f;
      if (LN) {
        st[grp * kBM + r] = make_float2(s, q);
        named_bar_sync(kAllBar, 512);
        s = 0.f;
        q = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // the same order in every group: identical statistics in all of them
          const float2 o = st[g * kBM + r];
          s += o.x;
          q += o.y;
        }
        mean = s * (1.f / N_TILE);
        rstd = rsqrtf(fmaxf(q * (1.f / N_TILE) - mean * mean, 0.f) + p.eps);
      }

      if (EPI == kEpiDot) {
        // ---- predictor head: out[row] = LayerNorm(z)[row] . dot_w + dot_b, masked positions 0; nothing else is stored ----
        float acc_dot = 0.f;
        const float* stw_dot = reinterpret_cast<const float*>(smem + L::kOffEdge);
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          tmem_ld32(taddr + c * 32, v);
          ln_chunk(v, vec, N_TILE, c * 32, mean, rstd, p.relu, p.slope);
          const float4* w4 = reinterpret_cast<const float4*>(stw_dot + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = w4[j];
            acc_dot = fmaf(v[4 * j], w.x, acc_dot);
            acc_dot = fmaf(v[4 * j + 1], w.y, acc_dot);
            acc_dot = fmaf(v[4 * j + 2], w.z, acc_dot);
            acc_dot = fmaf(v[4 * j + 3], w.w, acc_dot);
          }
        }
        named_bar_sync(kAllBar, 512);                 // every group has consumed the statistics of this tile
        st[grp * kBM + r].x = acc_dot;
        named_bar_sync(kAllBar, 512);
        if (grp == 0) {
          const int trow = t0 + r;
          if (b < p.batch && trow >= 0 && trow < p.t) {
            const size_t o = (size_t)b * p.t + trow;
            const float total = ((st[r].x + st[kBM + r].x) + (st[2 * kBM + r].x + st[3 * kBM + r].x)) + __ldg(p.dot_b);
            p.dot_out[o] = (p.dot_mask && p.dot_mask[o]) ? 0.f : total;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&tmem_empty[acc]);
        continue;
      }

      if (EPI == kEpiStencil) {
        // ---- u = depthwise3(LayerNorm(z)) of the next layer; rows outside the utterance count as zeros ----
        // The neighbour rows z[r-1], z[r+1] of a thread's 32 columns sit in the adjacent lanes of its warp: 64 shuffles per
        // chunk; only a warp's first and last row travel through shared memory (8 KB of edge rows, double-buffered by
        // chunk parity), so one 128-thread barrier per chunk covers the exchange AND "the staging buffer is free again".
        // (The first version staged all of z in shared memory and read the neighbours back: two more barriers per chunk,
        // a third of the kernel's time in tools/gemm_ab.py.)  u leaves as hi/lo planes through the staging buffer by TMA.
        const float* stw = reinterpret_cast<const float*>(smem + L::kOffEdge);  // w0 | w1 | w2 | bias, N_TILE floats each
        float* edge = reinterpret_cast<float*>(smem + L::kOffEdge + 4 * N_TILE * 4);
        const int trow = t0 + r;
        const bool live = b < p.batch && trow >= 0 && trow < p.t;
        const bool outrow = r >= 1 && r <= kBM - 2;   // rows 1..126 are this tile's outputs, staged as rows 0..125
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          tmem_ld32(taddr + c * 32, v);
          ln_chunk(v, vec, N_TILE, c * 32, mean, rstd, p.relu, p.slope);
          if (!live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          float* eg = edge + ((grp * 2 + (c & 1)) * 4) * 64;   // this group's edge rows of this chunk parity: [quad][2][32]
          if (lane == 0 || lane == 31) {
            float4* e4 = reinterpret_cast<float4*>(eg + quad * 64 + (lane == 0 ? 0 : 32));
#pragma unroll
            for (int i = 0; i < 8; ++i) e4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (issuer) tma_store_wait_read0();     // the previous store of this group has finished reading the staging buffer
          named_bar_sync(bar_id, 128);            // (A) edge rows visible, staging buffer free
          uint32_t hi[16], lo[16];
          {
            const float4* eup = reinterpret_cast<const float4*>(eg + (quad > 0 ? quad - 1 : 0) * 64 + 32);  // last row of the warp above
            const float4* edn = reinterpret_cast<const float4*>(eg + (quad < 3 ? quad + 1 : 3) * 64);       // first row of the warp below
            const float4* w0 = reinterpret_cast<const float4*>(stw + c * 32);
            const float4* w1 = reinterpret_cast<const float4*>(stw + N_TILE + c * 32);
            const float4* w2 = reinterpret_cast<const float4*>(stw + 2 * N_TILE + c * 32);
            const float4* bb = reinterpret_cast<const float4*>(stw + 3 * N_TILE + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 up, dn;
              up.x = __shfl_up_sync(0xffffffffu, v[4 * i], 1);
              up.y = __shfl_up_sync(0xffffffffu, v[4 * i + 1], 1);
              up.z = __shfl_up_sync(0xffffffffu, v[4 * i + 2], 1);
              up.w = __shfl_up_sync(0xffffffffu, v[4 * i + 3], 1);
              dn.x = __shfl_down_sync(0xffffffffu, v[4 * i], 1);
              dn.y = __shfl_down_sync(0xffffffffu, v[4 * i + 1], 1);
              dn.z = __shfl_down_sync(0xffffffffu, v[4 * i + 2], 1);
              dn.w = __shfl_down_sync(0xffffffffu, v[4 * i + 3], 1);
              if (lane == 0) up = eup[i];
              if (lane == 31) dn = edn[i];
              const float4 a0 = w0[i], a1 = w1[i], a2 = w2[i], ab = bb[i];
              // same association as dwconv1d_k_kernel: bias, then taps 0, 1, 2
              const float u0 = fmaf(a2.x, dn.x, fmaf(a1.x, v[4 * i], fmaf(a0.x, up.x, ab.x)));
              const float u1 = fmaf(a2.y, dn.y, fmaf(a1.y, v[4 * i + 1], fmaf(a0.y, up.y, ab.y)));
              const float u2 = fmaf(a2.z, dn.z, fmaf(a1.z, v[4 * i + 2], fmaf(a0.z, up.z, ab.z)));
              const float u3 = fmaf(a2.w, dn.w, fmaf(a1.w, v[4 * i + 3], fmaf(a0.w, up.w, ab.w)));
              split_pack2(u0, u1, hi[2 * i], lo[2 * i]);
              split_pack2(u2, u3, hi[2 * i + 1], lo[2 * i + 1]);
            }
          }
          if (outrow) {
            const int rr = r - 1;
            uint8_t* rh = sbuf + rr * 64;
            uint8_t* rl = rh + kStageChunk / 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int o = (i ^ ((rr >> 1) & 3)) << 4;
              *reinterpret_cast<uint4*>(rh + o) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              *reinterpret_cast<uint4*>(rl + o) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);            // (C) the u chunk is staged
#ifndef LFS2_DIAG_NO_STORES
          if (issuer) {  // maps with 126-row boxes; TMA clips the rows past the utterance's end
            tma_store_3d(&map_o0, sbuf, n0 + c * 32, t0 + 1, b);
            tma_store_3d(&map_o1, sbuf + kStageChunk / 2, n0 + c * 32, t0 + 1, b);
            tma_store_commit();
          }
#endif
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&tmem_empty[acc]);
        continue;
      }

#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.n) break;  // chunk entirely outside the tensor (n not a multiple of N_TILE); uniform per group
        tmem_ld32(taddr + c * 32, v);
        if (LN) {
          ln_chunk(v, vec, N_TILE, c * 32, mean, rstd, p.relu, p.slope);
        } else if (bias_vec && col0 + 32 <= p.n) {
          // the chunk's 32 bias values as 8 broadcast 16-byte loads
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = activate(v[j], p.relu, p.slope);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int col = col0 + j;
            v[j] = activate(v[j] + ((p.bias && col < p.n) ? __ldg(p.bias + col) : 0.f), p.relu, p.slope);
          }
        }
        if (row_masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (issuer) tma_store_wait_read0();       // the previous store of this group has finished reading the buffer
        named_bar_sync(bar_id, 128);
        if (OUT == kOutF32) {  // 128 rows x 128 B, SWIZZLE_128B: 16-byte unit i of row r lives at unit i ^ (r & 7)
          uint8_t* row = sbuf + r * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(row + ((i ^ (r & 7)) << 4)) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (OUT == kOutF16 || OUT == kOutBF16) {  // one 128 rows x 64 B 16-bit plane, SWIZZLE_64B
          uint32_t h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (OUT == kOutF16) h[j] = pack_f16_sat(v[2 * j], v[2 * j + 1]);
            else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
          }
          uint8_t* rh = sbuf + r * 64;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rh + ((i ^ ((r >> 1) & 3)) << 4)) = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        } else {        // two 128 rows x 64 B planes, SWIZZLE_64B: unit i of row r at i ^ ((r >> 1) & 3)
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_pack2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
          uint8_t* rh = sbuf + r * 64;
          uint8_t* rl = rh + kStageChunk / 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int u = (i ^ ((r >> 1) & 3)) << 4;
            *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
#ifndef LFS2_DIAG_NO_STORES
        if (issuer) {
          tma_store_3d(&map_o0, sbuf, col0, t0, b);
          if (OUT == kOutPlanes) tma_store_3d(&map_o1, sbuf + kStageChunk / 2, col0, t0, b);
          tma_store_commit();
        }
#endif
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty[acc]);
    }
    if (issuer) tma_store_wait_all();
  } else if (warp >= 2) {
    // ===================== epilogue of the 128- / 64-column tiles: warps 2..9 =====================
    // thread = one output row (TMEM lane quadrant = warp & 3); the two warps of a quadrant split
    // the tile's 32-column chunks between them (half 0: first chunks, half 1: the rest), each
    // half with two staging buffers, its own store-issuing thread and named barrier.  (No LayerNorm here: that
    // epilogue needs the whole 256-column row in one tile.)
    static_assert(N_TILE == 256 || (!LN && EPI == kEpiNone), "LayerNorm / fused predictor epilogues: 256-column tiles");
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const bool issuer = (warp == 2 || warp == 6) && lane == 0;  // issues / retires this half's TMA stores
    auto release_acc = [&](uint64_t* bar) {  // accumulator drained (pair: the leader's MMA warp waits for both CTAs)
      if (MC) mbar_arrive_cluster(bar, 0);
      else mbar_arrive(bar);
    };
    uint8_t* staging = smem + L::kOffStaging + half * 2 * kStageChunk;
    constexpr int kHalfChunks = kChunks / 2;
    const int c_begin = half * kHalfChunks, c_end = c_begin + kHalfChunks;
    const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0;
    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int tile = walk.first; tile < walk.count; tile += walk.stride, ++it) {
      int b, t0, n0;
      walk.coords(p, tile, N_TILE, b, t0, n0);
      int acc = it & 1;
      uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#ifdef LFS2_DIAG_NO_EPILOGUE  // timing diagnostics only (tools/gemm_ab.py): the accumulator is released unread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty[acc]);
      continue;
#endif
      uint32_t taddr = tmem_base + acc * kAccCols + ((uint32_t)(quad * 32) << 16);
      float v[32];
      const bool row_masked = p.row_mask && b < p.batch && t0 + r < p.t && p.row_mask[(size_t)b * p.t + t0 + r] != 0;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.n) break;  // chunk entirely outside the tensor (n not a multiple of N_TILE)
        tmem_ld32(taddr + c * 32, v);
        if (bias_vec && col0 + 32 <= p.n) {
          // the chunk's 32 bias values as 8 broadcast 16-byte loads (a per-element load + bounds test was 36 % of
          // this kernel's instructions)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = activate(v[j], p.relu, p.slope);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int col = col0 + j;
            v[j] = activate(v[j] + ((p.bias && col < p.n) ? __ldg(p.bias + col) : 0.f), p.relu, p.slope);
          }
        }
        if (row_masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        uint8_t* sb = staging + (chunk_ctr & 1) * kStageChunk;
        ++chunk_ctr;
        if (OUT == kOutF32) {  // 128 rows x 128 B, SWIZZLE_128B: 16-byte unit i of row r lives at unit i ^ (r & 7)
          uint8_t* row = sb + r * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(row + ((i ^ (r & 7)) << 4)) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (OUT == kOutF16 || OUT == kOutBF16) {  // one 128 rows x 64 B 16-bit plane, SWIZZLE_64B
          uint32_t h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (OUT == kOutF16) h[j] = pack_f16_sat(v[2 * j], v[2 * j + 1]);
            else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
          }
          uint8_t* rh = sb + r * 64;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rh + ((i ^ ((r >> 1) & 3)) << 4)) = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        } else {        // two 128 rows x 64 B planes, SWIZZLE_64B: unit i of row r at i ^ ((r >> 1) & 3)
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_pack2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
          uint8_t* rh = sb + r * 64;
          uint8_t* rl = rh + kStageChunk / 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int u = (i ^ ((r >> 1) & 3)) << 4;
            *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        // the store of the previous chunk (other buffer) must have finished READING before the
        // threads that pass this barrier start overwriting that buffer for the next chunk
        // (letting it stay in flight one chunk longer -- wait_group.read 1 + a second barrier -- measured no gain)
        if (issuer) tma_store_wait_read0();
        named_bar_sync(1 + half, 128);
#ifndef LFS2_DIAG_NO_STORES
        if (issuer) {
          tma_store_3d(&map_o0, sb, col0, t0, b);
          if (OUT == kOutPlanes) tma_store_3d(&map_o1, sb + kStageChunk / 2, col0, t0, b);
          tma_store_commit();
        }
#endif
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty[acc]);
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer may still arrive on this CTA's barriers until it has finished too
  if (warp == 1) {
    tc_fence_after();
    if (MC) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
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

// cuTensorMapEncodeTiled costs ~1-2 us of host time and a launch needs up to eleven maps; a step re-issues the same
// (pointer, shape, box) combinations over and over (weights are fixed, activations come back from the caching allocator
// at the same addresses), so encoded maps are memoised.  A map is a pure function of its key: a stale entry cannot exist.
struct TmapKey {
  const void* base;
  uint64_t d0, d1, d2;
  uint32_t box0, box1;
  int elem_bytes, swizzle_bytes;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ reinterpret_cast<uintptr_t>(k.base);
    const uint64_t v[5] = {k.d0, k.d1, k.d2, ((uint64_t)k.box0 << 32) | k.box1,
                           ((uint64_t)(uint32_t)k.elem_bytes << 32) | (uint32_t)k.swizzle_bytes};
    for (uint64_t x : v) h = (h ^ x) * 0xBF58476D1CE4E5B9ull + (h >> 29);
    return (size_t)h;
  }
};

bool make_tmap_3d_ex(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                     uint32_t box0, uint32_t box1, int swizzle_bytes) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key;
  memset(&key, 0, sizeof(key));  // padding bytes take part in the comparison
  key.base = base; key.d0 = d0; key.d1 = d1; key.d2 = d2; key.box0 = box0; key.box1 = box1;
  key.elem_bytes = elem_bytes; key.swizzle_bytes = swizzle_bytes;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * elem_bytes, d0 * d1 * elem_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 0  ? CU_TENSOR_MAP_SWIZZLE_NONE   // dense rows (dwconv_tma.cu)
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() >= 16384) cache.clear();
  cache.emplace(key, *out);
  return true;
}

bool make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0,
                  uint32_t box1, int swizzle_bytes) {
  return make_tmap_3d_ex(out, base, 2, d0, d1, d2, box0, box1, swizzle_bytes);
}

struct GemmTcMaps {
  CUtensorMap ah, al, wh, wl, rh, rl, ident, o0, o1;
};

template <int N_TILE, int NPASS, bool LN, int OUT, bool MC, int EPI = 0>
static int launch_gemm_tc(const GemmTcMaps& m, const GemmTcParams& p, cudaStream_t s) {
  using L = SmemLayout<N_TILE, NPASS, LN, EPI, MC>;
  auto kern = gemm_tc_kernel<N_TILE, NPASS, LN, OUT, MC, EPI>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) {
      set_error("gemm_tc: cannot reserve %d bytes of shared memory", L::kTotal);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (!MC) {
    kern<<<grid, gemm_tc_threads(N_TILE), L::kTotal, s>>>(m.ah, m.al, m.wh, m.wl, m.rh, m.rl, m.ident, m.o0, m.o1, p);
  } else {  // clusters of two CTAs (one per SM): pairs of row tiles share the multicast weight slabs
    grid &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(gemm_tc_threads(N_TILE));
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, m.ah, m.al, m.wh, m.wl, m.rh, m.rl, m.ident, m.o0, m.o1, p) != cudaSuccess) {
      set_error("gemm_tc: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      return LFS2_ERR_CUDA;
    }
  }
  LFS2_CHECK_LAUNCH("gemm_tc");
  return LFS2_OK;
}

template <int N_TILE, bool LN, bool MC>
static int dispatch_gemm_tc(const GemmTcMaps& m, const GemmTcParams& p, int npass, int out_kind, cudaStream_t s) {
  if (npass == 3) {
    if (out_kind == kOutF32) return launch_gemm_tc<N_TILE, 3, LN, kOutF32, MC>(m, p, s);
    if (out_kind == kOutF16) return launch_gemm_tc<N_TILE, 3, LN, kOutF16, MC>(m, p, s);
    if (out_kind == kOutBF16) return launch_gemm_tc<N_TILE, 3, LN, kOutBF16, MC>(m, p, s);
    return launch_gemm_tc<N_TILE, 3, LN, kOutPlanes, MC>(m, p, s);
  }
  if (npass == 2) {
    if constexpr (!LN) {
      if (out_kind == kOutF32) return launch_gemm_tc<N_TILE, 2, false, kOutF32, MC>(m, p, s);
      if (out_kind == kOutF16) return launch_gemm_tc<N_TILE, 2, false, kOutF16, MC>(m, p, s);
      if (out_kind == kOutBF16) return launch_gemm_tc<N_TILE, 2, false, kOutBF16, MC>(m, p, s);
      return launch_gemm_tc<N_TILE, 2, false, kOutPlanes, MC>(m, p, s);
    } else {
      set_error("gemm_tc: npass = 2 has no LayerNorm build");
      return LFS2_ERR_UNSUPPORTED;
    }
  }
  if (out_kind == kOutF32) return launch_gemm_tc<N_TILE, 1, LN, kOutF32, MC>(m, p, s);
  if (out_kind == kOutF16) return launch_gemm_tc<N_TILE, 1, LN, kOutF16, MC>(m, p, s);
  if (out_kind == kOutBF16) return launch_gemm_tc<N_TILE, 1, LN, kOutBF16, MC>(m, p, s);
  return launch_gemm_tc<N_TILE, 1, LN, kOutPlanes, MC>(m, p, s);
}

// LFS2_GEMM_MULTICAST=0 switches the CTA-pair variant off (A/B measurements)
static bool gemm_multicast_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LFS2_GEMM_MULTICAST");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// dx = y > 0 ? dy * scale : 0 written as bf16 hi/lo planes (what the input- and weight-gradient GEMMs read), and
// db += column sums of dx: relu_bwd + split_bf16 + colsum of the train step in one pass (10 instead of 24 bytes per
// element).  y is the saved ReLU output as fp32 or as the hi plane of its bf16 split (sign and zero survive the split).
// CTA = 32 float4 column groups x 8 row lanes over a row range.
template <bool Y_F32>
__global__ void __launch_bounds__(256)
relu_bwd_planes_kernel(const float4* __restrict__ dy, const void* __restrict__ y, uint2* __restrict__ hi,
                       uint2* __restrict__ lo, float* __restrict__ db, int m, int n4, int rows_per_cta, float scale) {
  __shared__ float4 part[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int cg = blockIdx.x * 32 + cl;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < n4) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 8) {
      const size_t i = (size_t)r * n4 + cg;
      float4 g = dy[i];
      bool p0, p1, p2, p3;
      if (Y_F32) {
        const float4 v = reinterpret_cast<const float4*>(y)[i];
        p0 = v.x > 0.f; p1 = v.y > 0.f; p2 = v.z > 0.f; p3 = v.w > 0.f;
      } else {  // bf16 pairs, first element in the low half: positive <=> sign clear and not zero
        const uint2 v = reinterpret_cast<const uint2*>(y)[i];
        p0 = (v.x & 0x8000u) == 0u && (v.x & 0x7fffu) != 0u;
        p1 = (v.x & 0x80000000u) == 0u && (v.x & 0x7fff0000u) != 0u;
        p2 = (v.y & 0x8000u) == 0u && (v.y & 0x7fffu) != 0u;
        p3 = (v.y & 0x80000000u) == 0u && (v.y & 0x7fff0000u) != 0u;
      }
      g.x = p0 ? g.x * scale : 0.f;
      g.y = p1 ? g.y * scale : 0.f;
      g.z = p2 ? g.z * scale : 0.f;
      g.w = p3 ? g.w * scale : 0.f;
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      uint2 h, l;
      split_pack2(g.x, g.y, h.x, l.x);
      split_pack2(g.z, g.w, h.y, l.y);
      hi[i] = h;
      lo[i] = l;
    }
  }
  if (db == nullptr) return;  // kernel argument: uniform
  part[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && cg < n4) {
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float4 v = part[q][cl];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    atomicAdd(db + cg * 4 + 0, s.x);
    atomicAdd(db + cg * 4 + 1, s.y);
    atomicAdd(db + cg * 4 + 2, s.z);
    atomicAdd(db + cg * 4 + 3, s.w);
  }
}

// One launch for the per-step re-formatting of every GEMM weight of the train step: bf16 hi/lo planes of W (rows, cols)
// and of W^T (cols, rows) for a table of matrices (lfs2_prep_entry, include/lfs2.h).  CTA = one 32 x 32 tile of one
// entry, found by binary search over the entries' first-tile indices.  Same values as split_bf16(W) and
// split_bf16(transpose(W)), which it replaces (~210 launches of a few microseconds each per step).
__global__ void __launch_bounds__(256)
weight_prep_kernel(const lfs2_prep_entry* __restrict__ entries, int n) {
  __shared__ float tile[32][33];
  const int tid = blockIdx.x;
  int lo_i = 0, hi_i = n - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (entries[mid].tile_begin <= tid) lo_i = mid;
    else hi_i = mid - 1;
  }
  const lfs2_prep_entry en = entries[lo_i];
  const int tiles_x = (en.cols + 31) >> 5;
  const int lt = tid - en.tile_begin;
  const int r0 = (lt / tiles_x) << 5, c0 = (lt % tiles_x) << 5;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(en.hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(en.lo);
  __nv_bfloat16* hi_t = reinterpret_cast<__nv_bfloat16*>(en.hi_t);
  __nv_bfloat16* lo_t = reinterpret_cast<__nv_bfloat16*>(en.lo_t);
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < en.rows && c < en.cols) {
      v = en.src[(size_t)r * en.cols + c];
      if (hi) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        hi[(size_t)r * en.cols + c] = h;
        lo[(size_t)r * en.cols + c] = l;
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (hi_t) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < en.rows && c < en.cols) {
        __nv_bfloat16 h, l;
        split_bf16(tile[threadIdx.x][i], h, l);
        hi_t[(size_t)c * en.rows + r] = h;
        lo_t[(size_t)c * en.rows + r] = l;
      }
    }
  }
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" {

long long lfs2_gemm_tc_limited_workspace_bytes(int batch, int t) {
  return batch > 0 && t > 0 ? (1 + (long long)batch * ((t + kBM - 1) / kBM)) * (long long)sizeof(int) : 0;
}

int lfs2_gemm_tc(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps, const void* w_hi,
                 const void* w_lo, int n, const float* bias, int relu, const void* res_hi, const void* res_lo,
                 const void* ident_hi, const float* gamma, const float* beta, float eps, float* out_f32,
                 void* out_hi, void* out_lo, int npass, void* stream) {
  return lfs2_gemm_tc_limited(a_hi, a_lo, batch, t, d, taps, w_hi, w_lo, n, bias, relu, res_hi, res_lo, ident_hi, gamma,
                              beta, eps, out_f32, out_hi, out_lo, npass, nullptr, 0, nullptr, stream);
}

int lfs2_gemm_tc_limited(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps, const void* w_hi,
                         const void* w_lo, int n, const float* bias, int relu, const void* res_hi, const void* res_lo,
                         const void* ident_hi, const float* gamma, const float* beta, float eps, float* out_f32,
                         void* out_hi, void* out_lo, int npass, const int* row_limit, int limit_extra,
                         void* workspace, void* stream) {
  LFS2_REQUIRE((out_f32 != nullptr) != (out_hi != nullptr), LFS2_ERR_INVALID_ARG,
               "gemm_tc: exactly one of out_f32 / out_hi+out_lo");
  LFS2_REQUIRE(!out_hi || out_lo, LFS2_ERR_INVALID_ARG, "gemm_tc: out_hi without out_lo");
  return lfs2_gemm_tc_ex(a_hi, a_lo, batch, t, d, taps, 1, w_hi, w_lo, n, bias, relu ? 1 : 0, 0.f, res_hi, res_lo,
                         ident_hi, gamma, beta, eps, out_f32 ? (void*)out_f32 : out_hi, out_lo,
                         out_f32 ? LFS2_OUT_F32 : LFS2_OUT_PLANES, npass, row_limit, limit_extra, workspace, nullptr,
                         stream);
}

int lfs2_gemm_tc_ex(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps, int dilation,
                    const void* w_hi, const void* w_lo, int n, const float* bias, int activation, float slope,
                    const void* res_hi, const void* res_lo, const void* ident_hi, const float* gamma, const float* beta,
                    float eps, void* out0, void* out1, int out_kind, int npass, const int* row_limit, int limit_extra,
                    void* workspace, const uint8_t* row_mask, void* stream) {
  LFS2_REQUIRE(a_hi && w_hi && out0, LFS2_ERR_INVALID_ARG, "gemm_tc: null operand");
  LFS2_REQUIRE(npass >= 1 && npass <= 3, LFS2_ERR_INVALID_ARG, "gemm_tc: npass must be 1, 2 or 3");
  LFS2_REQUIRE(npass == 1 || w_lo, LFS2_ERR_INVALID_ARG, "gemm_tc: npass >= 2 needs the weight lo plane");
  LFS2_REQUIRE(npass != 3 || a_lo, LFS2_ERR_INVALID_ARG, "gemm_tc: npass=3 needs the lo plane of a");
  LFS2_REQUIRE(npass != 2 || (!gamma && !res_hi), LFS2_ERR_UNSUPPORTED,
               "gemm_tc: npass = 2 (one fp16 activation plane) has no LayerNorm / residual build");
  LFS2_REQUIRE(out_kind >= LFS2_OUT_PLANES && out_kind <= LFS2_OUT_BF16, LFS2_ERR_INVALID_ARG,
               "gemm_tc: out_kind must be LFS2_OUT_PLANES, LFS2_OUT_F32, LFS2_OUT_F16 or LFS2_OUT_BF16");
  LFS2_REQUIRE(out_kind != LFS2_OUT_PLANES || out1, LFS2_ERR_INVALID_ARG, "gemm_tc: plane output needs out1 (the lo plane)");
  LFS2_REQUIRE(activation >= 0 && activation <= 2, LFS2_ERR_INVALID_ARG, "gemm_tc: activation must be 0, 1 or 2");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && n > 0 && taps > 0 && dilation > 0, LFS2_ERR_INVALID_ARG, "gemm_tc: bad shape");
  LFS2_REQUIRE(taps % 2 == 1, LFS2_ERR_UNSUPPORTED, "gemm_tc: kernel size %d must be odd", taps);
  LFS2_REQUIRE(d % kBK == 0, LFS2_ERR_UNSUPPORTED, "gemm_tc: d=%d must be a multiple of %d", d, kBK);
  LFS2_REQUIRE(n % 16 == 0, LFS2_ERR_UNSUPPORTED, "gemm_tc: n=%d must be a multiple of 16", n);
  LFS2_REQUIRE(aligned16(a_hi) && aligned16(w_hi) && (!a_lo || aligned16(a_lo)) && (!w_lo || aligned16(w_lo)) &&
                   aligned16(out0) && (!out1 || aligned16(out1)) && (!res_hi || (aligned16(res_hi) && aligned16(res_lo))),
               LFS2_ERR_INVALID_ARG, "gemm_tc: pointers must be 16-byte aligned");
  const bool ln = gamma != nullptr;
  LFS2_REQUIRE(!ln || beta, LFS2_ERR_INVALID_ARG, "gemm_tc: gamma without beta");
  LFS2_REQUIRE(!res_hi || (res_lo && ident_hi), LFS2_ERR_INVALID_ARG, "gemm_tc: the residual needs hi, lo planes and the identity");
  // (the residual slabs use the lo-plane slot of a pipeline stage, which exists in 3-pass and LayerNorm builds)
  LFS2_REQUIRE(!res_hi || ln || npass == 3, LFS2_ERR_UNSUPPORTED,
               "gemm_tc: a residual without LayerNorm needs npass = 3");
  LFS2_REQUIRE(!res_hi || !activation, LFS2_ERR_UNSUPPORTED, "gemm_tc: an activation after the residual add is not a reference pattern");
  int n_tile = ln ? n : (n % 256 == 0 ? 256 : (n > 64 ? 128 : 64));
  {  // LFS2_GEMM_NTILE=128: A/B knob (tools/gemm_ab.py) -- narrower column tiles for the plain epilogue
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("LFS2_GEMM_NTILE");
      forced = e ? atoi(e) : 0;
    }
    if (!ln && forced == 128 && n_tile == 256) n_tile = 128;
  }
  if (ln) LFS2_REQUIRE(n == 256, LFS2_ERR_UNSUPPORTED, "gemm_tc: LayerNorm epilogue needs n == 256 (got %d)", n);

  // 2-CTA multicast variant: full 256-column tiles, at least one pair of row tiles per cluster
  const int m_tiles_all = batch * ((t + kBM - 1) / kBM);
  // (not for the LayerNorm builds: they are paced by their epilogue, and coupling two CTAs' epilogues to one MMA
  // stream costs them 6-10 %; measured in profiles/r2p_*)
  const bool mc = !ln && n_tile == 256 && n % 256 == 0 && m_tiles_all >= 2 * num_sms() && gemm_multicast_enabled();
  const uint32_t w_box = mc ? n_tile / 2 : n_tile;  // MC: each CTA of a pair fetches half a weight slab

  GemmTcMaps m;
  const uint64_t ktot = (uint64_t)taps * d;
  bool ok = make_tmap_3d(&m.ah, a_hi, d, t, batch, kBK, kBM, 64) && make_tmap_3d(&m.wh, w_hi, ktot, n, 1, kBK, w_box, 64);
  m.al = m.ah;
  m.wl = m.wh;
  if (npass == 3) ok = ok && make_tmap_3d(&m.al, a_lo, d, t, batch, kBK, kBM, 64);
  if (npass >= 2) ok = ok && make_tmap_3d(&m.wl, w_lo, ktot, n, 1, kBK, w_box, 64);
  if (res_hi)
    ok = ok && make_tmap_3d(&m.rh, res_hi, n, t, batch, kBK, kBM, 64) &&
         make_tmap_3d(&m.rl, res_lo, n, t, batch, kBK, kBM, 64) && make_tmap_3d(&m.ident, ident_hi, n, n, 1, kBK, mc ? 16 : 32, 64);
  else {
    m.rh = m.ah;
    m.rl = m.ah;
    m.ident = m.wh;
  }
  if (out_kind == LFS2_OUT_F32) {
    ok = ok && make_tmap_3d_ex(&m.o0, out0, 4, n, t, batch, 32, kBM, 128);
    m.o1 = m.o0;
  } else {  // 16-bit planes (bf16 hi/lo, or one fp16 plane: the map only moves bytes)
    ok = ok && make_tmap_3d(&m.o0, out0, n, t, batch, 32, kBM, 64);
    if (out_kind == LFS2_OUT_PLANES) ok = ok && make_tmap_3d(&m.o1, out1, n, t, batch, 32, kBM, 64);
    else m.o1 = m.o0;
  }
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed (driver entry point %s)",
               get_encode_tiled() ? "found" : "missing");

  GemmTcParams p;
  p.batch = batch; p.t = t; p.d = d; p.taps = taps; p.half = (taps - 1) / 2; p.dil = dilation;
  p.n = n;
  p.m_tiles_per_batch = (t + kBM - 1) / kBM;
  p.n_tiles = (n + n_tile - 1) / n_tile;
  p.total_tiles = batch * p.m_tiles_per_batch * p.n_tiles;
  p.has_residual = res_hi != nullptr;
  p.bias = bias; p.relu = activation; p.slope = slope; p.row_mask = row_mask;
  p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.m_step = kBM; p.t_shift = 0;
  p.st_w = p.st_b = p.dot_w = p.dot_b = nullptr; p.dot_mask = nullptr; p.dot_out = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  p.tile_list = nullptr;
  if (!row_limit && workspace) {
    p.tile_list = reinterpret_cast<const int*>(workspace);  // a list built earlier for the same (batch, t, limit): reuse
  } else if (row_limit) {  // compact list of the row tiles that are needed, built on the device (no host read-back)
    LFS2_REQUIRE(workspace, LFS2_ERR_INVALID_ARG, "gemm_tc: a row limit needs the tile-list workspace");
    int* list = reinterpret_cast<int*>(workspace);
    if (cudaMemsetAsync(list, 0, sizeof(int), s) != cudaSuccess) {
      set_error("gemm_tc: memset failed");
      return LFS2_ERR_CUDA;
    }
    gemm_tile_list_kernel<<<1, 256, 0, s>>>(row_limit, limit_extra, batch, p.m_tiles_per_batch, list, kBM);
    LFS2_CHECK_LAUNCH("gemm_tile_list");
    p.tile_list = list;
  }
  if (mc) return dispatch_gemm_tc<256, false, true>(m, p, npass, out_kind, s);
  if (ln) return dispatch_gemm_tc<256, true, false>(m, p, npass, out_kind, s);
  if (n_tile == 256) return dispatch_gemm_tc<256, false, false>(m, p, npass, out_kind, s);
  if (n_tile == 128) return dispatch_gemm_tc<128, false, false>(m, p, npass, out_kind, s);
  return dispatch_gemm_tc<64, false, false>(m, p, npass, out_kind, s);
}

long long lfs2_predictor_layer_tc_workspace_bytes(int batch, int t) {
  return batch > 0 && t > 0 ? (1 + (long long)batch * ((t + 125) / 126)) * (long long)sizeof(int) : 0;
}

int lfs2_predictor_layer_tc(const void* a_hi, const void* a_lo, int batch, int t, const void* w_hi, const void* w_lo,
                            const float* bias, const float* gamma, const float* beta, float eps, int npass,
                            const float* next_dw_w, const float* next_dw_b, void* out_hi, void* out_lo,
                            const float* head_w, const float* head_b, const uint8_t* head_mask, float* head_out,
                            const int* row_limit, int limit_extra, void* workspace, void* stream) {
  constexpr int d = 256, n = 256;
  LFS2_REQUIRE(a_hi && w_hi && bias && gamma && beta, LFS2_ERR_INVALID_ARG, "predictor_layer_tc: null operand");
  LFS2_REQUIRE(npass == 1 || npass == 3, LFS2_ERR_INVALID_ARG, "predictor_layer_tc: npass must be 1 or 3");
  LFS2_REQUIRE(npass == 1 || (a_lo && w_lo), LFS2_ERR_INVALID_ARG, "predictor_layer_tc: npass=3 needs the lo planes");
  const bool stencil = next_dw_w != nullptr;
  LFS2_REQUIRE(stencil != (head_w != nullptr), LFS2_ERR_INVALID_ARG,
               "predictor_layer_tc: exactly one of the next layer's depthwise conv / the head");
  LFS2_REQUIRE(!stencil || (next_dw_b && out_hi && out_lo), LFS2_ERR_INVALID_ARG, "predictor_layer_tc: stencil outputs missing");
  LFS2_REQUIRE(stencil || (head_b && head_out), LFS2_ERR_INVALID_ARG, "predictor_layer_tc: head outputs missing");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0, LFS2_ERR_INVALID_ARG, "predictor_layer_tc: bad shape");
  LFS2_REQUIRE(aligned16(a_hi) && aligned16(w_hi) && (!a_lo || aligned16(a_lo)) && (!w_lo || aligned16(w_lo)) &&
                   (!out_hi || (aligned16(out_hi) && aligned16(out_lo))),
               LFS2_ERR_INVALID_ARG, "predictor_layer_tc: pointers must be 16-byte aligned");
  const int m_step = stencil ? kBM - 2 : kBM;
  GemmTcMaps m;
  bool ok = make_tmap_3d(&m.ah, a_hi, d, t, batch, kBK, kBM, 64) && make_tmap_3d(&m.wh, w_hi, d, n, 1, kBK, n, 64);
  if (npass == 3)
    ok = ok && make_tmap_3d(&m.al, a_lo, d, t, batch, kBK, kBM, 64) && make_tmap_3d(&m.wl, w_lo, d, n, 1, kBK, n, 64);
  else {
    m.al = m.ah;
    m.wl = m.wh;
  }
  m.rh = m.ah;
  m.rl = m.ah;
  m.ident = m.wh;
  if (stencil) {  // 126-row boxes: a tile stores its rows 1..126
    ok = ok && make_tmap_3d(&m.o0, out_hi, n, t, batch, 32, kBM - 2, 64) && make_tmap_3d(&m.o1, out_lo, n, t, batch, 32, kBM - 2, 64);
  } else {
    m.o0 = m.ah;
    m.o1 = m.ah;
  }
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "predictor_layer_tc: cuTensorMapEncodeTiled failed");
  GemmTcParams p;
  p.batch = batch; p.t = t; p.d = d; p.taps = 1; p.half = 0; p.dil = 1;
  p.n = n;
  p.m_step = m_step; p.t_shift = stencil ? -1 : 0;
  p.m_tiles_per_batch = (t + m_step - 1) / m_step;
  p.n_tiles = 1;
  p.total_tiles = batch * p.m_tiles_per_batch;
  p.has_residual = 0;
  p.bias = bias; p.relu = 1; p.slope = 0.f; p.row_mask = nullptr;
  p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.st_w = next_dw_w; p.st_b = next_dw_b; p.dot_w = head_w; p.dot_b = head_b; p.dot_mask = head_mask; p.dot_out = head_out;
  cudaStream_t s = (cudaStream_t)stream;
  p.tile_list = nullptr;
  if (!row_limit && workspace) {
    p.tile_list = reinterpret_cast<const int*>(workspace);
  } else if (row_limit) {
    LFS2_REQUIRE(workspace, LFS2_ERR_INVALID_ARG, "predictor_layer_tc: a row limit needs the tile-list workspace");
    int* list = reinterpret_cast<int*>(workspace);
    if (cudaMemsetAsync(list, 0, sizeof(int), s) != cudaSuccess) {
      set_error("predictor_layer_tc: memset failed");
      return LFS2_ERR_CUDA;
    }
    gemm_tile_list_kernel<<<1, 256, 0, s>>>(row_limit, limit_extra, batch, p.m_tiles_per_batch, list, m_step);
    LFS2_CHECK_LAUNCH("gemm_tile_list");
    p.tile_list = list;
  }
  if (stencil)
    return npass == 3 ? launch_gemm_tc<256, 3, true, kOutPlanes, false, kEpiStencil>(m, p, s)
                      : launch_gemm_tc<256, 1, true, kOutPlanes, false, kEpiStencil>(m, p, s);
  return npass == 3 ? launch_gemm_tc<256, 3, true, kOutPlanes, false, kEpiDot>(m, p, s)
                    : launch_gemm_tc<256, 1, true, kOutPlanes, false, kEpiDot>(m, p, s);
}

// x (n) fp32 -> hi/lo fp16 planes: hi = fp16(x) (saturating), lo = fp16(x - hi); hi + lo = x to ~2^-22 |x| (2^-25
// absolute once lo is subnormal) -- the weight operand of the 2-pass recipe
__device__ __forceinline__ void split_f16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16_sat(a, b);
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  lo = pack_f16_sat(a - __low2float(h), b - __high2float(h));
}
__global__ void split_f16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  uint2 h, l;
  split_f16_pair(v.x, v.y, h.x, l.x);
  split_f16_pair(v.z, v.w, h.y, l.y);
  hi[i] = h;
  lo[i] = l;
}

// x (n) fp32 -> hi/lo bf16 planes
__global__ void split_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  uint2* __restrict__ f16, size_t n4) {
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
  if (f16) f16[i] = make_uint2(pack_f16_sat(v.x, v.y), pack_f16_sat(v.z, v.w));
}

int lfs2_relu_bwd_planes(const float* dy, const float* y_f32, const void* y_hi, void* dx_hi, void* dx_lo, float* db,
                         int rows, int cols, float scale, void* stream) {
  LFS2_REQUIRE(dy && dx_hi && dx_lo && ((y_f32 != nullptr) != (y_hi != nullptr)), LFS2_ERR_INVALID_ARG,
               "relu_bwd_planes: null pointer, or not exactly one of y_f32 / y_hi given");
  if (rows == 0 || cols == 0) return LFS2_OK;
  LFS2_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, LFS2_ERR_UNSUPPORTED,
               "relu_bwd_planes: cols=%d must be a positive multiple of 4", cols);
  LFS2_REQUIRE(aligned16(dy) && aligned16(y_f32) && aligned16(y_hi) && aligned16(dx_hi) && aligned16(dx_lo),
               LFS2_ERR_INVALID_ARG, "relu_bwd_planes: pointers must be 16-byte aligned");
  const int n4 = cols / 4, rows_per_cta = 128;
  dim3 grid(ceil_div(n4, 32), ceil_div(rows, rows_per_cta));
  LFS2_REQUIRE(grid.y <= 65535u, LFS2_ERR_UNSUPPORTED, "relu_bwd_planes: rows=%d exceeds the grid limit", rows);
  if (y_f32)
    relu_bwd_planes_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)dy, y_f32, (uint2*)dx_hi,
                                                                          (uint2*)dx_lo, db, rows, n4, rows_per_cta, scale);
  else
    relu_bwd_planes_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)dy, y_hi, (uint2*)dx_hi,
                                                                           (uint2*)dx_lo, db, rows, n4, rows_per_cta, scale);
  LFS2_CHECK_LAUNCH("relu_bwd_planes");
  return LFS2_OK;
}

int lfs2_weight_planes_batched(const lfs2_prep_entry* entries, int n_entries, int total_tiles, void* stream) {
  LFS2_REQUIRE(entries, LFS2_ERR_INVALID_ARG, "weight_planes_batched: null table");
  if (n_entries == 0 || total_tiles == 0) return LFS2_OK;
  LFS2_REQUIRE(n_entries > 0 && total_tiles > 0, LFS2_ERR_INVALID_ARG, "weight_planes_batched: bad counts");
  weight_prep_kernel<<<total_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(entries, n_entries);
  LFS2_CHECK_LAUNCH("weight_planes_batched");
  return LFS2_OK;
}

int lfs2_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream) {
  return lfs2_split_bf16_ex(x, hi, lo, nullptr, n, stream);
}

int lfs2_split_f16(const float* x, void* hi, void* lo, long long n, void* stream) {
  LFS2_REQUIRE(x && hi && lo, LFS2_ERR_INVALID_ARG, "split_f16: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "split_f16: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(x) && aligned16(hi) && aligned16(lo), LFS2_ERR_INVALID_ARG,
               "split_f16: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  split_f16_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n4);
  LFS2_CHECK_LAUNCH("split_f16");
  return LFS2_OK;
}

int lfs2_split_bf16_ex(const float* x, void* hi, void* lo, void* f16, long long n, void* stream) {
  LFS2_REQUIRE(x && hi && lo && aligned16(f16), LFS2_ERR_INVALID_ARG, "split_bf16: null or misaligned pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "split_bf16: n must be a positive multiple of 4");
  LFS2_REQUIRE(aligned16(x) && aligned16(hi) && aligned16(lo), LFS2_ERR_INVALID_ARG,
               "split_bf16: pointers must be 16-byte aligned");
  size_t n4 = (size_t)n / 4;
  split_bf16_kernel<<<ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo,
                                                                         (uint2*)f16, n4);
  LFS2_CHECK_LAUNCH("split_bf16");
  return LFS2_OK;
}

}  // extern "C"
