// Fused position-wise FFN of the depthwise FFTBlock (reference model.py:118-122 after the depthwise conv):
//
//     out = LayerNorm( x1 + relu(u . W11^T + b11) . W_eff^T + b_eff )          u = dwconv(x1), d = 256
//
// in ONE kernel: the F-wide intermediate v = relu(.) never leaves the SM.  Per 128-row tile and per
// 128-column chunk c of F (two acc1 buffers, so the conversion of chunk c runs under the MMAs of chunk c+1):
//     G1  acc1[c&1] (TMEM, fp32 128 x 128)  = u . W11[c]^T            SS form: u and W11 slabs by TMA
//     E1  8 epilogue warps: acc1 + b11, ReLU, split into bf16 hi/lo pairs, written back IN PLACE into the same
//         tensor-memory columns (per 32-column block: 16 packed hi columns, 16 packed lo columns)
//     G2  acc2 (TMEM, fp32 128 x 256) += v[c] . W_eff[:, c]^T        TS form: A = v read from tensor memory
// tensor-pipe order: G1(0) | G1(1) G2(0) | G1(2) G2(1) | ...
// then the residual x1 rides the tensor core: per 32-column slab, acc2[:, slab] += R_hi . I32 + R_lo . I32 with ONE
// 32 x 32 identity block held in shared memory (N = 32 instructions: 1/8 of the work of a full-width identity slab
// and no identity traffic), and the epilogue does + b_eff, LayerNorm over the 256 columns and leaves as bf16 hi/lo
// planes (plus, on request, the same rows as ONE fp16 plane for the next block's 2-pass QKV GEMM) through swizzled
// staging + TMA stores.
// HBM traffic per row: read u (1 KB) + x1 (1 KB), write out (1 KB) -- the unfused pair moves 11 KB.
// tcgen05.mma executes in issue order, so G1 of chunk c+1 may be issued right behind G2 of chunk c although it
// overwrites the columns G2 reads (same pattern as S/P in attention_tc.cu).
// NPASS = 3: hi.hi + lo.hi + hi.lo for every product; NPASS = 1: hi.hi only (bf16 mode);
// NPASS = 2: the activation operand is ONE fp16 value (u arrives as an fp16 plane, v is packed as fp16 in tensor
// memory) against fp16 hi/lo weight planes (lfs2_split_f16; one instruction cannot mix an fp16 A with a bf16 B):
// a.w_hi + a.w_lo -- 11 significant bits on the activation side, full weights; two thirds of the tensor work
// (tools/precision_emulation.py has the mel error of this recipe per site).
// MC: clusters of two CTAs (the kernel is bound by the L2 -> SM stream of the weights, ~2 MB per 128-row tile against
// ~0.6 MB of activations): the pair works on two row tiles at once, each CTA fetches HALF of every weight slab and
// multicasts it into both CTAs' shared memory -- half the weight traffic per row (same scheme as gemm_tc.cu).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..9 = epilogue (thread = row, two warps per quadrant).
// (Sixteen epilogue warps -- which help the plain LayerNorm GEMMs of gemm_tc.cu -- made this kernel 3.5 % slower: the
// register cap of a 576-thread block is 96, both epilogues spill, and the fp16 plane needs two more barriers per chunk
// once every group has a single staging buffer.)
#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kFM = 128;      // rows per tile
constexpr int kFK = 32;       // k-slab
constexpr int kFD = 256;      // model width = K of G1 = N of G2 = LayerNorm width
constexpr int kFC = 128;      // F chunk = N of G1 = K of G2
constexpr int kFThreads = 320;
constexpr int kFStageChunk = kFM * 32 * 4;  // 16 KB staging chunk (hi | lo planes of 128 x 32)

struct FfnParams {
  int m_pad;        // first row past every tile (an out-of-range tile of an odd pair starts here: TMA fills zeros / drops)
  int m;            // rows
  int f;            // hidden width (multiple of 256)
  int total_tiles;
  const int* tile_list;  // null, or [0] = number of active 128-row tiles, [1..] = their indices (row-limited launch)
  const float* b1;  // (f)
  const float* b2;  // (256) folded bias
  const float* gamma;
  const float* beta;
  float eps;
  int out_f16;      // also store the rows as one fp16 plane (map_o_f16)
};

// work item i of this CTA's stride loop -> 128-row tile
__device__ __forceinline__ int tile_count(const FfnParams& p) { return p.tile_list ? __ldg(p.tile_list) : p.total_tiles; }
__device__ __forceinline__ int tile_at(const FfnParams& p, int i) { return p.tile_list ? __ldg(p.tile_list + 1 + i) : i; }

// active tiles of a row-limited launch over (batch, t) rows flattened to (batch * t): a 128-row tile is needed when it
// holds a row of some utterance b before that utterance's last kept 128-row group, i.e. t_row < roundup128(limit[b] + extra)
__global__ void ffn_tile_list_kernel(const int* __restrict__ row_limit, int extra, int batch, int t, int total_tiles,
                                     int* __restrict__ list) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_tiles; i += gridDim.x * blockDim.x) {
    const long long r0 = (long long)i * kFM, r1 = min(r0 + kFM, (long long)batch * t);
    bool need = false;
    for (int b = (int)(r0 / t); b < batch && (long long)b * t < r1 && !need; ++b) {
      const int lim = min(t, (row_limit[b] + extra + 127) & ~127);
      const long long first = max(r0, (long long)b * t);
      need = first < (long long)b * t + lim;
    }
    if (need) list[1 + atomicAdd(list, 1)] = i;
  }
}

template <int NPASS>
struct FfnSmem {
  // one 32 KB stage, three uses:
  //   G1       : u_hi 8K | W11_hi 8K | u_lo 8K | W11_lo 8K          (slab = 128 rows x 32 k per operand; NPASS 2: no u_lo)
  //   G2       : W_eff_hi 16K | W_eff_lo 16K                        (slab = 256 n-rows x 32 k)
  //   residual : R_hi 8K | R_lo 8K
  static constexpr int kAPlane = kFM * kFK * 2;   // 8 KB
  static constexpr int kW2Plane = kFD * kFK * 2;  // 16 KB
  static constexpr int kStage = 32 * 1024;
  static constexpr int kG1W1Hi = kAPlane, kG1ALo = 2 * kAPlane, kG1W1Lo = 3 * kAPlane;
  static constexpr int kG2Lo = kW2Plane;
  static constexpr int kResLo = kAPlane;
  static constexpr int kMaxF = 2048;
  static constexpr int kI32 = 32 * kFK * 2;       // the 32 x 32 identity block of the residual products (2 KB)
  static constexpr int kF16Stage = kFM * 32 * 2;  // fp16 output plane: one 8 KB staging buffer per epilogue half
  static constexpr int kFixed = 4 * kFStageChunk + kI32 + 2 * kF16Stage + kMaxF * 4 + 3 * kFD * 4 + 2 * 2 * kFM * 8 + 1024;
  static constexpr int kStages = (226 * 1024 - kFixed) / kStage > 6 ? 6 : (226 * 1024 - kFixed) / kStage;
  static constexpr int kOffStaging = kStages * kStage;
  static constexpr int kOffI32 = kOffStaging + 4 * kFStageChunk;  // 1024-aligned (swizzled operand tile)
  static constexpr int kOffF16 = kOffI32 + kI32;
  static constexpr int kOffB1 = kOffF16 + 2 * kF16Stage;          // b1: kMaxF floats
  static constexpr int kOffVec = kOffB1 + kMaxF * 4;              // b2 | gamma | beta
  static constexpr int kOffStats = kOffVec + 3 * kFD * 4;
  static constexpr int kTotal = kStages * kStage + kFixed;
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};

__device__ __forceinline__ void f_tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void f_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void f_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

#ifdef LFS2_FFN_TIMELINE  // diagnostics build only (tools/ffn_ab.py timeline): per-CTA, per-tile clock64 stamps
__device__ long long g_ffn_tl[148][16][8];
#define FFN_TL(slot) do { if (it < 16) g_ffn_tl[blockIdx.x][it][slot] = clock64(); } while (0)
#else
#define FFN_TL(slot) do { } while (0)
#endif

template <int NPASS, bool MC>
__global__ void __launch_bounds__(kFThreads, 1)
ffn_fused_tc_kernel(const __grid_constant__ CUtensorMap map_u_hi, const __grid_constant__ CUtensorMap map_u_lo,
                    const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                    const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                    const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                    const __grid_constant__ CUtensorMap map_ident, const __grid_constant__ CUtensorMap map_o_hi,
                    const __grid_constant__ CUtensorMap map_o_lo, const __grid_constant__ CUtensorMap map_o_f16,
                    const FfnParams p) {
  using L = FfnSmem<NPASS>;
  constexpr int kStages = L::kStages;
  constexpr int kSlabs = kFD / kFK;  // 8 k-slabs per product (K = 256 everywhere)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc1_full[2], v_ready[2], acc2_full, acc2_empty,
      ident_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = p.f / kFC;
  const int ntiles = tile_count(p);
  // work items: MC -> the pair (cluster) walks tile pairs (2i, 2i + 1), this CTA takes the one of its rank
  const int rank = MC ? (int)cluster_ctarank() : 0;
  const int w_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_stride = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_count = MC ? (ntiles + 1) >> 1 : ntiles;
  auto row0_of = [&](int wi) {
    const int ti = MC ? 2 * wi + rank : wi;
    return ti < ntiles ? tile_at(p, ti) * kFM : p.m_pad;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_u_hi);
    prefetch_tmap(&map_w1_hi);
    prefetch_tmap(&map_w2_hi);
    prefetch_tmap(&map_o_hi);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC ? 2 : 1);  // MC: free when BOTH CTAs' MMAs have read the slot (the peer writes into it too)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&v_ready[i], 8);
    }
    mbar_init(&acc2_full, 1);
    mbar_init(&acc2_empty, 8);
    mbar_init(&ident_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  if (warp >= 2) {
    float* b1s = reinterpret_cast<float*>(smem + L::kOffB1);
    float* vec = reinterpret_cast<float*>(smem + L::kOffVec);
    for (int i = threadIdx.x - 64; i < p.f; i += 256) b1s[i] = p.b1 ? p.b1[i] : 0.f;
    for (int i = threadIdx.x - 64; i < kFD; i += 256) {
      vec[i] = p.b2 ? p.b2[i] : 0.f;
      vec[kFD + i] = p.gamma[i];
      vec[2 * kFD + i] = p.beta[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into this CTA
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 256;  // acc1 buffer b: columns [128 b, 128 b + 128)

  if (warp == 0) {
    // ===================== TMA producer: slabs in exactly the order the MMA warp consumes them =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      if (w_first < w_count) {  // the identity block: once per CTA that has work
        mbar_expect_tx(&ident_bar, L::kI32);
        tma_load_3d(smem + L::kOffI32, &map_ident, &ident_bar, 0, 0, 0);
      }
      // MC: the weight maps have half-height boxes; this CTA fetches its half of a slab and multicasts it to the same
      // place in both CTAs (every full barrier still sees a whole slab's bytes)
      auto load_w = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c0, int row, int rows, int plane) {
        if (MC) tma_load_3d_mc(dst + rank * (plane / 2), map, bar, c0, row + rank * (rows / 2), 0, (uint16_t)3);
        else tma_load_3d(dst, map, bar, c0, row, 0);
      };
      auto load_g1 = [&](int r0, int c) {  // u slab + W11 slab, 8 slabs
        for (int ks = 0; ks < kSlabs; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStage;
#ifdef LFS2_FFN_DIAG_NO_WLO_LOADS  // timing diagnostics only: the lo weight planes are not fetched (stale operands)
          mbar_expect_tx(&full_bar[stage], (NPASS == 3 ? 3 : 2) * L::kAPlane);
#else
          mbar_expect_tx(&full_bar[stage], (NPASS + 1) * L::kAPlane);
#endif
          tma_load_3d(st, &map_u_hi, &full_bar[stage], ks * kFK, r0, 0);
          load_w(st + L::kG1W1Hi, &map_w1_hi, &full_bar[stage], ks * kFK, c * kFC, kFC, L::kAPlane);
          if (NPASS == 3) tma_load_3d(st + L::kG1ALo, &map_u_lo, &full_bar[stage], ks * kFK, r0, 0);
#ifndef LFS2_FFN_DIAG_NO_WLO_LOADS
          if (NPASS >= 2) load_w(st + L::kG1W1Lo, &map_w1_lo, &full_bar[stage], ks * kFK, c * kFC, kFC, L::kAPlane);
#endif
          next();
        }
      };
      auto load_g2 = [&](int c) {  // W_eff slab only (A = v lives in tensor memory), 4 slabs
        for (int ks = 0; ks < kFC / kFK; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStage;
#ifdef LFS2_FFN_DIAG_NO_WLO_LOADS
          mbar_expect_tx(&full_bar[stage], L::kW2Plane);
          load_w(st, &map_w2_hi, &full_bar[stage], c * kFC + ks * kFK, 0, kFD, L::kW2Plane);
#else
          mbar_expect_tx(&full_bar[stage], (NPASS >= 2 ? 2 : 1) * L::kW2Plane);
          load_w(st, &map_w2_hi, &full_bar[stage], c * kFC + ks * kFK, 0, kFD, L::kW2Plane);
          if (NPASS >= 2) load_w(st + L::kG2Lo, &map_w2_lo, &full_bar[stage], c * kFC + ks * kFK, 0, kFD, L::kW2Plane);
#endif
          next();
        }
      };
      for (int wi = w_first; wi < w_count; wi += w_stride) {
        const int r0 = row0_of(wi);
        load_g1(r0, 0);
        for (int c = 0; c < nchunks; ++c) {  // same order as the MMA warp: G1(c+1) is issued before G2(c)
          if (c + 1 < nchunks) load_g1(r0, c + 1);
          load_g2(c);
        }
        for (int ks = 0; ks < kSlabs; ++ks) {  // residual: R_hi, R_lo (against the resident identity block)
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::kStage;
          mbar_expect_tx(&full_bar[stage], 2 * L::kAPlane);
          tma_load_3d(st, &map_r_hi, &full_bar[stage], ks * kFK, r0, 0);
          tma_load_3d(st + L::kResLo, &map_r_lo, &full_bar[stage], ks * kFK, r0, 0);
          next();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr int kFmt = NPASS == 2 ? kFmtF16 : kFmtBF16;                      // 2-pass recipe: fp16 operands on both sides
    constexpr uint32_t idesc1 = make_idesc(kFmt, kFM, kFC, 0, 0);              // G1: M128 x N128
    constexpr uint32_t idesc2 = make_idesc(kFmt, kFM, kFD, 0, 0);              // G2: M128 x N256
    constexpr uint32_t idesc_r = make_idesc(kFmtBF16, kFM, 32, 0, 0);          // residual: M128 x N32 per 32-column slab
    const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 512, kSwizzle64);
    const uint64_t d_i32 = make_smem_desc(smem_u32(smem + L::kOffI32), 16, 512, kSwizzle64);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t chunk_ctr = 0;  // chunks since kernel start: buffer = ctr & 1, barrier parity = (ctr >> 1) & 1
    int it = 0;
    auto next = [&]() {
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto release = [&](uint64_t* bar) {  // smem slot reusable once these MMAs retire (MC: in both CTAs, either may refill it)
      if (MC) umma_commit_mc(bar, (uint16_t)3);
      else umma_commit(bar);
    };
    auto issue_g1 = [&](uint32_t ctr) {  // acc1[ctr & 1] = u . W11[chunk]^T
      const uint32_t acc = t_acc1 + 128 * (ctr & 1);
      for (int ks = 0; ks < kSlabs; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_hi = desc_advance(d0, stage * L::kStage);
          const uint64_t w_hi = desc_advance(a_hi, L::kG1W1Hi);
          const uint64_t a_lo = desc_advance(a_hi, L::kG1ALo);
          const uint64_t w_lo = desc_advance(a_hi, L::kG1W1Lo);
          if (ks == 0) umma_f16_c<false>(acc, a_hi, w_hi, idesc1);
          else umma_f16_c<true>(acc, a_hi, w_hi, idesc1);
          umma_f16_c<true>(acc, desc_advance(a_hi, 32), desc_advance(w_hi, 32), idesc1);
          if (NPASS == 3) {
            umma_f16_c<true>(acc, a_lo, w_hi, idesc1);
            umma_f16_c<true>(acc, desc_advance(a_lo, 32), desc_advance(w_hi, 32), idesc1);
          }
          if (NPASS >= 2) {
            umma_f16_c<true>(acc, a_hi, w_lo, idesc1);
            umma_f16_c<true>(acc, desc_advance(a_hi, 32), desc_advance(w_lo, 32), idesc1);
          }
          release(&empty_bar[stage]);
          if (ks + 1 == kSlabs) umma_commit(&acc1_full[ctr & 1]);
        }
        __syncwarp();
        next();
      }
    };
    for (int wi = w_first; wi < w_count; wi += w_stride, ++it) {
      if (lane == 0) FFN_TL(5);
      issue_g1(chunk_ctr);
      for (int c = 0; c < nchunks; ++c, ++chunk_ctr) {
        // G1 of the next chunk goes first: it runs on the tensor pipe while the epilogue warps convert chunk c.
        // (its accumulator buffer was last read by G2(c-1), issued earlier: tcgen05.mma executes in issue order)
        if (c + 1 < nchunks) issue_g1(chunk_ctr + 1);
        // ---- G2: acc2 += v[c] . W_eff[:, c]^T, A = v from tensor memory (written in place over acc1) ----
        mbar_wait(&v_ready[chunk_ctr & 1], (chunk_ctr >> 1) & 1);
        // the first G2 of a tile overwrites acc2: the previous tile's LayerNorm epilogue must have drained it
        if (c == 0) mbar_wait(&acc2_empty, (it & 1) ^ 1);
        if (c == 0 && lane == 0) FFN_TL(6);
        tc_fence_after();
        const uint32_t vbase = t_acc1 + 128 * (chunk_ctr & 1);
        for (int ks = 0; ks < kFC / kFK; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t w_hi = desc_advance(d0, stage * L::kStage);
            const uint64_t w_lo = desc_advance(w_hi, L::kG2Lo);
            const uint32_t v_hi = vbase + 32 * ks, v_lo = v_hi + 16;  // packed pairs: 8 columns per k16
            if (c == 0 && ks == 0) umma_f16_ts_c<false>(t_acc2, v_hi, w_hi, idesc2);
            else umma_f16_ts_c<true>(t_acc2, v_hi, w_hi, idesc2);
            umma_f16_ts_c<true>(t_acc2, v_hi + 8, desc_advance(w_hi, 32), idesc2);
            if (NPASS == 3) {
              umma_f16_ts_c<true>(t_acc2, v_lo, w_hi, idesc2);
              umma_f16_ts_c<true>(t_acc2, v_lo + 8, desc_advance(w_hi, 32), idesc2);
            }
            if (NPASS >= 2) {
              umma_f16_ts_c<true>(t_acc2, v_hi, w_lo, idesc2);
              umma_f16_ts_c<true>(t_acc2, v_hi + 8, desc_advance(w_lo, 32), idesc2);
            }
            release(&empty_bar[stage]);
          }
          __syncwarp();
          next();
        }
      }
      // ---- residual x1 on the tensor core: acc2[:, 32 ks .. 32 ks + 32) += R_hi[ks] . I32 + R_lo[ks] . I32 ----
      if (it == 0) mbar_wait(&ident_bar, 0);
      if (lane == 0) FFN_TL(7);
      for (int ks = 0; ks < kSlabs; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_hi = desc_advance(d0, stage * L::kStage);
          const uint64_t a_lo = desc_advance(a_hi, L::kResLo);
          const uint32_t acc = t_acc2 + 32 * ks;
          umma_f16_c<true>(acc, a_hi, d_i32, idesc_r);
          umma_f16_c<true>(acc, desc_advance(a_hi, 32), desc_advance(d_i32, 32), idesc_r);
          umma_f16_c<true>(acc, a_lo, d_i32, idesc_r);
          umma_f16_c<true>(acc, desc_advance(a_lo, 32), desc_advance(d_i32, 32), idesc_r);
          release(&empty_bar[stage]);
          if (ks + 1 == kSlabs) umma_commit(&acc2_full);
        }
        __syncwarp();
        next();
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool issuer = (warp == 2 || warp == 6) && lane == 0;
    const float* b1s = reinterpret_cast<const float*>(smem + L::kOffB1);
    const float* vec = reinterpret_cast<const float*>(smem + L::kOffVec);
    float2* stats = reinterpret_cast<float2*>(smem + L::kOffStats);
    uint8_t* staging = smem + L::kOffStaging + half * 2 * kFStageChunk;
    uint32_t chunk_ctr = 0, st_ctr = 0;
    int it = 0;
    float v[32];
    for (int wi = w_first; wi < w_count; wi += w_stride, ++it) {
      const int r0 = row0_of(wi);
      if (warp == 2 && lane == 0) FFN_TL(0);
      // ---- E1 per F chunk: acc1 -> relu(acc1 + b1) as bf16 hi/lo pairs, in place ----
      for (int c = 0; c < nchunks; ++c, ++chunk_ctr) {
        mbar_wait(&acc1_full[chunk_ctr & 1], (chunk_ctr >> 1) & 1);
        tc_fence_after();
#ifdef LFS2_FFN_DIAG_NO_E1  // timing diagnostics only (tools/ffn_ab.py): wrong results
        if (false)
#endif
#pragma unroll 1
        for (int j = 2 * half; j < 2 * half + 2; ++j) {  // 4 blocks of 32 columns per chunk, 2 per warp of the pair
          const uint32_t ta = t_acc1 + 128 * (chunk_ctr & 1) + lane_off + 32 * j;
          tmem_ld32(ta, v);
          uint32_t hi[16], lo[16];
          const float4* bb = reinterpret_cast<const float4*>(b1s + c * kFC + 32 * j);  // broadcast 16-byte reads
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 b = bb[e];
            const float x0 = fmaxf(v[4 * e] + b.x, 0.f), x1 = fmaxf(v[4 * e + 1] + b.y, 0.f);
            const float x2 = fmaxf(v[4 * e + 2] + b.z, 0.f), x3 = fmaxf(v[4 * e + 3] + b.w, 0.f);
            if (NPASS == 2) {
              hi[2 * e] = pack_f16_sat(x0, x1);
              hi[2 * e + 1] = pack_f16_sat(x2, x3);
            } else {
              split_pack2(x0, x1, hi[2 * e], lo[2 * e]);
              split_pack2(x2, x3, hi[2 * e + 1], lo[2 * e + 1]);
            }
          }
          f_tmem_st16(ta, hi);
          if (NPASS == 3) f_tmem_st16(ta + 16, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_ready[chunk_ctr & 1]);
      }
      // ---- final epilogue: acc2 + b2 -> LayerNorm -> hi/lo planes ----
      if (warp == 2 && lane == 0) FFN_TL(1);
      mbar_wait(&acc2_full, it & 1);
      if (warp == 2 && lane == 0) FFN_TL(2);
      tc_fence_after();
      const uint32_t ta2 = t_acc2 + lane_off;
      float s = 0.f, q = 0.f;
#ifdef LFS2_FFN_DIAG_NO_LN
      if (false)
#endif
#pragma unroll 1
      for (int j = 4 * half; j < 4 * half + 4; ++j) {
        tmem_ld32(ta2 + 32 * j, v);
        const float4* b4 = reinterpret_cast<const float4*>(vec + 32 * j);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 b = b4[e];
          const float x0 = v[4 * e] + b.x, x1 = v[4 * e + 1] + b.y, x2 = v[4 * e + 2] + b.z, x3 = v[4 * e + 3] + b.w;
          s += x0; q = fmaf(x0, x0, q);
          s += x1; q = fmaf(x1, x1, q);
          s += x2; q = fmaf(x2, x2, q);
          s += x3; q = fmaf(x3, x3, q);
        }
      }
      float2* stt = stats + (it & 1) * 2 * kFM;
      stt[half * kFM + r] = make_float2(s, q);
      f_bar_sync(3, 256);
      if (warp == 2 && lane == 0) FFN_TL(3);
      const float2 o = stt[(half ^ 1) * kFM + r];
      s += o.x;
      q += o.y;
      const float mean = s * (1.f / kFD);
      const float rstd = rsqrtf(fmaxf(q * (1.f / kFD) - mean * mean, 0.f) + p.eps);
#ifdef LFS2_FFN_DIAG_NO_LN
      if (false)
#endif
#pragma unroll 1
      for (int j = 4 * half; j < 4 * half + 4; ++j) {
        tmem_ld32(ta2 + 32 * j, v);
        {
          const float4* b4 = reinterpret_cast<const float4*>(vec + 32 * j);
          const float4* g4 = reinterpret_cast<const float4*>(vec + kFD + 32 * j);
          const float4* e4 = reinterpret_cast<const float4*>(vec + 2 * kFD + 32 * j);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 b = b4[e], g = g4[e], bt = e4[e];
            v[4 * e] = (v[4 * e] + b.x - mean) * rstd * g.x + bt.x;
            v[4 * e + 1] = (v[4 * e + 1] + b.y - mean) * rstd * g.y + bt.y;
            v[4 * e + 2] = (v[4 * e + 2] + b.z - mean) * rstd * g.z + bt.z;
            v[4 * e + 3] = (v[4 * e + 3] + b.w - mean) * rstd * g.w + bt.w;
          }
        }
        uint8_t* sb = staging + (st_ctr & 1) * kFStageChunk;
        ++st_ctr;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split_pack2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        uint8_t* rh = sb + r * 64;
        uint8_t* rl = rh + kFStageChunk / 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = (i ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
          *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
        }
        fence_proxy_async_smem();
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        f_bar_sync(1 + half, 128);
#ifdef LFS2_FFN_DIAG_NO_STORES
        if (false)
#endif
        if (issuer) {
          f_tma_store_3d(&map_o_hi, sb, 32 * j, r0, 0);
          f_tma_store_3d(&map_o_lo, sb + kFStageChunk / 2, 32 * j, r0, 0);
        }
        if (p.out_f16) {
          // single staging buffer per half: every earlier store has been read (wait_group.read above).  (Storing the
          // plane straight from registers -- 16 bytes per row and instruction, half-filled sectors -- cost the kernel 7 %.)
          uint8_t* rf = smem + L::kOffF16 + half * L::kF16Stage + r * 64;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int u = (i ^ ((r >> 1) & 3)) << 4;
            *reinterpret_cast<uint4*>(rf + u) =
                make_uint4(pack_f16_sat(v[8 * i], v[8 * i + 1]), pack_f16_sat(v[8 * i + 2], v[8 * i + 3]),
                           pack_f16_sat(v[8 * i + 4], v[8 * i + 5]), pack_f16_sat(v[8 * i + 6], v[8 * i + 7]));
          }
          fence_proxy_async_smem();
          f_bar_sync(1 + half, 128);
          if (issuer) f_tma_store_3d(&map_o_f16, smem + L::kOffF16 + half * L::kF16Stage, 32 * j, r0, 0);
        }
        if (issuer) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) FFN_TL(4);
      if (lane == 0) mbar_arrive(&acc2_empty);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer may still arrive on this CTA's barriers until it has finished too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NPASS, bool MC>
static int launch_ffn(const CUtensorMap* m, const FfnParams& p, cudaStream_t s) {
  using L = FfnSmem<NPASS>;
  auto kern = ffn_fused_tc_kernel<NPASS, MC>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) {
      set_error("ffn_fused_tc: cannot reserve %d bytes of shared memory", L::kTotal);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  if (!MC) {
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    kern<<<grid, kFThreads, L::kTotal, s>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], p);
  } else {  // clusters of two CTAs (one per SM): pairs of row tiles share the multicast weight slabs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() & ~1);
    cfg.blockDim = dim3(kFThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], p) !=
        cudaSuccess) {
      set_error("ffn_fused_tc: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      return LFS2_ERR_CUDA;
    }
  }
  LFS2_CHECK_LAUNCH("ffn_fused_tc");
  return LFS2_OK;
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" int lfs2_ffn_fused_tc(const void* u_hi, const void* u_lo, int m, const void* w1_hi, const void* w1_lo, int f,
                                 const float* b1, const void* w2_hi, const void* w2_lo, const float* b2,
                                 const void* res_hi, const void* res_lo, const void* ident_hi, const float* gamma,
                                 const float* beta, float eps, void* out_hi, void* out_lo, int npass, void* stream) {
  return lfs2_ffn_fused_tc_limited(u_hi, u_lo, 1, m, w1_hi, w1_lo, f, b1, w2_hi, w2_lo, b2, res_hi, res_lo, ident_hi, gamma,
                                   beta, eps, out_hi, out_lo, npass, nullptr, 0, nullptr, stream);
}

#ifdef LFS2_FFN_TIMELINE
extern "C" __attribute__((visibility("default"))) int lfs2_ffn_timeline(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_ffn_tl, sizeof(g_ffn_tl)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" long long lfs2_ffn_fused_tc_limited_workspace_bytes(int batch, int t) {
  return batch > 0 && t > 0 ? (1 + (long long)ceil_div((long long)batch * t, kFM)) * (long long)sizeof(int) : 0;
}

extern "C" int lfs2_ffn_fused_tc_limited(const void* u_hi, const void* u_lo, int batch, int t, const void* w1_hi,
                                         const void* w1_lo, int f, const float* b1, const void* w2_hi, const void* w2_lo,
                                         const float* b2, const void* res_hi, const void* res_lo, const void* ident_hi,
                                         const float* gamma, const float* beta, float eps, void* out_hi, void* out_lo,
                                         int npass, const int* row_limit, int limit_extra, void* workspace, void* stream) {
  return lfs2_ffn_fused_tc_ex(u_hi, u_lo, batch, t, w1_hi, w1_lo, f, b1, w2_hi, w2_lo, b2, res_hi, res_lo, ident_hi, gamma,
                              beta, eps, out_hi, out_lo, nullptr, npass, row_limit, limit_extra, workspace, stream);
}

extern "C" int lfs2_ffn_fused_tc_ex(const void* u_hi, const void* u_lo, int batch, int t, const void* w1_hi,
                                    const void* w1_lo, int f, const float* b1, const void* w2_hi, const void* w2_lo,
                                    const float* b2, const void* res_hi, const void* res_lo, const void* ident_hi,
                                    const float* gamma, const float* beta, float eps, void* out_hi, void* out_lo,
                                    void* out_f16, int npass, const int* row_limit, int limit_extra, void* workspace,
                                    void* stream) {
  LFS2_REQUIRE(batch >= 0 && t >= 0 && (long long)batch * t <= 0x7fffffffLL, LFS2_ERR_INVALID_ARG, "ffn_fused_tc: bad shape");
  LFS2_REQUIRE(!row_limit || workspace, LFS2_ERR_INVALID_ARG, "ffn_fused_tc: a row-limited launch needs its workspace");
  const int m = batch * t;
  LFS2_REQUIRE(u_hi && w1_hi && w2_hi && res_hi && res_lo && ident_hi && gamma && beta && out_hi && out_lo,
               LFS2_ERR_INVALID_ARG, "ffn_fused_tc: null pointer");
  LFS2_REQUIRE(npass >= 1 && npass <= 3, LFS2_ERR_INVALID_ARG, "ffn_fused_tc: npass must be 1, 2 or 3");
  LFS2_REQUIRE(npass == 1 || (w1_lo && w2_lo), LFS2_ERR_INVALID_ARG, "ffn_fused_tc: npass >= 2 needs the weight lo planes");
  LFS2_REQUIRE(npass != 3 || u_lo, LFS2_ERR_INVALID_ARG, "ffn_fused_tc: npass=3 needs the lo plane of u");
  if (m == 0) return LFS2_OK;
  LFS2_REQUIRE(m > 0 && f > 0, LFS2_ERR_INVALID_ARG, "ffn_fused_tc: bad shape");
  LFS2_REQUIRE(f % kFC == 0 && f <= FfnSmem<3>::kMaxF, LFS2_ERR_UNSUPPORTED,
               "ffn_fused_tc: hidden width %d must be a multiple of %d and <= %d (model width is fixed at %d)", f, kFC,
               FfnSmem<3>::kMaxF, kFD);
  LFS2_REQUIRE(aligned16(u_hi) && aligned16(w1_hi) && aligned16(w2_hi) && aligned16(res_hi) && aligned16(res_lo) &&
                   aligned16(out_hi) && aligned16(out_lo) && (!u_lo || aligned16(u_lo)) && aligned16(out_f16),
               LFS2_ERR_INVALID_ARG, "ffn_fused_tc: pointers must be 16-byte aligned");
  // 2-CTA multicast variant: at least one pair of row tiles per cluster (LFS2_FFN_MULTICAST=0 switches it off: A/B runs)
  static int mc_on = -1;
  if (mc_on < 0) {
    const char* e = getenv("LFS2_FFN_MULTICAST");
    mc_on = (e && e[0] == '0') ? 0 : 1;
  }
  const bool mc = mc_on == 1 && ceil_div(m, kFM) >= 2 * num_sms();
  const uint32_t w1_box = mc ? kFC / 2 : kFC, w2_box = mc ? kFD / 2 : kFD;  // MC: each CTA fetches half a weight slab
  CUtensorMap maps[12];
  bool ok = make_tmap_3d(&maps[0], u_hi, kFD, m, 1, kFK, kFM, 64) && make_tmap_3d(&maps[2], w1_hi, kFD, f, 1, kFK, w1_box, 64) &&
            make_tmap_3d(&maps[4], w2_hi, f, kFD, 1, kFK, w2_box, 64) && make_tmap_3d(&maps[6], res_hi, kFD, m, 1, kFK, kFM, 64) &&
            make_tmap_3d(&maps[7], res_lo, kFD, m, 1, kFK, kFM, 64) &&
            make_tmap_3d(&maps[8], ident_hi, kFD, kFD, 1, kFK, 32, 64) &&
            make_tmap_3d(&maps[9], out_hi, kFD, m, 1, 32, kFM, 64) && make_tmap_3d(&maps[10], out_lo, kFD, m, 1, 32, kFM, 64);
  maps[1] = maps[0];
  maps[3] = maps[2];
  maps[5] = maps[4];
  maps[11] = maps[9];
  if (npass == 3) ok = ok && make_tmap_3d(&maps[1], u_lo, kFD, m, 1, kFK, kFM, 64);
  if (npass >= 2)
    ok = ok && make_tmap_3d(&maps[3], w1_lo, kFD, f, 1, kFK, w1_box, 64) && make_tmap_3d(&maps[5], w2_lo, f, kFD, 1, kFK, w2_box, 64);
  if (out_f16) ok = ok && make_tmap_3d(&maps[11], out_f16, kFD, m, 1, 32, kFM, 64);
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "ffn_fused_tc: cuTensorMapEncodeTiled failed");
  FfnParams p;
  p.m = m; p.f = f; p.total_tiles = ceil_div(m, kFM); p.m_pad = p.total_tiles * kFM;
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps; p.out_f16 = out_f16 != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  p.tile_list = nullptr;
  if (!row_limit && workspace) {
    p.tile_list = reinterpret_cast<const int*>(workspace);  // a list built earlier for the same (batch, t, limit): reuse
  } else if (row_limit) {  // compact list of the needed row tiles, built on the device (no host read-back)
    int* list = reinterpret_cast<int*>(workspace);
    if (cudaMemsetAsync(list, 0, sizeof(int), s) != cudaSuccess) {
      set_error("ffn_fused_tc: cudaMemsetAsync failed");
      return LFS2_ERR_CUDA;
    }
    ffn_tile_list_kernel<<<ceil_div(p.total_tiles, 256), 256, 0, s>>>(row_limit, limit_extra, batch, t, p.total_tiles, list);
    LFS2_CHECK_LAUNCH("ffn_tile_list");
    p.tile_list = list;
  }
  if (mc) return npass == 3 ? launch_ffn<3, true>(maps, p, s) : (npass == 2 ? launch_ffn<2, true>(maps, p, s) : launch_ffn<1, true>(maps, p, s));
  return npass == 3 ? launch_ffn<3, false>(maps, p, s) : (npass == 2 ? launch_ffn<2, false>(maps, p, s) : launch_ffn<1, false>(maps, p, s));
}
