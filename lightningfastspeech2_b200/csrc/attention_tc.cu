// tcgen05 flash attention for the FFTBlock (head_dim 128), operands as bf16 hi/lo planes.
//
//   one CTA = 128 queries of one (utterance, head); key tiles of 64 stream through shared
//   memory by TMA; S = Q.K^T and O += P.V run on the 5th-gen tensor cores with both A
//   operands (Q, P) read from TENSOR MEMORY (tcgen05.mma "TS" form), so shared memory only
//   feeds the B operands (K K-major, V MN-major straight from the row-major qkv tensor --
//   no transposed copy of V exists anywhere).
//
//   TMEM columns (512):  [0,128)   Q   bf16 pairs: hi plane cols 0..63, lo plane cols 64..127
//                        [128,..)  S/P buffers: S_J fp32 (a group of kG 64-key tiles, see AttnCfg) is
//                                  overwritten in place by P_J (hi: first half of the columns, lo: second half);
//                                  two 128-column buffers (kG = 2)
//                        then      O   fp32 accumulator (128 head-dim columns)
//
//   warps: 0 = TMA producer of the Q tile and the K ring, 10 = TMA producer of the V ring
//   (3 stages each), 1 = MMA issuer (one elected thread),
//   2..9 = softmax (thread = query row; TMEM lane quadrant = warp & 3; the two warps of a
//   quadrant split each group's keys and the 128 O columns, and agree on the row maximum through
//   shared memory + one 64-thread named barrier per step).
//   tensor-pipe order:  QK_0 .. QK_{n-1} | PV_0 QK_n | PV_1 QK_{n+1} | ...  (indices = tile groups, n = kSBufs)
//   so the softmax warps always find S ready and only PV waits for them.  The Q tile is
//   fetched by TMA into (still idle) V-ring memory and moved to TMEM by the softmax warps; the
//   normalised O leaves through swizzled staging in the (idle) K ring and TMA tensor stores.  Online softmax with a lazily updated exponent reference: the O
//   accumulator is rescaled in TMEM only when a row's logits outgrow the reference by 2^8.
//
//   NPASS = 3: hi.hi + lo.hi + hi.lo for both products (fp32-parity mode); NPASS = 1: hi.hi.
//   The softmax itself (max, exp2, sum, 1/l) is fp32.  PAD keys (key_padding_mask) and keys
//   beyond T get -inf; rows whose keys are all masked produce NaN like the reference.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kAQ = 128;        // queries per CTA (UMMA M)
constexpr int kAK = 64;         // keys per tile
constexpr int kDH = 128;        // head dim
constexpr int kKVStages = 3;
constexpr int kAttnTcThreads = 352;  // K-TMA, MMA, 8 softmax warps (two per TMEM lane quadrant), V-TMA
constexpr int kTileBytes = kAK * kDH * 2;          // one plane of a K or V tile: 16 KB
// S/P buffers in TMEM.  A buffer holds a GROUP of kG consecutive key tiles (kG * 64 fp32 columns); the softmax
// warps work through one group per step, so their per-step fixed latencies (TMEM load, maximum exchange, barrier,
// TMEM store, mbarrier) are paid once per kG * 64 keys.  Measured on the C2 decoder shape (tools/attn_ab.py, one
// launch): two 2-tile buffers vs three 1-tile buffers = 0.371 vs 0.458 ms in bf16 mode (bound by the softmax
// warps), 0.595 vs 0.607 ms in fp32-parity mode (bound by the tensor pipe).
#ifndef LFS2_ATTN_G1  // tuning knobs (tools/attn_ab.py builds the alternatives): tiles per buffer / buffers, per mode
#define LFS2_ATTN_G1 2
#define LFS2_ATTN_BUFS1 2
#endif
#ifndef LFS2_ATTN_G3
#define LFS2_ATTN_G3 2
#define LFS2_ATTN_BUFS3 2
#endif
template <int NPASS>
struct AttnCfg {
  static constexpr int kG = NPASS == 1 ? LFS2_ATTN_G1 : LFS2_ATTN_G3;              // key tiles per S/P buffer
  static constexpr int kSBufs = NPASS == 1 ? LFS2_ATTN_BUFS1 : LFS2_ATTN_BUFS3;    // QK runs up to kSBufs groups ahead of PV
  static constexpr int kSW = kG * kAK;              // keys (= fp32 columns) per buffer
  static constexpr int kE = kSW / 2;                // keys per softmax thread and step
  static constexpr uint32_t kColO = 128 + kSBufs * kSW;
  static_assert(kColO + kDH <= 512, "tensor memory budget");
};
constexpr uint32_t kColQ = 0, kColS = 128;
constexpr int kMaxSBufs = 3;

// 32 lanes x 32 columns without the trailing wait (caller batches the wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32_u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st16_u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void pair_bar_sync(int quad) {
  asm volatile("bar.sync %0, 64;" ::"r"(4 + quad) : "memory");
}

struct AttnTcParams {
  const __nv_bfloat16* qkv_hi;  // (B, T, 3d)
  const __nv_bfloat16* qkv_lo;
  const uint8_t* kpm;           // (B, T) 1 = PAD, or null
  const int* kend;              // (B): 1 + index of the last non-PAD key (0 = all masked)
  __nv_bfloat16* ctx_hi;        // (B, T, d)
  __nv_bfloat16* ctx_lo;
  float* ctx_f32;               // (B, T, d) or null
  int t, d;
  float scale_log2e;            // head_dim^-1/2 * log2(e)
  const int* row_limit;         // null, or (B): query tiles that start at or after row_limit[b] + limit_extra are skipped
  int limit_extra;
};

// last valid key + 1 per utterance; one warp per utterance
__global__ void attn_kend_kernel(const uint8_t* __restrict__ kpm, int* __restrict__ kend, int batch, int t) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= batch) return;
  int last = 0;
  if (!kpm) {
    last = t;
  } else {
    for (int i = lane; i < t; i += 32)
      if (!kpm[(size_t)b * t + i]) last = i + 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if (lane == 0) kend[b] = last;
}

__device__ __forceinline__ void tma_store_3d_a(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// two fp32 -> packed fp16 pair (element a in the low half)
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// FMT = operand format of Q, K, V and P: kFmtBF16 (hi [+ lo] planes) or kFmtF16 (NPASS = 1 only: ONE fp16 plane written
// by lfs2_gemm_tc_ex(LFS2_OUT_F16); 11 significant bits instead of 8 -- see profiles/r2c_precision_emulation_*.txt)
template <int NPASS, int FMT>
__global__ void __launch_bounds__(kAttnTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                    const AttnTcParams p) {
  using Cfg = AttnCfg<NPASS>;
  constexpr int kG = Cfg::kG, kSBufs = Cfg::kSBufs, kSW = Cfg::kSW, kE = Cfg::kE;
  constexpr uint32_t kColO = Cfg::kColO;
  constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  constexpr int kStageBytes = kPlanes * kTileBytes;  // K (or V) tile, hi [+ lo]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                // K ring; reused as the O staging buffer at the end
  uint8_t* sV = smem + kKVStages * kStageBytes;      // V ring; holds the Q tile until it has moved to TMEM
  __shared__ __align__(8) uint64_t k_full[kKVStages], k_empty[kKVStages], v_full[kKVStages], v_empty[kKVStages];
  __shared__ __align__(8) uint64_t q_smem_full, q_full, s_full[kMaxSBufs], p_full[kMaxSBufs], pv_done[kMaxSBufs], o_final;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][2][kAQ];  // [tile parity][half][row]: row maxima / partial sums between the warp pair

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
  // query rows the caller does not need (rows past an utterance's end + conv halo; same rule as the row-limited GEMM)
  if (p.row_limit && q0 >= __ldg(p.row_limit + b) + p.limit_extra) return;
  const int kend = p.kend[b];
  const int ntiles = (kend + kAK - 1) / kAK;
  const int col_q = h * kDH, col_k = p.d + h * kDH, col_v = 2 * p.d + h * kDH;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_hi);
    prefetch_tmap(&map_q_hi);
    if (NPASS == 3) prefetch_tmap(&map_lo);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(&q_smem_full, 1);
    mbar_init(&q_full, 8);
    for (int i = 0; i < kSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(&o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer: Q tile, then the K ring =====================
    if (lane == 0) {
      // Q (128 rows x 128 head-dim columns per plane) as two [64 cols x 128 rows] SWIZZLE_128B boxes
      mbar_expect_tx(&q_smem_full, kPlanes * 2 * 16384);
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_3d(sV + bx * 16384, &map_q_hi, &q_smem_full, col_q + bx * 64, q0, b);
        if (NPASS == 3) tma_load_3d(sV + 32768 + bx * 16384, &map_q_lo, &q_smem_full, col_q + bx * 64, q0, b);
      }
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % kKVStages;
        mbar_wait(&k_empty[st], ((j / kKVStages) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], kStageBytes);
        uint8_t* dk = sK + st * kStageBytes;
#pragma unroll
        for (int c = 0; c < kDH / 32; ++c) {
          tma_load_3d(dk + c * 4096, &map_hi, &k_full[st], col_k + c * 32, j * kAK, b);
          if (NPASS == 3) tma_load_3d(dk + kTileBytes + c * 4096, &map_lo, &k_full[st], col_k + c * 32, j * kAK, b);
        }
      }
    }
  } else if (warp == 10) {
    // ===================== TMA producer: the V ring (after Q has left shared memory) =====================
    if (lane == 0) {
      mbar_wait(&q_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % kKVStages;
        mbar_wait(&v_empty[st], ((j / kKVStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], kStageBytes);
        uint8_t* dv = sV + st * kStageBytes;
#pragma unroll
        for (int c = 0; c < kDH / 32; ++c) {
          tma_load_3d(dv + c * 4096, &map_hi, &v_full[st], col_v + c * 32, j * kAK, b);
          if (NPASS == 3) tma_load_3d(dv + kTileBytes + c * 4096, &map_lo, &v_full[st], col_v + c * 32, j * kAK, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) control flow so descriptors stay in uniform
    // registers; one elected lane issues the tcgen05 instructions.  QK products run up to
    // kSBufs tiles ahead of the PV products, so the softmax warps always find S ready.
    if (ntiles > 0) {
      constexpr uint32_t idesc_qk = make_idesc(FMT, kAQ, kAK, 0, 0);  // A: TMEM, B: K tile, K-major
      constexpr uint32_t idesc_pv = make_idesc(FMT, kAQ, kDH, 0, 1);  // A: TMEM, B: V tile, MN-major
      const uint32_t tq = tmem_base + kColQ, to = tmem_base + kColO;
      const uint64_t dk0 = make_smem_desc(smem_u32(sK), 16, 512, kSwizzle64);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 4096, 512, kSwizzle64);

      const int npairs = (ntiles + kG - 1) / kG;
      // S of tile group jp: tiles kG*jp .. kG*jp + kG-1 (those that exist) into the kG parts of S/P buffer jp % kSBufs
      auto issue_qk = [&](int jp) {
        const int sb = jp % kSBufs;
#pragma unroll
        for (int hh = 0; hh < kG; ++hh) {
          const int j = kG * jp + hh;
          if (j >= ntiles) break;
          const int st = j % kKVStages;
          mbar_wait(&k_full[st], (j / kKVStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t ts = tmem_base + kColS + sb * kSW + hh * kAK;
            const uint64_t dkh = desc_advance(dk0, st * kStageBytes), dkl = desc_advance(dkh, kTileBytes);
#pragma unroll
            for (int i = 0; i < kDH / 16; ++i) {
              const uint32_t off = (i >> 1) * 4096 + (i & 1) * 32;
              if (i == 0) umma_f16_ts_c<false>(ts, tq, dkh, idesc_qk);
              else umma_f16_ts_c<true>(ts, tq + i * 8, desc_advance(dkh, off), idesc_qk);
              if (NPASS == 3) {
                umma_f16_ts_c<true>(ts, tq + 64 + i * 8, desc_advance(dkh, off), idesc_qk);
                umma_f16_ts_c<true>(ts, tq + i * 8, desc_advance(dkl, off), idesc_qk);
              }
            }
            umma_commit(&k_empty[st]);
            if (hh == kG - 1 || j + 1 == ntiles) umma_commit(&s_full[sb]);  // the group's last existing tile
          }
          __syncwarp();
        }
      };

      mbar_wait(&q_full, 0);
      tc_fence_after();
      for (int jp = 0; jp < kSBufs && jp < npairs; ++jp) issue_qk(jp);
      for (int jp = 0; jp < npairs; ++jp) {
        const int sb = jp % kSBufs;
        mbar_wait(&p_full[sb], (jp / kSBufs) & 1);
#pragma unroll
        for (int hh = 0; hh < kG; ++hh) {
          const int j = kG * jp + hh;
          if (j >= ntiles) break;
          const int st = j % kKVStages;
          mbar_wait(&v_full[st], (j / kKVStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t tp = tmem_base + kColS + sb * kSW + hh * 32;  // P hi: first kSW/2 cols of the buffer, P lo: the rest
            const uint64_t dvh = desc_advance(dv0, st * kStageBytes), dvl = desc_advance(dvh, kTileBytes);
#pragma unroll
            for (int i = 0; i < kAK / 16; ++i) {
              if (i == 0) umma_f16_ts(to, tp, dvh, idesc_pv, j ? 1u : 0u);
              else umma_f16_ts_c<true>(to, tp + i * 8, desc_advance(dvh, i * 1024), idesc_pv);
              if (NPASS == 3) {
                umma_f16_ts_c<true>(to, tp + kSW / 2 + i * 8, desc_advance(dvh, i * 1024), idesc_pv);
                umma_f16_ts_c<true>(to, tp + i * 8, desc_advance(dvl, i * 1024), idesc_pv);
              }
            }
            umma_commit(&v_empty[st]);
            if (hh == kG - 1 || j + 1 == ntiles) {
              umma_commit(&pv_done[sb]);
              if (j + 1 == ntiles) umma_commit(&o_final);  // every product of this CTA has landed
            }
          }
          __syncwarp();
        }
        if (jp + kSBufs < npairs) issue_qk(jp + kSBufs);
      }
    }
  } else {
    // ===================== softmax warps 2..9: thread = query row =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which tile of the pair (64 of its 128 keys) / which 64 of the 128 O columns
    const int r = quad * 32 + lane;
    const int tq_row = q0 + r;
    const bool row_ok = tq_row < p.t;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;

    // ---- Q row: swizzled shared memory -> TMEM (bf16 pairs; half 0 moves the hi plane -> cols 0..63,
    //      half 1 the lo plane -> cols 64..127) ----
    {
      mbar_wait(&q_smem_full, 0);
      if (half == 0 || NPASS == 3) {
        const uint8_t* qrow = sV + half * 32768 + r * 128;
        uint32_t w[32];
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(qrow + bx * 16384 + ((i ^ (r & 7)) << 4));
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
          tmem_st32_u(tmem_base + kColQ + half * 64 + bx * 32 + lane_off, w);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full);
    }

    float m_run = -INFINITY;  // exponent reference of this row (shared by both halves)
    float l_run = 0.f;        // partial row sum over this half's keys
    const float c = p.scale_log2e;
    const uint8_t* mrow = p.kpm ? p.kpm + (size_t)b * p.t : nullptr;
    // "key is masked" bytes of this lane's keys in the NEXT tile group, prefetched one step ahead and only
    // TESTED at the top of the next iteration, so the loads' latency hides behind a whole step of work.
    // This warp's keys of group jp: jp*kSW + half*kE + [0,kE); keys of a tile past the last one (bf16 mode, odd
    // tile count) are all "masked": their S columns were never written and their P must be 0.
    const int npairs = (ntiles + kG - 1) / kG;
    auto key_masked = [&](int jp, int sub) -> uint32_t {
      const int k0 = jp * kSW + half * kE;
      const int k1 = k0 + sub * 32 + lane;
      uint32_t v = 1u;
      if (k1 < p.t && k0 / kAK < ntiles) v = mrow ? (uint32_t)mrow[k1] : 0u;
      return v;
    };
    uint32_t next_m0 = npairs > 0 ? key_masked(0, 0) : 1u;
    uint32_t next_m1 = (kE > 32 && npairs > 0) ? key_masked(0, 1) : 0u;

    for (int jp = 0; jp < npairs; ++jp) {
      const int sb = jp % kSBufs;
      const uint32_t mbits0 = __ballot_sync(0xffffffffu, next_m0 != 0u);
      const uint32_t mbits1 = kE > 32 ? __ballot_sync(0xffffffffu, next_m1 != 0u) : 0u;
      if (jp + 1 < npairs) {
        next_m0 = key_masked(jp + 1, 0);
        if (kE > 32) next_m1 = key_masked(jp + 1, 1);
      }
      mbar_wait(&s_full[sb], (jp / kSBufs) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_base + kColS + sb * kSW + lane_off;
      float s[kE];
      tmem_ld32_nowait(ts + half * kE, s);
      if (kE > 32) tmem_ld32_nowait(ts + half * kE + 32, s + (kE > 32 ? 32 : 0));
      tmem_wait_ld();
      if (mbits0 | mbits1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if ((mbits0 >> i) & 1u) s[i] = -INFINITY;
          if (kE > 32 && ((mbits1 >> i) & 1u)) s[(kE > 32 ? 32 : 0) + i] = -INFINITY;
        }
      }
      float tmax0 = s[0], tmax1 = s[1];
#pragma unroll
      for (int i = 2; i < kE; i += 2) {
        tmax0 = fmaxf(tmax0, s[i]);
        tmax1 = fmaxf(tmax1, s[i + 1]);
      }
      float tmax = fmaxf(tmax0, tmax1);
      // row maximum over all 128 keys: exchange with the partner warp (parity double buffer); the
      // barrier also orders "both halves have read S" before either overwrites it with P
      xch[jp & 1][half][r] = tmax;
      pair_bar_sync(quad);
      tmax = fmaxf(tmax, xch[jp & 1][half ^ 1][r]);
      // Lazy rescale: m_run is the exponent reference, not necessarily the true running maximum.
      // It is only moved (and O, l rescaled) when some row's logits exceed it by more than 2^8
      // in the exp2 domain, so p <= 256 always and the O accumulator is almost never touched;
      // softmax is shift-invariant, so the result is exact either way.  Both halves see the same
      // tmax and m_run, hence take the same (warp-uniform) branch.
      const bool grow = (jp > 0) && ((tmax - m_run) * c > 8.f);  // (x - -inf) = inf: first finite tile moves it
      if (jp == 0) {
        m_run = tmax;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_run, tmax);
        // every PV product up to pair jp-1 must have landed in O (PV_jp cannot start before our P_jp)
        mbar_wait(&pv_done[(jp - 1) % kSBufs], ((jp - 1) / kSBufs) & 1);
        tc_fence_after();
        const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx((m_run - m_new) * c);
        l_run *= alpha;
        float o[32];
#pragma unroll 1
        for (int cc = 2 * half; cc < 2 * half + 2; ++cc) {
          tmem_ld32(tmem_base + kColO + cc * 32 + lane_off, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          tmem_st32(tmem_base + kColO + cc * 32 + lane_off, o);
        }
        tmem_wait_st();
        m_run = m_new;
      }
      const float mc = (m_run == -INFINITY) ? 0.f : m_run * c;
      uint32_t ph[kE / 2], pl[kE / 2];
      float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
      for (int i = 0; i < kE / 2; ++i) {
        float p0 = ex2_approx(fmaf(s[2 * i], c, -mc));
        float p1 = ex2_approx(fmaf(s[2 * i + 1], c, -mc));
        lsum0 += p0;
        lsum1 += p1;
        if (FMT == kFmtF16) ph[i] = pack_f16x2(p0, p1);
        else split_pack2(p0, p1, ph[i], pl[i]);
      }
      l_run += lsum0 + lsum1;
      // P hi: first kSW/2 columns of the S buffer (bf16 pairs in key order), P lo: the other kSW/2
      if (kE > 32) {
        tmem_st32_u(ts + half * (kE / 2), ph);
        if (NPASS == 3) tmem_st32_u(ts + kSW / 2 + half * (kE / 2), pl);
      } else {
        tmem_st16_u(ts + half * (kE / 2), ph);
        if (NPASS == 3) tmem_st16_u(ts + kSW / 2 + half * (kE / 2), pl);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }

    // ---- epilogue: O / l (this half's 64 columns) -> ctx ----
    xch[npairs & 1][half][r] = l_run;
    pair_bar_sync(quad);
    l_run += xch[npairs & 1][half ^ 1][r];
    if (ntiles > 0) {
      mbar_wait(&o_final, 0);
      tc_fence_after();
    }
    const float inv_l = 1.f / l_run;  // l == 0 (no unmasked key) -> inf -> NaN rows, like the reference
    const size_t orow = ((size_t)b * p.t + tq_row) * p.d + col_q;
    float o[32];
#pragma unroll 1
    for (int cc = 2 * half; cc < 2 * half + 2; ++cc) {
      if (ntiles > 0) {
        tmem_ld32(tmem_base + kColO + cc * 32 + lane_off, o);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv_l;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __int_as_float(0x7fc00000);
      }
      if (p.ctx_f32 && row_ok) {  // fp32 copy (tests / callers that want the reference's dtype)
        float4* of = reinterpret_cast<float4*>(p.ctx_f32 + orow + cc * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) of[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
      if (p.ctx_hi) {  // planes: swizzled staging in the (now idle) K ring, then TMA tensor stores
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack2(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
        uint8_t* rh = sK + cc * 16384 + r * 64;  // chunk cc: hi plane [128 rows x 64 B], lo plane 8 KB further
        uint8_t* rl = rh + 8192;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = (i ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
          *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
        }
      }
    }
    if (p.ctx_hi) {
      fence_proxy_async_smem();
      asm volatile("bar.sync 8, 256;" ::: "memory");  // all 8 softmax warps have staged their columns
      if (warp == 2 && lane == 0) {
#pragma unroll
        for (int cc = 0; cc < kDH / 32; ++cc) {
          tma_store_3d_a(&map_o_hi, sK + cc * 16384, col_q + cc * 32, q0, b);
          tma_store_3d_a(&map_o_lo, sK + cc * 16384 + 8192, col_q + cc * 32, q0, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem may be released once it has been read
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct AttnMaps {
  CUtensorMap kv_hi, kv_lo, q_hi, q_lo, o_hi, o_lo;
};

template <int NPASS, int FMT>
static int launch_attention_tc(const AttnMaps& m, const AttnTcParams& p, int batch, int nhead, cudaStream_t s) {
  static_assert(FMT == kFmtBF16 || NPASS == 1, "fp16 operands are single-plane");
  constexpr int kSmem = 2 * kKVStages * (NPASS == 3 ? 2 : 1) * kTileBytes + 1024;
  auto kern = attention_tc_kernel<NPASS, FMT>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) {
      set_error("attention_tc: cannot reserve %d bytes of shared memory", kSmem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.t + kAQ - 1) / kAQ, nhead, batch);
  kern<<<grid, kAttnTcThreads, kSmem, s>>>(m.kv_hi, m.kv_lo, m.q_hi, m.q_lo, m.o_hi, m.o_lo, p);
  LFS2_CHECK_LAUNCH("attention_tc");
  return LFS2_OK;
}

// attention_tc_pp.cu: the single-plane kernel laid out for two CTAs per SM
int launch_attention_tc_pp(const void* qkv, int f16, const uint8_t* kpm, const int* kend, void* ctx_hi, void* ctx_lo,
                           float* ctx_f32, int batch, int t, int d, int nhead, const int* row_limit, int limit_extra,
                           cudaStream_t s);

// single-plane operands, head_dim 128: LFS2_ATTN_PP = 1 (default) two CTAs per SM with thread = whole query row
// (attention_tc_pp.cu); 2 = two CTAs per SM with warp pairs (attention_tc_wide.cu, NC = 1); 0 = the one-CTA-per-SM kernel
// of this file (A/B measurements)
static int attn_pp_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LFS2_ATTN_PP");
    v = e ? atoi(e) : 1;
  }
  return v;
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" {

int lfs2_attention_tc_workspace_bytes(int batch) { return batch > 0 ? batch * (int)sizeof(int) : 0; }

int lfs2_mask_lengths(const uint8_t* pad_mask, int* lengths, int batch, int t, void* stream) {
  LFS2_REQUIRE(lengths, LFS2_ERR_INVALID_ARG, "mask_lengths: null pointer");
  if (batch == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t >= 0, LFS2_ERR_INVALID_ARG, "mask_lengths: bad shape");
  attn_kend_kernel<<<ceil_div((long long)batch * 32, 128), 128, 0, (cudaStream_t)stream>>>(pad_mask, lengths, batch, t);
  LFS2_CHECK_LAUNCH("mask_lengths");
  return LFS2_OK;
}

int lfs2_attention_tc(const void* qkv_hi, const void* qkv_lo, const uint8_t* key_padding_mask, void* ctx_hi,
                      void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t, int d, int nhead, int npass,
                      void* stream) {
  return lfs2_attention_tc_limited(qkv_hi, qkv_lo, key_padding_mask, ctx_hi, ctx_lo, ctx_f32, workspace, batch, t, d,
                                   nhead, npass, nullptr, 0, stream);
}

int lfs2_attention_tc_limited(const void* qkv_hi, const void* qkv_lo, const uint8_t* key_padding_mask, void* ctx_hi,
                              void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t, int d, int nhead,
                              int npass, const int* row_limit, int limit_extra, void* stream) {
  return lfs2_attention_tc_ex(qkv_hi, qkv_lo, LFS2_OPERAND_BF16, key_padding_mask, ctx_hi, ctx_lo, ctx_f32, workspace, batch,
                              t, d, nhead, npass, row_limit, limit_extra, stream);
}

int lfs2_attention_tc_ex(const void* qkv_hi, const void* qkv_lo, int operand_format, const uint8_t* key_padding_mask,
                         void* ctx_hi, void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t, int d, int nhead,
                         int npass, const int* row_limit, int limit_extra, void* stream) {
  LFS2_REQUIRE(qkv_hi && workspace, LFS2_ERR_INVALID_ARG, "attention_tc: null pointer");
  LFS2_REQUIRE(npass == 1 || npass == 3, LFS2_ERR_INVALID_ARG, "attention_tc: npass must be 1 or 3");
  LFS2_REQUIRE(operand_format == LFS2_OPERAND_BF16 || (operand_format == LFS2_OPERAND_F16 && npass == 1),
               LFS2_ERR_INVALID_ARG, "attention_tc: operand_format must be bf16, or fp16 with npass = 1");
  LFS2_REQUIRE(npass == 1 || qkv_lo, LFS2_ERR_INVALID_ARG, "attention_tc: npass=3 needs the lo plane");
  LFS2_REQUIRE((ctx_hi && ctx_lo) || ctx_f32, LFS2_ERR_INVALID_ARG, "attention_tc: no output");
  LFS2_REQUIRE(!ctx_hi == !ctx_lo, LFS2_ERR_INVALID_ARG, "attention_tc: ctx_hi and ctx_lo go together");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && nhead > 0 && d % nhead == 0, LFS2_ERR_INVALID_ARG,
               "attention_tc: bad shape");
  LFS2_REQUIRE(d / nhead == kDH, LFS2_ERR_UNSUPPORTED, "attention_tc: head_dim %d (only %d is implemented)",
               d / nhead, kDH);
  LFS2_REQUIRE(batch <= 65535 && nhead <= 65535, LFS2_ERR_UNSUPPORTED, "attention_tc: batch/heads exceed grid limits");
  LFS2_REQUIRE(aligned16(qkv_hi) && (!qkv_lo || aligned16(qkv_lo)) && (!ctx_hi || (aligned16(ctx_hi) && aligned16(ctx_lo))) &&
                   (!ctx_f32 || aligned16(ctx_f32)),
               LFS2_ERR_INVALID_ARG, "attention_tc: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int* kend = reinterpret_cast<int*>(workspace);
  attn_kend_kernel<<<ceil_div((long long)batch * 32, 128), 128, 0, s>>>(key_padding_mask, kend, batch, t);
  LFS2_CHECK_LAUNCH("attn_kend");

  if (npass == 1 && attn_pp_variant() == 2)
    return lfs2_attention_tc_wide(qkv_hi, operand_format, key_padding_mask, ctx_hi, ctx_lo, ctx_f32, workspace, batch, t, d,
                                  nhead, row_limit, limit_extra, stream);
  if (npass == 1 && attn_pp_variant() == 1)
    return launch_attention_tc_pp(qkv_hi, operand_format == LFS2_OPERAND_F16, key_padding_mask, kend, ctx_hi, ctx_lo,
                                  ctx_f32, batch, t, d, nhead, row_limit, limit_extra, s);

  AttnMaps m;
  bool ok = make_tmap_3d(&m.kv_hi, qkv_hi, 3ull * d, t, batch, 32, kAK, 64) &&
            make_tmap_3d(&m.q_hi, qkv_hi, 3ull * d, t, batch, 64, kAQ, 128);
  if (npass == 3)
    ok = ok && make_tmap_3d(&m.kv_lo, qkv_lo, 3ull * d, t, batch, 32, kAK, 64) &&
         make_tmap_3d(&m.q_lo, qkv_lo, 3ull * d, t, batch, 64, kAQ, 128);
  else {
    m.kv_lo = m.kv_hi;
    m.q_lo = m.q_hi;
  }
  if (ctx_hi)
    ok = ok && make_tmap_3d(&m.o_hi, ctx_hi, d, t, batch, 32, kAQ, 64) && make_tmap_3d(&m.o_lo, ctx_lo, d, t, batch, 32, kAQ, 64);
  else {
    m.o_hi = m.kv_hi;
    m.o_lo = m.kv_hi;
  }
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "attention_tc: cuTensorMapEncodeTiled failed");

  AttnTcParams p;
  p.qkv_hi = (const __nv_bfloat16*)qkv_hi;
  p.qkv_lo = (const __nv_bfloat16*)qkv_lo;
  p.kpm = key_padding_mask;
  p.kend = kend;
  p.ctx_hi = (__nv_bfloat16*)ctx_hi;
  p.ctx_lo = (__nv_bfloat16*)ctx_lo;
  p.ctx_f32 = ctx_f32;
  p.t = t;
  p.d = d;
  p.scale_log2e = (float)(1.4426950408889634 / sqrt((double)kDH));
  p.row_limit = row_limit;
  p.limit_extra = limit_extra;
  if (operand_format == LFS2_OPERAND_F16) return launch_attention_tc<1, kFmtF16>(m, p, batch, nhead, s);
  return npass == 3 ? launch_attention_tc<3, kFmtBF16>(m, p, batch, nhead, s)
                    : launch_attention_tc<1, kFmtBF16>(m, p, batch, nhead, s);
}

}  // extern "C"
