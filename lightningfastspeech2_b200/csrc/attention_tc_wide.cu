// tcgen05 flash attention for wide heads: head_dim = 128 * NC (NC = 2, 3: the 76 M configuration's head_dim 384),
// operands as ONE 16-bit plane (bf16: compute mode "bf16"; fp16: compute mode "fp32", see attention_tc.cu).
//
// Same structure as attention_tc.cu (one CTA = 128 queries of one (utterance, head); 64-key tiles stream through
// shared memory by TMA; online softmax with a lazily moved exponent reference; no T x T tensor ever reaches HBM), with
// the budget re-cut for the wide head:
//   tensor memory (512 columns):  [0, 128)   two S/P buffers of 64 keys (S fp32, overwritten in place by P as packed
//                                            16-bit pairs in the first 32 columns of the buffer)
//                                 [128, 128 + 128 NC)   O accumulator, fp32
//   -> no room for Q in tensor memory: Q (128 x 128 NC, 32 KB per 128 columns) stays in SHARED memory for the whole CTA
//      and S = Q.K^T is an SS-form MMA accumulated over the NC 128-column chunks of the head.
//   shared memory: Q | K ring | V ring, both rings in 16 KB slots = [64 keys x 128 columns]; a key tile is NC slots
//      of K (consumed along the contraction of Q.K^T) and NC slots of V (the NC 128-column blocks of O += P.V).
//   the normalised O rows leave straight from registers (thread = row, 64-byte pieces per plane): with 4 NC chunks per
//   row there is no shared memory left to stage them.
// Warp roles as in attention_tc.cu: 0 = TMA producer (Q, K ring), 10 = TMA producer (V ring), 1 = MMA issuer,
// 2..9 = softmax (thread = query row; the two warps of a TMEM lane quadrant split a tile's 64 keys and the O columns).
#include <math.h>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kWQ = 128;          // queries per CTA (UMMA M)
constexpr int kWK = 64;           // keys per tile
constexpr int kWC = 128;          // head-dim columns per chunk / ring slot
constexpr int kWSlot = kWK * kWC * 2;     // 16 KB
constexpr int kWThreads = 352;
constexpr int kWE = kWK / 2;      // keys per softmax thread and step

// CTAS = CTAs that share an SM (1, or 2 for head_dim 128: half the tensor memory and shared memory each, so that two
// independent CTAs de-phase each other's softmax / MMA phases, see attention_tc_pp.cu)
template <int NC, int CTAS = 1>
struct WideCfg {
  static constexpr int kQBytes = NC * kWQ * kWC * 2;                          // 32 KB per chunk
  // 227 KB per SM minus the static barriers / exchange buffers (~2.4 KB per CTA) and the 1 KB alignment slack
  static constexpr int kBudget = (CTAS == 2 ? 110 : 222) * 1024;
  static constexpr int kRing = (kBudget - kQBytes) / kWSlot;                  // slots available to both rings
  static constexpr int kKS = kRing / 2 + (kRing & 1), kVS = kRing / 2;        // K ring, V ring
  static constexpr int kSmem = kQBytes + (kKS + kVS) * kWSlot + 1024;
  static constexpr uint32_t kColS = 0, kColO = 128;
  static constexpr uint32_t kTmemCols = CTAS == 2 ? 256 : 512;
  static_assert(kColO + NC * kWC <= kTmemCols, "tensor memory budget");
  static_assert(kKS >= 2 && kVS >= 2, "rings need at least two slots");
  static_assert(CTAS * (kSmem + 4096) <= 227 * 1024 + (CTAS - 1) * 2048, "shared memory budget");
};
constexpr int kWMaxSlots = 8;
static_assert(WideCfg<2>::kKS <= kWMaxSlots && WideCfg<3>::kKS <= kWMaxSlots && WideCfg<1, 2>::kKS <= kWMaxSlots,
              "barrier arrays");
__device__ __forceinline__ void w_tmem_alloc(uint32_t* smem_out, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void w_tmem_ld32_nowait(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void w_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void w_tmem_st16_u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void w_pair_bar_sync(int quad) { asm volatile("bar.sync %0, 64;" ::"r"(4 + quad) : "memory"); }
__device__ __forceinline__ uint32_t w_pack16(float a, float b, int fmt) {
  uint32_t r;
  if (fmt == kFmtF16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

struct WideParams {
  const uint8_t* kpm;   // (B, T) 1 = PAD, or null
  const int* kend;      // (B): 1 + index of the last non-PAD key
  __nv_bfloat16* ctx_hi;
  __nv_bfloat16* ctx_lo;
  float* ctx_f32;
  int t, d;
  float scale_log2e;
  const int* row_limit;
  int limit_extra;
};

__global__ void wide_kend_kernel(const uint8_t* __restrict__ kpm, int* __restrict__ kend, int batch, int t) {
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

template <int NC, int FMT, int CTAS>
__global__ void __launch_bounds__(kWThreads, CTAS)
attention_tc_wide_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                         const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                         const WideParams p) {
  using Cfg = WideCfg<NC, CTAS>;
  constexpr int kKS = Cfg::kKS, kVS = Cfg::kVS;
  constexpr uint32_t kColS = Cfg::kColS, kColO = Cfg::kColO;
  constexpr int kDH = NC * kWC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // NC * 4 boxes of [128 rows x 32 cols] (8 KB, SWIZZLE_64B)
  uint8_t* sK = sQ + Cfg::kQBytes;          // K ring
  uint8_t* sV = sK + kKS * kWSlot;          // V ring
  __shared__ __align__(8) uint64_t k_full[kWMaxSlots], k_empty[kWMaxSlots], v_full[kWMaxSlots], v_empty[kWMaxSlots];
  __shared__ __align__(8) uint64_t q_full, s_full[2], p_full[2], pv_done[2], o_final;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][2][kWQ];  // [tile parity][half][row]: row maxima / partial sums between the warp pair

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kWQ, h = blockIdx.y, b = blockIdx.z;
  if (p.row_limit && q0 >= __ldg(p.row_limit + b) + p.limit_extra) return;
  const int kend = p.kend[b];
  const int ntiles = (kend + kWK - 1) / kWK;
  const int col_q = h * kDH, col_k = p.d + h * kDH, col_v = 2 * p.d + h * kDH;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_q);
    for (int s = 0; s < kKS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < kVS; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(&o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) w_tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer: the Q tile (resident), then the K ring =====================
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(&q_full, Cfg::kQBytes);
#pragma unroll 1
      for (int bx = 0; bx < NC * 4; ++bx) tma_load_3d(sQ + bx * 8192, &map_q, &q_full, col_q + bx * 32, q0, b);
#pragma unroll 1
      for (int n = 0; n < ntiles * NC; ++n) {
        const int j = n / NC, c = n % NC, st = n % kKS;
        mbar_wait(&k_empty[st], ((n / kKS) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], kWSlot);
        uint8_t* dk = sK + st * kWSlot;
#pragma unroll
        for (int x = 0; x < 4; ++x) tma_load_3d(dk + x * 4096, &map_kv, &k_full[st], col_k + c * kWC + x * 32, j * kWK, b);
      }
    }
  } else if (warp == 10) {
    // ===================== TMA producer: the V ring =====================
    if (lane == 0) {
#pragma unroll 1
      for (int n = 0; n < ntiles * NC; ++n) {
        const int j = n / NC, c = n % NC, st = n % kVS;
        mbar_wait(&v_empty[st], ((n / kVS) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], kWSlot);
        uint8_t* dv = sV + st * kWSlot;
#pragma unroll
        for (int x = 0; x < 4; ++x) tma_load_3d(dv + x * 4096, &map_kv, &v_full[st], col_v + c * kWC + x * 32, j * kWK, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ntiles > 0) {
      constexpr uint32_t idesc_qk = make_idesc(FMT, kWQ, kWK, 0, 0);  // A: Q (smem, K-major), B: K slot (K-major)
      constexpr uint32_t idesc_pv = make_idesc(FMT, kWQ, kWC, 0, 1);  // A: P (TMEM), B: V slot (MN-major)
      const uint32_t to = tmem_base + kColO;
      const uint64_t dq0 = make_smem_desc(smem_u32(sQ), 16, 512, kSwizzle64);
      const uint64_t dk0 = make_smem_desc(smem_u32(sK), 16, 512, kSwizzle64);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 4096, 512, kSwizzle64);

      auto issue_qk = [&](int j) {  // S_j = Q . K_j^T into S/P buffer j & 1, accumulated over the NC head-dim chunks
        const int sb = j & 1;
        const uint32_t ts = tmem_base + kColS + sb * kWK;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          const int n = j * NC + c, st = n % kKS;
          mbar_wait(&k_full[st], (n / kKS) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dk = desc_advance(dk0, st * kWSlot);
            const uint64_t dq = desc_advance(dq0, c * 4 * 8192);
#pragma unroll
            for (int i = 0; i < kWC / 16; ++i) {
              const uint64_t a = desc_advance(dq, (i >> 1) * 8192 + (i & 1) * 32);
              const uint64_t bd = desc_advance(dk, (i >> 1) * 4096 + (i & 1) * 32);
              umma_f16(ts, a, bd, idesc_qk, (c | i) ? 1u : 0u);
            }
            umma_commit(&k_empty[st]);
            if (c == NC - 1) umma_commit(&s_full[sb]);
          }
          __syncwarp();
        }
      };

      mbar_wait(&q_full, 0);
      tc_fence_after();
      issue_qk(0);
      if (ntiles > 1) issue_qk(1);
#pragma unroll 1
      for (int j = 0; j < ntiles; ++j) {
        const int sb = j & 1;
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        const uint32_t tp = tmem_base + kColS + sb * kWK;  // P: packed pairs in the first 32 columns of the buffer
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          const int n = j * NC + c, st = n % kVS;
          mbar_wait(&v_full[st], (n / kVS) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dv = desc_advance(dv0, st * kWSlot);
#pragma unroll
            for (int i = 0; i < kWK / 16; ++i)
              umma_f16_ts(to + c * kWC, tp + i * 8, desc_advance(dv, i * 1024), idesc_pv, (j | i) ? 1u : 0u);
            umma_commit(&v_empty[st]);
            if (c == NC - 1) {
              umma_commit(&pv_done[sb]);
              if (j + 1 == ntiles) umma_commit(&o_final);
            }
          }
          __syncwarp();
        }
        if (j + 2 < ntiles) issue_qk(j + 2);
      }
    }
  } else {
    // ===================== softmax warps 2..9: thread = query row =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int tq_row = q0 + r;
    const bool row_ok = tq_row < p.t;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    constexpr int kChunksPerHalf = 2 * NC;  // 32-column chunks of O per warp of the pair

    float m_run = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2e;
    const uint8_t* mrow = p.kpm ? p.kpm + (size_t)b * p.t : nullptr;
    auto key_masked = [&](int j) -> uint32_t {  // this lane's key of tile j: j * 64 + half * 32 + lane
      const int k1 = j * kWK + half * kWE + lane;
      uint32_t v = 1u;
      if (k1 < p.t) v = mrow ? (uint32_t)mrow[k1] : 0u;
      return v;
    };
    uint32_t next_m = ntiles > 0 ? key_masked(0) : 1u;

#pragma unroll 1
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      const uint32_t mbits = __ballot_sync(0xffffffffu, next_m != 0u);
      if (j + 1 < ntiles) next_m = key_masked(j + 1);
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_base + kColS + sb * kWK + lane_off;
      float s[kWE];
      w_tmem_ld32_nowait(ts + half * kWE, s);
      w_tmem_wait_ld();
      if (mbits) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if ((mbits >> i) & 1u) s[i] = -INFINITY;
      }
      float tmax0 = s[0], tmax1 = s[1];
#pragma unroll
      for (int i = 2; i < kWE; i += 2) {
        tmax0 = fmaxf(tmax0, s[i]);
        tmax1 = fmaxf(tmax1, s[i + 1]);
      }
      float tmax = fmaxf(tmax0, tmax1);
      xch[j & 1][half][r] = tmax;
      w_pair_bar_sync(quad);  // also: both halves have read S before either overwrites it with P
      tmax = fmaxf(tmax, xch[j & 1][half ^ 1][r]);
      const bool grow = (j > 0) && ((tmax - m_run) * c > 8.f);
      if (j == 0) {
        m_run = tmax;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_run, tmax);
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // every PV product up to tile j-1 has landed in O
        tc_fence_after();
        const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx((m_run - m_new) * c);
        l_run *= alpha;
        float o[32];
#pragma unroll 1
        for (int cc = half * kChunksPerHalf; cc < (half + 1) * kChunksPerHalf; ++cc) {
          tmem_ld32(tmem_base + kColO + cc * 32 + lane_off, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          tmem_st32(tmem_base + kColO + cc * 32 + lane_off, o);
        }
        tmem_wait_st();
        m_run = m_new;
      }
      const float mc = (m_run == -INFINITY) ? 0.f : m_run * c;
      uint32_t ph[kWE / 2];
      // scale-and-shift and the row sums as packed pairs (FFMA2 / FADD2: half the issue slots, the same bits)
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
      float2 lsum = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kWE / 2; ++i) {
        const float2 e = ffma2(make_float2(s[2 * i], s[2 * i + 1]), c2, nmc2);
        const float p0 = ex2_approx(e.x);
        const float p1 = ex2_approx(e.y);
        lsum = fadd2(lsum, make_float2(p0, p1));
        ph[i] = w_pack16(p0, p1, FMT);
      }
      l_run += lsum.x + lsum.y;
      w_tmem_st16_u(ts + half * (kWE / 2), ph);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }

    // ---- epilogue: O / l (this half's columns) -> ctx.  The hi/lo planes leave through swizzled staging in the (by then
    //      idle) Q tile and TMA stores, one or two 16 KB buffers per half (16-byte register stores half-fill their sectors:
    //      in attention_tc_pp.cu they were 17 % of a CTA's life); the optional fp32 copy goes straight from registers ----
    xch[ntiles & 1][half][r] = l_run;
    w_pair_bar_sync(quad);
    l_run += xch[ntiles & 1][half ^ 1][r];
    if (ntiles > 0) {
      mbar_wait(&o_final, 0);
      tc_fence_after();
    }
    const float inv_l = 1.f / l_run;  // l == 0 (no unmasked key) -> inf -> NaN rows, like the reference
    const size_t orow = ((size_t)b * p.t + tq_row) * p.d + col_q;
    float o[32];
    constexpr int kStBufs = NC >= 2 ? 2 : 1;             // staging buffers per half inside the Q tile (NC * 32 KB)
    const bool st_issuer = quad == 0 && lane == 0;        // one thread per half issues / retires its stores
    int st_k = 0;
#pragma unroll 1
    for (int cc = half * kChunksPerHalf; cc < (half + 1) * kChunksPerHalf; ++cc, ++st_k) {
      if (ntiles > 0) {
        tmem_ld32(tmem_base + kColO + cc * 32 + lane_off, o);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv_l;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __int_as_float(0x7fc00000);
      }
      if (p.ctx_hi) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack2(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
        uint8_t* sb = sQ + (half * kStBufs + (st_k % kStBufs)) * 16384;
        if (st_k >= kStBufs) {  // the buffer's previous store has finished reading it
          if (st_issuer) {
            if (kStBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          asm volatile("bar.sync %0, 128;" ::"r"(8 + half) : "memory");
        }
        uint8_t* rh = sb + r * 64;
        uint8_t* rl = rh + 8192;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = (i ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
          *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(8 + half) : "memory");
        if (st_issuer) {
          const uint64_t mh = reinterpret_cast<uint64_t>(&map_o_hi), ml = reinterpret_cast<uint64_t>(&map_o_lo);
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(mh),
                       "r"(smem_u32(sb)), "r"(col_q + cc * 32), "r"(q0), "r"(b)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(ml),
                       "r"(smem_u32(sb + 8192)), "r"(col_q + cc * 32), "r"(q0), "r"(b)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (!row_ok) continue;
      if (p.ctx_f32) {
        float4* of = reinterpret_cast<float4*>(p.ctx_f32 + orow + cc * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) of[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    }
    if (p.ctx_hi && st_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int NC, int FMT, int CTAS = 1>
static int launch_wide(const CUtensorMap& kv, const CUtensorMap& q, const CUtensorMap& oh, const CUtensorMap& ol,
                       const WideParams& p, int batch, int nhead, cudaStream_t s) {
  constexpr int kSmem = WideCfg<NC, CTAS>::kSmem;
  auto kern = attention_tc_wide_kernel<NC, FMT, CTAS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) {
      set_error("attention_tc_wide: cannot reserve %d bytes of shared memory", kSmem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.t + kWQ - 1) / kWQ, nhead, batch);
  kern<<<grid, kWThreads, kSmem, s>>>(kv, q, oh, ol, p);
  LFS2_CHECK_LAUNCH("attention_tc_wide");
  return LFS2_OK;
}

}  // namespace tc
}  // namespace lfs2

using namespace lfs2;
using namespace lfs2::tc;

extern "C" int lfs2_attention_tc_wide(const void* qkv, int operand_format, const uint8_t* key_padding_mask, void* ctx_hi,
                                      void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t, int d, int nhead,
                                      const int* row_limit, int limit_extra, void* stream) {
  LFS2_REQUIRE(qkv && workspace, LFS2_ERR_INVALID_ARG, "attention_tc_wide: null pointer");
  LFS2_REQUIRE(operand_format == LFS2_OPERAND_BF16 || operand_format == LFS2_OPERAND_F16, LFS2_ERR_INVALID_ARG,
               "attention_tc_wide: operand_format must be LFS2_OPERAND_BF16 or LFS2_OPERAND_F16");
  LFS2_REQUIRE((ctx_hi && ctx_lo) || ctx_f32, LFS2_ERR_INVALID_ARG, "attention_tc_wide: no output");
  LFS2_REQUIRE(!ctx_hi == !ctx_lo, LFS2_ERR_INVALID_ARG, "attention_tc_wide: ctx_hi and ctx_lo go together");
  if (batch == 0 || t == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && t > 0 && d > 0 && nhead > 0 && d % nhead == 0, LFS2_ERR_INVALID_ARG,
               "attention_tc_wide: bad shape");
  const int dh = d / nhead;
  LFS2_REQUIRE(dh == 128 || dh == 256 || dh == 384, LFS2_ERR_UNSUPPORTED,
               "attention_tc_wide: head_dim %d (128, 256 and 384 are implemented)", dh);
  LFS2_REQUIRE(batch <= 65535 && nhead <= 65535, LFS2_ERR_UNSUPPORTED, "attention_tc_wide: batch/heads exceed grid limits");
  LFS2_REQUIRE(aligned16(qkv) && (!ctx_hi || (aligned16(ctx_hi) && aligned16(ctx_lo))) && (!ctx_f32 || aligned16(ctx_f32)),
               LFS2_ERR_INVALID_ARG, "attention_tc_wide: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int* kend = reinterpret_cast<int*>(workspace);
  wide_kend_kernel<<<ceil_div((long long)batch * 32, 128), 128, 0, s>>>(key_padding_mask, kend, batch, t);
  LFS2_CHECK_LAUNCH("attn_kend");
  CUtensorMap kv, q, oh, ol;
  bool ok = make_tmap_3d(&kv, qkv, 3ull * d, t, batch, 32, kWK, 64) && make_tmap_3d(&q, qkv, 3ull * d, t, batch, 32, kWQ, 64);
  oh = kv;
  ol = kv;
  if (ctx_hi) ok = ok && make_tmap_3d(&oh, ctx_hi, d, t, batch, 32, kWQ, 64) && make_tmap_3d(&ol, ctx_lo, d, t, batch, 32, kWQ, 64);
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "attention_tc_wide: cuTensorMapEncodeTiled failed");
  WideParams p;
  p.kpm = key_padding_mask;
  p.kend = kend;
  p.ctx_hi = (__nv_bfloat16*)ctx_hi;
  p.ctx_lo = (__nv_bfloat16*)ctx_lo;
  p.ctx_f32 = ctx_f32;
  p.t = t;
  p.d = d;
  p.scale_log2e = (float)(1.4426950408889634 / sqrt((double)dh));
  p.row_limit = row_limit;
  p.limit_extra = limit_extra;
  const bool f16 = operand_format == LFS2_OPERAND_F16;
  if (dh == 128)  // two CTAs per SM, warp pairs (variant 2 of lfs2_attention_tc_ex's single-plane kernels)
    return f16 ? launch_wide<1, kFmtF16, 2>(kv, q, oh, ol, p, batch, nhead, s) : launch_wide<1, kFmtBF16, 2>(kv, q, oh, ol, p, batch, nhead, s);
  if (dh == 384)
    return f16 ? launch_wide<3, kFmtF16>(kv, q, oh, ol, p, batch, nhead, s) : launch_wide<3, kFmtBF16>(kv, q, oh, ol, p, batch, nhead, s);
  return f16 ? launch_wide<2, kFmtF16>(kv, q, oh, ol, p, batch, nhead, s) : launch_wide<2, kFmtBF16>(kv, q, oh, ol, p, batch, nhead, s);
}
