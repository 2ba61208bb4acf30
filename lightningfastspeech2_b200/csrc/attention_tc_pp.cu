// tcgen05 flash attention, head_dim 128, single-plane operands (fp16: compute mode "fp32"; bf16: "bf16" mode), laid out
// for TWO CTAs PER SM.
//
// With single-pass operands attention_tc_kernel is bound by its softmax warps, not by the tensor pipe (ncu: tensor 32 %,
// XU 33 %): the two warps that share an SM sub-partition split every row's keys, exchange the row maximum through a
// 64-thread barrier each step and therefore run load -> max -> exp -> pack -> store in lockstep, so the MUFU pipe and the
// tensor pipe each sit idle two thirds of the time.  Here a CTA is half the size -- 128 queries of one (utterance, head),
// FOUR softmax warps with thread = one whole query row (no partner, no exchange, no barrier inside a step), 256 tensor-memory
// columns, < 100 KB of shared memory -- and two of them share an SM.  They are independent, so their phases drift apart and
// one CTA's exp phase runs under the other's loads / MMAs; the hardware scheduler does the ping-pong.
// (Dispatching the utterances longest first -- a shorter grid tail on paper -- cost 5 %: CTAs of equal length then run
// in lockstep on an SM and want the same pipe at the same time.  The batch order stays as it comes.)
//
//   tensor memory (256 columns):  [0, 128)    two S/P buffers of 64 keys (S fp32, overwritten in place by P as packed 16-bit
//                                             pairs in the first 32 columns of the buffer)
//                                 [128, 256)  O accumulator, fp32
//   shared memory: Q tile (128 x 128, 32 KB, resident: S = Q.K^T is an SS-form MMA) | K ring | V ring (2 x 16 KB each)
//   warps: 0 = TMA producer (Q, K ring), 1 = MMA issuer, 2..5 = softmax (TMEM lane quadrant = warp & 3), 6 = TMA producer (V)
//   tensor-pipe order: QK_0 QK_1 | PV_0 QK_2 | PV_1 QK_3 | ...
// The normalised O rows leave as hi/lo planes through swizzled staging in the (by then idle) K ring and TMA tensor stores.
// (Straight from registers -- 16 bytes per row and instruction, half-filled sectors -- the epilogue took 9.8 k of a CTA's
// 58 k cycles: clock64 timeline, tools/attn_timeline.py, profiles/r3f_attn_timeline.txt.)
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace lfs2 {
namespace tc {

constexpr int kPQ = 128;          // queries per CTA (UMMA M)
constexpr int kPK = 64;           // keys per tile
constexpr int kPD = 128;          // head dim
constexpr int kPSlot = kPK * kPD * 2;     // 16 KB
constexpr int kPThreads = 224;
constexpr int kPKS = 2, kPVS = 2;         // ring depths (tiles)
constexpr int kPQBytes = kPQ * kPD * 2;   // 32 KB
constexpr int kPSmem = kPQBytes + (kPKS + kPVS) * kPSlot + 1024;
constexpr uint32_t kPColS = 0, kPColO = 128, kPTmemCols = 256;
static_assert(2 * (kPSmem + 2048) <= 227 * 1024, "two CTAs must fit one SM");

__device__ __forceinline__ void p_tmem_alloc(uint32_t* smem_out, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void p_tmem_ld32_nowait(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void p_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void p_tmem_st32_u(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ uint32_t p_pack16(float a, float b, int fmt) {
  uint32_t r;
  if (fmt == kFmtF16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

struct PpParams {
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

#ifdef LFS2_ATTN_TIMELINE  // diagnostics build only (tools/attn_ab.py timeline): per-CTA clock64 stamps / wait totals
__device__ long long g_attn_tl[4096][12];
#define ATL_SET(slot, val) do { const int cid_ = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x; if (cid_ < 4096) g_attn_tl[cid_][slot] = (val); } while (0)
#define ATL_T() clock64()
#else
#define ATL_SET(slot, val) do { } while (0)
#define ATL_T() 0ll
#endif

template <int FMT>
__global__ void __launch_bounds__(kPThreads, 2)
attention_tc_pp_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                       const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                       const PpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 4 boxes of [128 rows x 32 cols] (8 KB, SWIZZLE_64B)
  uint8_t* sK = sQ + kPQBytes;              // K ring
  uint8_t* sV = sK + kPKS * kPSlot;         // V ring
  __shared__ __align__(8) uint64_t k_full[kPKS], k_empty[kPKS], v_full[kPVS], v_empty[kPVS];
  __shared__ __align__(8) uint64_t q_full, s_full[2], p_full[2], pv_done[2], o_final;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kPQ, h = blockIdx.y, b = blockIdx.z;
  const long long tl_start = ATL_T();
  if (p.row_limit && q0 >= __ldg(p.row_limit + b) + p.limit_extra) return;
  const int kend = p.kend[b];
  const int ntiles = (kend + kPK - 1) / kPK;
  const int col_q = h * kPD, col_k = p.d + h * kPD, col_v = 2 * p.d + h * kPD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_q);
    for (int s = 0; s < kPKS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < kPVS; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(&o_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) p_tmem_alloc(&tmem_base_smem, kPTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer: the Q tile (resident), then the K ring =====================
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(&q_full, kPQBytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx) tma_load_3d(sQ + bx * 8192, &map_q, &q_full, col_q + bx * 32, q0, b);
#pragma unroll 1
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % kPKS;
        mbar_wait(&k_empty[st], ((j / kPKS) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], kPSlot);
        uint8_t* dk = sK + st * kPSlot;
#pragma unroll
        for (int x = 0; x < 4; ++x) tma_load_3d(dk + x * 4096, &map_kv, &k_full[st], col_k + x * 32, j * kPK, b);
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer: the V ring =====================
    if (lane == 0) {
#pragma unroll 1
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % kPVS;
        mbar_wait(&v_empty[st], ((j / kPVS) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], kPSlot);
        uint8_t* dv = sV + st * kPSlot;
#pragma unroll
        for (int x = 0; x < 4; ++x) tma_load_3d(dv + x * 4096, &map_kv, &v_full[st], col_v + x * 32, j * kPK, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ntiles > 0) {
      constexpr uint32_t idesc_qk = make_idesc(FMT, kPQ, kPK, 0, 0);  // A: Q (smem, K-major), B: K slot (K-major)
      constexpr uint32_t idesc_pv = make_idesc(FMT, kPQ, kPD, 0, 1);  // A: P (TMEM), B: V slot (MN-major)
      const uint32_t to = tmem_base + kPColO;
      const uint64_t dq0 = make_smem_desc(smem_u32(sQ), 16, 512, kSwizzle64);
      const uint64_t dk0 = make_smem_desc(smem_u32(sK), 16, 512, kSwizzle64);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 4096, 512, kSwizzle64);

      auto issue_qk = [&](int j) {  // S_j = Q . K_j^T into S/P buffer j & 1
        const int sb = j & 1, st = j % kPKS;
        mbar_wait(&k_full[st], (j / kPKS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ts = tmem_base + kPColS + sb * kPK;
          const uint64_t dk = desc_advance(dk0, st * kPSlot);
#pragma unroll
          for (int i = 0; i < kPD / 16; ++i) {
            const uint64_t a = desc_advance(dq0, (i >> 1) * 8192 + (i & 1) * 32);
            const uint64_t bd = desc_advance(dk, (i >> 1) * 4096 + (i & 1) * 32);
            if (i == 0) umma_f16_c<false>(ts, a, bd, idesc_qk);
            else umma_f16_c<true>(ts, a, bd, idesc_qk);
          }
          umma_commit(&k_empty[st]);
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
      };

      long long tl_wait_p = 0, tl_wait_v = 0;
      mbar_wait(&q_full, 0);
      if (lane == 0) ATL_SET(1, ATL_T() - tl_start);   // Q tile landed
      tc_fence_after();
      issue_qk(0);
      if (ntiles > 1) issue_qk(1);
#pragma unroll 1
      for (int j = 0; j < ntiles; ++j) {
        const int sb = j & 1, st = j % kPVS;
        const long long tl_a = ATL_T();
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        const long long tl_b = ATL_T();
        mbar_wait(&v_full[st], (j / kPVS) & 1);
        tl_wait_p += tl_b - tl_a;
        tl_wait_v += ATL_T() - tl_b;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tp = tmem_base + kPColS + sb * kPK;  // P: packed pairs in the first 32 columns of the buffer
          const uint64_t dv = desc_advance(dv0, st * kPSlot);
#pragma unroll
          for (int i = 0; i < kPK / 16; ++i) {
            if (i == 0) umma_f16_ts(to, tp, dv, idesc_pv, j ? 1u : 0u);
            else umma_f16_ts_c<true>(to, tp + i * 8, desc_advance(dv, i * 1024), idesc_pv);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&pv_done[sb]);
          if (j + 1 == ntiles) umma_commit(&o_final);
        }
        __syncwarp();
        if (j + 2 < ntiles) issue_qk(j + 2);
      }
      if (lane == 0) {
        ATL_SET(2, tl_wait_p);                         // MMA warp: cycles waiting for P (softmax)
        ATL_SET(3, tl_wait_v);                         // ... for V tiles
        ATL_SET(4, ATL_T() - tl_start);                // last PV issued
      }
    }
  } else {
    // ===================== softmax warps 2..5: thread = one whole query row =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int tq_row = q0 + r;
    const bool row_ok = tq_row < p.t;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;

    float m_run = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2e;
    const uint8_t* mrow = p.kpm ? p.kpm + (size_t)b * p.t : nullptr;
    auto key_masked = [&](int j, int sub) -> uint32_t {  // this lane's keys of tile j: j * 64 + sub * 32 + lane
      const int k1 = j * kPK + sub * 32 + lane;
      uint32_t v = 1u;
      if (k1 < p.t) v = mrow ? (uint32_t)mrow[k1] : 0u;
      return v;
    };
    uint32_t next_m0 = ntiles > 0 ? key_masked(0, 0) : 1u, next_m1 = ntiles > 0 ? key_masked(0, 1) : 1u;
    long long tl_wait_s = 0, tl_first = 0;

#pragma unroll 1
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      const uint32_t mbits0 = __ballot_sync(0xffffffffu, next_m0 != 0u);
      const uint32_t mbits1 = __ballot_sync(0xffffffffu, next_m1 != 0u);
      if (j + 1 < ntiles) {
        next_m0 = key_masked(j + 1, 0);
        next_m1 = key_masked(j + 1, 1);
      }
      const long long tl_c = ATL_T();
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tl_wait_s += ATL_T() - tl_c;
      if (j == 0) tl_first = ATL_T() - tl_start;
      tc_fence_after();
      const uint32_t ts = tmem_base + kPColS + sb * kPK + lane_off;
      float s[kPK];
      p_tmem_ld32_nowait(ts, s);
      p_tmem_ld32_nowait(ts + 32, s + 32);
      p_tmem_wait_ld();
      if (mbits0 | mbits1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if ((mbits0 >> i) & 1u) s[i] = -INFINITY;
          if ((mbits1 >> i) & 1u) s[32 + i] = -INFINITY;
        }
      }
      // row maximum: four chains of three-input maxima (FMNMX3: two new logits per issue slot)
      float t0 = fmaxf(s[0], s[4]), t1 = fmaxf(s[1], s[5]), t2 = fmaxf(s[2], s[6]), t3 = fmaxf(s[3], s[7]);
      static_assert(kPK % 8 == 0, "row maximum: the chains take eight logits per step");
#pragma unroll
      for (int i = 8; i < kPK; i += 8) {
        t0 = fmax3(t0, s[i], s[i + 4]);
        t1 = fmax3(t1, s[i + 1], s[i + 5]);
        t2 = fmax3(t2, s[i + 2], s[i + 6]);
        t3 = fmax3(t3, s[i + 3], s[i + 7]);
      }
      const float tmax = fmax3(fmaxf(t0, t1), t2, t3);
      // lazy rescale: the exponent reference only moves (and O, l are rescaled) when some row's logits outgrow it by 2^8
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
        for (int cc = 0; cc < kPD / 32; ++cc) {
          tmem_ld32(tmem_base + kPColO + cc * 32 + lane_off, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          tmem_st32(tmem_base + kPColO + cc * 32 + lane_off, o);
        }
        tmem_wait_st();
        m_run = m_new;
      }
      const float mc = (m_run == -INFINITY) ? 0.f : m_run * c;
      uint32_t ph[kPK / 2];
      // These warps are bound by issue slots and dependent-instruction latency, not by the MUFU pipe (one exponential in
      // four as a Cody-Waite + degree-4 polynomial on the FMA pipe: 1430 -> 2340 cycles per key tile).  The scale-and-
      // shift and the row sums therefore run as packed pairs (FFMA2 / FADD2: half the issue slots, the same bits).
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
      float2 l01 = make_float2(0.f, 0.f), l23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kPK / 4; ++i) {
        const float2 e01 = ffma2(make_float2(s[4 * i], s[4 * i + 1]), c2, nmc2);
        const float2 e23 = ffma2(make_float2(s[4 * i + 2], s[4 * i + 3]), c2, nmc2);
        const float p0 = ex2_approx(e01.x);
        const float p1 = ex2_approx(e01.y);
        const float p2 = ex2_approx(e23.x);
        const float p3 = ex2_approx(e23.y);
        l01 = fadd2(l01, make_float2(p0, p1));
        l23 = fadd2(l23, make_float2(p2, p3));
        ph[2 * i] = p_pack16(p0, p1, FMT);
        ph[2 * i + 1] = p_pack16(p2, p3, FMT);
      }
      l_run += (l01.x + l01.y) + (l23.x + l23.y);
      p_tmem_st32_u(ts, ph);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }

    // ---- epilogue: O / l -> ctx (hi/lo planes: staged, TMA stores; the optional fp32 copy straight from registers) ----
    const long long tl_loop_end = ATL_T() - tl_start;
    if (ntiles > 0) {
      mbar_wait(&o_final, 0);
      tc_fence_after();
    }
    const long long tl_ofinal = ATL_T() - tl_start;
    const float inv_l = 1.f / l_run;  // l == 0 (no unmasked key) -> inf -> NaN rows, like the reference
    const size_t orow = ((size_t)b * p.t + tq_row) * p.d + col_q;
    float o[32];
#pragma unroll 1
    for (int cc = 0; cc < kPD / 32; ++cc) {
      if (ntiles > 0) {
        tmem_ld32(tmem_base + kPColO + cc * 32 + lane_off, o);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 v = fmul2(make_float2(o[i], o[i + 1]), make_float2(inv_l, inv_l));
          o[i] = v.x;
          o[i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __int_as_float(0x7fc00000);
      }
      if (p.ctx_hi) {
        // hi | lo planes of this 128 x 32 chunk -> swizzled staging (two 16 KB buffers in the idle K ring) -> TMA stores;
        // rows past the end of the utterance tensor are clipped by the store
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack2(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
        uint8_t* sb = sK + (cc & 1) * kPSlot;
        const bool st_issuer = warp == 2 && lane == 0;
        if (cc >= 2) {  // the buffer's previous store has finished reading it
          if (st_issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        uint8_t* rh = sb + r * 64;
        uint8_t* rl = rh + kPSlot / 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = (i ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(rh + u) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
          *reinterpret_cast<uint4*>(rl + u) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (st_issuer) {
          const uint64_t mh = reinterpret_cast<uint64_t>(&map_o_hi), ml = reinterpret_cast<uint64_t>(&map_o_lo);
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(mh),
                       "r"(smem_u32(sb)), "r"(col_q + cc * 32), "r"(q0), "r"(b)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(ml),
                       "r"(smem_u32(sb + kPSlot / 2)), "r"(col_q + cc * 32), "r"(q0), "r"(b)
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
    if (p.ctx_hi && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef LFS2_ATTN_TIMELINE
    if (warp == 2 && lane == 0) {
      ATL_SET(0, (long long)ntiles);
      ATL_SET(5, tl_first);        // softmax: first S tile seen
      ATL_SET(6, tl_wait_s);       // softmax: cycles waiting for S
      ATL_SET(7, tl_loop_end);     // softmax: last P written
      ATL_SET(8, tl_ofinal);       // O complete
      ATL_SET(9, ATL_T() - tl_start);   // O rows stored
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      ATL_SET(10, (long long)smid);
      ATL_SET(11, tl_start);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPTmemCols);
  }
}

#ifdef LFS2_ATTN_TIMELINE
extern "C" __attribute__((visibility("default"))) int lfs2_attn_timeline(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_attn_tl, sizeof(g_attn_tl)) == cudaSuccess ? 0 : 1;
}
#endif

template <int FMT>
static int launch_pp(const CUtensorMap& kv, const CUtensorMap& q, const CUtensorMap& oh, const CUtensorMap& ol,
                     const PpParams& p, int batch, int nhead, cudaStream_t s) {
  auto kern = attention_tc_pp_kernel<FMT>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem) != cudaSuccess) {
      set_error("attention_tc_pp: cannot reserve %d bytes of shared memory", kPSmem);
      return LFS2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.t + kPQ - 1) / kPQ, nhead, batch);
  kern<<<grid, kPThreads, kPSmem, s>>>(kv, q, oh, ol, p);
  LFS2_CHECK_LAUNCH("attention_tc_pp");
  return LFS2_OK;
}

// single-plane attention launch used by lfs2_attention_tc_ex (attention_tc.cu); kend = workspace already filled
int launch_attention_tc_pp(const void* qkv, int f16, const uint8_t* kpm, const int* kend, void* ctx_hi, void* ctx_lo,
                           float* ctx_f32, int batch, int t, int d, int nhead, const int* row_limit, int limit_extra,
                           cudaStream_t s) {
  CUtensorMap kv, q, oh, ol;
  bool ok = make_tmap_3d(&kv, qkv, 3ull * d, t, batch, 32, kPK, 64) && make_tmap_3d(&q, qkv, 3ull * d, t, batch, 32, kPQ, 64);
  oh = kv;
  ol = kv;
  if (ctx_hi) ok = ok && make_tmap_3d(&oh, ctx_hi, d, t, batch, 32, kPQ, 64) && make_tmap_3d(&ol, ctx_lo, d, t, batch, 32, kPQ, 64);
  LFS2_REQUIRE(ok, LFS2_ERR_CUDA, "attention_tc_pp: cuTensorMapEncodeTiled failed");
  PpParams p;
  p.kpm = kpm;
  p.kend = kend;
  p.ctx_hi = (__nv_bfloat16*)ctx_hi;
  p.ctx_lo = (__nv_bfloat16*)ctx_lo;
  p.ctx_f32 = ctx_f32;
  p.t = t;
  p.d = d;
  p.scale_log2e = (float)(1.4426950408889634 / sqrt((double)kPD));
  p.row_limit = row_limit;
  p.limit_extra = limit_extra;
  return f16 ? launch_pp<kFmtF16>(kv, q, oh, ol, p, batch, nhead, s) : launch_pp<kFmtBF16>(kv, q, oh, ol, p, batch, nhead, s);
}

}  // namespace tc
}  // namespace lfs2
