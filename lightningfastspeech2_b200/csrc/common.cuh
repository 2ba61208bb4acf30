// Shared helpers for the lfs2 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lfs2.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "lfs2 kernels are written for sm_100a (Blackwell B200) only"
#endif

namespace lfs2 {

void set_error(const char* fmt, ...);

#define LFS2_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::lfs2::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

#define LFS2_CHECK_LAUNCH(name)                                                      \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::lfs2::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return LFS2_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// SM count of the current device (grids of the persistent kernels, split heuristics): queried once per process;
// 148 on the B200 this library is written for
static inline int num_sms() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        v <= 0)
      return 148;
    return v;
  }();
  return n;
}

// acc += w * x on all four lanes of a float4 as two packed fp32 FMAs (sm_100 FFMA2: two IEEE fused multiply-adds per
// issue slot, bit-identical to four fmaf).  The stencil loops of the depthwise conv are issue-bound, not FMA-pipe-bound.
__device__ __forceinline__ void fma4(float4& acc, const float4& w, const float4& x) {
  unsigned long long a0, a1, w0, w1, x0, x1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(acc.z), "f"(acc.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w0) : "f"(w.x), "f"(w.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w1) : "f"(w.z), "f"(w.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(w0), "l"(x0));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(w1), "l"(x1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a1));
}

// packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): two IEEE operations per issue slot, bit-identical to the scalar forms
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ua, ub, uc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(uc) : "l"(ua), "l"(ub));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(uc));
  return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ua, ub;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ua) : "l"(ub));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ua));
  return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ua, ub;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(ua) : "l"(ub));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ua));
  return r;
}

// max(a, b, c) in one instruction (sm_100 FMNMX3)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 16-byte accesses (read-once / write-once data: keep them out of L1)
__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- counter-based dropout masks (dropout.cu and the kernels that fuse a dropout site) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep-scale factors of elements 4i .. 4i+3
__device__ __forceinline__ float4 dropout_scale4(size_t i, uint32_t threshold, float inv_keep, uint2 key, uint32_t site) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), site, 0u), key);
  return make_float4(r.x >= threshold ? inv_keep : 0.f, r.y >= threshold ? inv_keep : 0.f,
                     r.z >= threshold ? inv_keep : 0.f, r.w >= threshold ? inv_keep : 0.f);
}

// (p, seed, site) of one dropout site as kernel arguments; threshold == 0 and inv_keep == 1 mean "no dropout"
struct DropSite {
  uint32_t threshold;
  float inv_keep;
  uint2 key;
  uint32_t site;
};
static inline DropSite make_drop_site(float p, unsigned long long seed, unsigned int site) {
  DropSite d;
  double t = (double)p * 4294967296.0;
  d.threshold = p <= 0.f ? 0u : (t >= 4294967295.0 ? 4294967295u : (uint32_t)t);
  d.inv_keep = p <= 0.f ? 1.f : 1.f / (1.f - p);
  d.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  d.site = site;
  return d;
}

}  // namespace lfs2
