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

constexpr int kNumSMs = 148;  // B200

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

}  // namespace lfs2
