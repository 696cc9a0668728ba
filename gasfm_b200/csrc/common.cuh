// Shared device helpers for the gasfm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace gasfm {

// ---- error reporting ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> 0 / code (+ message)

#define GASFM_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::gasfm::set_error(__VA_ARGS__);      \
      return 1;                             \
    }                                       \
  } while (0)

constexpr int kNumSMs = 148;

// ---- vector memory access -------------------------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (each edge row is used once).
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// Cached 128-bit load (per-target rows are reused by every edge of the segment).
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float comp(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ float& comp(float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }

// lane mask of the LPR-lane group this lane belongs to
template <int LPR>
__device__ __forceinline__ unsigned group_mask(int lane) {
  if constexpr (LPR == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << LPR) - 1u) << (lane & ~(LPR - 1));
  }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: launchers remember the size they opted into
// per (kernel slot, current device), so that a process driving several GPUs opts in on each of them.
static inline size_t& smem_opt_in_slot(int kernel_slot) {
  static size_t allowed[4][32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return allowed[kernel_slot & 3][dev & 31];
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace gasfm
