// Shared helpers for libynet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ynet_b200.h"

namespace ynet {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();

#define YNET_CHECK_ARG(cond, msg)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      ::ynet::set_error("%s: %s", __func__, msg);       \
      return YNET_E_INVALID;                            \
    }                                                   \
  } while (0)

#define YNET_CHECK_ALIGN(ptr, a)                                     \
  do {                                                               \
    if ((reinterpret_cast<uintptr_t>(ptr) % (a)) != 0) {             \
      ::ynet::set_error("%s: %s not %d-byte aligned", __func__, #ptr, (int)(a)); \
      return YNET_E_ALIGN;                                           \
    }                                                                \
  } while (0)

#define YNET_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return ::ynet::cuda_fail(e__, __func__); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

template <typename T>
__host__ __device__ constexpr T tmin(T a, T b) {
  return a < b ? a : b;
}
template <typename T>
__host__ __device__ constexpr T tmax(T a, T b) {
  return a > b ? a : b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace ynet
