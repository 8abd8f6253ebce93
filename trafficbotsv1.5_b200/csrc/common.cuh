// Shared device helpers for libtbknarpe (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tb_knarpe.h"

#define TB_FULL_MASK 0xffffffffu

#define TB_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return TB_ERR_CUDA;             \
  } while (0)

static inline bool tb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Accurate-enough sin/cos for |a| up to a few thousand rad: two-constant Cody-Waite reduction to [-pi, pi]
// followed by the SFU approximations (abs error 2^-21.4 on that interval). The reference never wraps angles
// (utils/rpe.py:31 cast=False), so arguments like x * f (|x| <= 500 m, f <= 1) must be reduced here.
__device__ __forceinline__ float tb_reduce_2pi(float a) {
  const float inv2pi = 0.15915494309189535f;
  const float c1 = 6.28318548202514648f;    // fp32(2*pi)
  const float c2 = -1.74845553146e-07f;     // 2*pi - c1
  float k = rintf(a * inv2pi);
  float r = fmaf(k, -c1, a);
  return fmaf(k, -c2, r);
}
__device__ __forceinline__ void tb_sincos(float a, float* s, float* c) {
  float r = tb_reduce_2pi(a);
  *s = __sinf(r);
  *c = __cosf(r);
}
__device__ __forceinline__ float tb_cos(float a) { return __cosf(tb_reduce_2pi(a)); }
__device__ __forceinline__ float tb_sin(float a) { return __sinf(tb_reduce_2pi(a)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
