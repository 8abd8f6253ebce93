// Shared device helpers for libtbknarpe (sm_100a).
#pragma once
#include <cuda_fp16.h>
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

// ---- fp16 range guard of the tensor-core mode. Every fp32 -> fp16 conversion of an activation that is
// not bounded by construction ([q|u], [k|v], ReLU hidden rows, the history encoder's MLP rows) saturates to +-65504
// instead of overflowing to inf (cvt.rn.satfinite: same single F2FP instruction), and the producing kernel ORs bit 0
// into a caller-owned device word when a converted value sits AT the saturation bound (tb_set_fp16_flag). The engine
// reads the word after a rollout and refuses the result instead of returning silently clamped trajectories.
extern unsigned int* tb_fp16_flag_ptr;  // host global (api.cu): device address of the sticky flag word, nullptr = off

__device__ __forceinline__ uint32_t tb_pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// running maximum of |halves| (one HMNMX2 with |.| operand modifiers); hm starts at 0
__device__ __forceinline__ void tb_track_h2(uint32_t& hm, uint32_t p) {
  const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&hm), __habs2(*reinterpret_cast<const __half2*>(&p)));
  hm = *reinterpret_cast<const uint32_t*>(&m);
}
__device__ __forceinline__ void tb_flag_if_sat(uint32_t hm, unsigned int* flag) {
  if (flag && ((hm & 0xFFFFu) == 0x7BFFu || (hm >> 16) == 0x7BFFu)) atomicOr(flag, 1u);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
