// Shared device helpers for the DGDM B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include "../../include/dgdm_b200.h"

#define DGDM_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return DGDM_ERR_LAUNCH; } while (0)

namespace dgdm {

constexpr int kWarp = 32;

// cudaFuncSetAttribute is per device / context: remember which devices of this process have been configured (bit per ordinal).
// Racing callers at worst repeat the (idempotent) attribute call.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0ull};
  bool needed() const { int d = 0; cudaGetDevice(&d); return !((done.load(std::memory_order_acquire) >> (d & 63)) & 1ull); }
  void mark() { int d = 0; cudaGetDevice(&d); done.fetch_or(1ull << (d & 63), std::memory_order_release); }
};

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
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf-form GELU (torch.nn.GELU() default) with a branch-free erf: erfc(t) = 2^-Q(t), Q(t) = t * poly7(t) fitted (weighted minimax)
// on [0, 4] — erf(4) rounds to 1 in fp32 — so erf(x) = sign(x) * (1 - ex2(-Q(|x|))). Max abs error of erf 9.1e-8 in fp32
// arithmetic (+ <= 2.4e-7 from ex2.approx), the same class as erff's ~2 ulp, at ~15 instructions instead of ~45 (two divergent
// branches): the exact-erf GELU was the dominant cost of the LayerNorm+GELU GEMM epilogues.
__device__ __forceinline__ float erf_fast(float x) {
  const float t = fminf(fabsf(x), 4.0f);
  float p = 4.53585889772512e-05f;
  p = fmaf(p, t, -0.00044550737948156893f);
  p = fmaf(p, t, 0.0014894403284415603f);
  p = fmaf(p, t, 0.0007746326155029237f);
  p = fmaf(p, t, -0.02825368382036686f);
  p = fmaf(p, t, 0.14848162233829498f);
  p = fmaf(p, t, 0.9184163808822632f);
  p = fmaf(p, t, 1.6279085874557495f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-p * t));
  return copysignf(1.0f - e, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_fast(x * 0.70710678118654752440f), hx);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case DGDM_ACT_GELU: return gelu_erf(x);
    case DGDM_ACT_RELU: return fmaxf(x, 0.0f);
    case DGDM_ACT_SILU: return silu(x);
    case DGDM_ACT_TANH: return tanhf(x);
    case DGDM_ACT_ELU:  return x > 0.0f ? x : expm1f(x);
    default: return x;
  }
}

// Activation over a register chunk with the switch hoisted OUT of the element loop: a per-element runtime switch
// compiles to an indirect branch (BRX) per element, which dominated the GEMM epilogue (60 us for a single CTA).
template <int ACT>
__device__ __forceinline__ float act_const(float x) {
  if constexpr (ACT == DGDM_ACT_GELU) return gelu_erf(x);
  else if constexpr (ACT == DGDM_ACT_RELU) return fmaxf(x, 0.0f);
  else if constexpr (ACT == DGDM_ACT_SILU) return silu(x);
  else if constexpr (ACT == DGDM_ACT_TANH) return tanhf(x);
  else if constexpr (ACT == DGDM_ACT_ELU) return x > 0.0f ? x : expm1f(x);
  else return x;
}
template <int N>
__device__ __forceinline__ void apply_act_chunk(float* v, int act) {
  switch (act) {
    case DGDM_ACT_GELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_const<DGDM_ACT_GELU>(v[j]);
      break;
    case DGDM_ACT_RELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_const<DGDM_ACT_RELU>(v[j]);
      break;
    case DGDM_ACT_SILU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_const<DGDM_ACT_SILU>(v[j]);
      break;
    case DGDM_ACT_TANH:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_const<DGDM_ACT_TANH>(v[j]);
      break;
    case DGDM_ACT_ELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_const<DGDM_ACT_ELU>(v[j]);
      break;
    default: break;
  }
}
__device__ __forceinline__ void apply_act4(float4& a, int act) {
  float v[4] = {a.x, a.y, a.z, a.w};
  apply_act_chunk<4>(v, act);
  a = make_float4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// order-preserving float <-> uint mapping (for atomic min/max and radix select)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace dgdm
