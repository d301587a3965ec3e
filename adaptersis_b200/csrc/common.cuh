// common.cuh -- shared helpers for libasis_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/asis_b200.h"

namespace asis {

typedef __nv_bfloat16 bf16;

// ---- error reporting -------------------------------------------------------------------------
char *err_buf();                 // thread-local, 512 bytes
extern std::atomic<uint64_t> g_launches;

#define ASIS_FAIL(code, ...)                          \
  do {                                                \
    snprintf(asis::err_buf(), 512, __VA_ARGS__);      \
    return (code);                                    \
  } while (0)

#define ASIS_REQUIRE(cond, ...) \
  do {                          \
    if (!(cond)) ASIS_FAIL(ASIS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define ASIS_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      ASIS_FAIL(ASIS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                __FILE__, __LINE__);                                                        \
  } while (0)

// after a <<<>>> launch
#define ASIS_LAUNCHED()                                                                     \
  do {                                                                                      \
    asis::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
    cudaError_t e__ = cudaPeekAtLastError();                                                \
    if (e__ != cudaSuccess) {                                                               \
      cudaGetLastError();                                                                   \
      ASIS_FAIL(ASIS_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                __FILE__, __LINE__);                                                        \
    }                                                                                       \
  } while (0)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool dtype_ok(int d) { return d == ASIS_F32 || d == ASIS_BF16; }
static inline size_t dtype_size(int d) { return d == ASIS_BF16 ? 2 : 4; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// 4-element vector load/store (16 B for float, 8 B for bf16); p must be aligned accordingly
__device__ __forceinline__ void load4(const float *p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16 *p, float (&v)[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16 *p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t *>(&a);
  t.y = *reinterpret_cast<uint32_t *>(&b);
  *reinterpret_cast<uint2 *>(p) = t;
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

// exact (erf) GELU and its derivative -- nn.GELU() default (dinov2/layers/mlp.py:22)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU pair for the bf16 tensor-core epilogues: erfc by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7,
// far below bf16 rounding), one MUFU.RCP + one MUFU.EX2 instead of the ~40-instruction erff.
//   h(x) = 0.5 * erfc(|x|/sqrt2) = 0.5 * poly(t) * t * exp(-x^2/2),  t = 1 / (1 + p |x|/sqrt2)
//   Phi(x) = x >= 0 ? 1 - h : h          gelu(x) = x Phi(x) = max(x, 0) - |x| h
// (the 0.5 is folded into the polynomial, the 1/sqrt2 into p; exp(-x^2/2) is shared with the density)
__device__ __forceinline__ float gelu_half_erfc(float x, float &e) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * (-0.5f * 1.4426950408889634f)));   // exp(-x^2/2)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  return (poly * t) * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e;
  const float h = gelu_half_erfc(x, e);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
// activation and derivative from ONE erfc / exp evaluation (GELU_GRAD epilogue)
__device__ __forceinline__ float gelu_and_grad_fast(float x, float &grad) {
  float e;
  const float h = gelu_half_erfc(x, e);
  grad = fmaf(x * e, 0.3989422804014327f, 0.5f + copysignf(0.5f - h, x));
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float e;
  const float h = gelu_half_erfc(x, e);
  const float cdf = 0.5f + copysignf(0.5f - h, x);
  return fmaf(x * e, 0.3989422804014327f, cdf);
}

// dispatch on a runtime dtype to a compile-time type
#define ASIS_DISPATCH_DTYPE(dt, T, ...)          \
  do {                                           \
    if ((dt) == ASIS_F32) {                      \
      typedef float T;                           \
      __VA_ARGS__;                               \
    } else {                                     \
      typedef asis::bf16 T;                      \
      __VA_ARGS__;                               \
    }                                            \
  } while (0)

}  // namespace asis
