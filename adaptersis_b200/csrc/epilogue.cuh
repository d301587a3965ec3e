// epilogue.cuh -- the fused GEMM epilogues shared by the FFMA (fp32) and tcgen05 (bf16) kernels.
// bias / exact GELU (dinov2/layers/mlp.py:34-40), LayerScale + residual
// (dinov2/layers/block.py:91-94,112-113; layer_scale.py:26-27; adapter_blocks.py:136,139,176).
#pragma once
#include "common.cuh"

namespace asis {

struct EpiArgs {
  int kind;
  const float *bias;      // [N] or null
  const float *gamma;     // [N]           (SCALE_RESIDUAL)
  const float *residual;  // [M, ldc] f32  (SCALE_RESIDUAL)
  void *aux;              // [M, ldaux]    (GELU / SCALE_RESIDUAL: optional output; DGELU: input)
  int aux_dtype;
  int64_t ldaux;
  void *C;
  int c_dtype;
  int64_t ldc;
};

__device__ __forceinline__ float ld_any(const void *p, int dtype, size_t i) {
  return dtype == ASIS_F32 ? reinterpret_cast<const float *>(p)[i] : __bfloat162float(reinterpret_cast<const bf16 *>(p)[i]);
}
__device__ __forceinline__ void st_any(void *p, int dtype, size_t i, float v) {
  if (dtype == ASIS_F32)
    reinterpret_cast<float *>(p)[i] = v;
  else
    reinterpret_cast<bf16 *>(p)[i] = __float2bfloat16_rn(v);
}

// one output element
__device__ __forceinline__ void epi_scalar(const EpiArgs &e, int m, int n, float acc) {
  const size_t ci = (size_t)m * e.ldc + n;
  const float b = e.bias ? e.bias[n] : 0.f;
  float out;
  switch (e.kind) {
    case ASIS_EPI_GELU: {
      const float h = acc + b;
      if (e.aux) st_any(e.aux, e.aux_dtype, (size_t)m * e.ldaux + n, h);
      out = gelu_erf(h);
    } break;
    case ASIS_EPI_SCALE_RESIDUAL: {
      const float u = acc + b;
      if (e.aux) st_any(e.aux, e.aux_dtype, (size_t)m * e.ldaux + n, u);
      out = e.residual[ci] + e.gamma[n] * u;
    } break;
    case ASIS_EPI_DGELU:
      out = acc * dgelu_erf(ld_any(e.aux, e.aux_dtype, (size_t)m * e.ldaux + n));
      break;
    case ASIS_EPI_ACCUMULATE:
      out = reinterpret_cast<float *>(e.C)[ci] + acc;
      break;
    case ASIS_EPI_GELU_GRAD: {
      const float h = acc + b;
      st_any(e.aux, e.aux_dtype, (size_t)m * e.ldaux + n, dgelu_erf(h));
      out = gelu_erf(h);
    } break;
    case ASIS_EPI_MUL_AUX:
      out = acc * ld_any(e.aux, e.aux_dtype, (size_t)m * e.ldaux + n);
      break;
    default:
      out = acc + b;
  }
  st_any(e.C, e.c_dtype, ci, out);
}

static inline int check_epilogue(int epilogue, const float *gamma, const float *residual, void *aux, int aux_dtype,
                                 int c_dtype) {
  ASIS_REQUIRE(epilogue >= ASIS_EPI_NONE && epilogue <= ASIS_EPI_MUL_AUX, "gemm: unknown epilogue %d", epilogue);
  if (epilogue == ASIS_EPI_SCALE_RESIDUAL) ASIS_REQUIRE(gamma && residual, "gemm: SCALE_RESIDUAL needs gamma and residual");
  if (epilogue == ASIS_EPI_DGELU) ASIS_REQUIRE(aux, "gemm: DGELU needs aux (the saved pre-activation)");
  if (epilogue == ASIS_EPI_GELU_GRAD || epilogue == ASIS_EPI_MUL_AUX) ASIS_REQUIRE(aux, "gemm: GELU_GRAD / MUL_AUX need aux (the GELU derivative)");
  if (epilogue == ASIS_EPI_ACCUMULATE) ASIS_REQUIRE(c_dtype == ASIS_F32, "gemm: ACCUMULATE needs an f32 C");
  if (aux) ASIS_REQUIRE(dtype_ok(aux_dtype), "gemm: bad aux dtype");
  return ASIS_OK;
}

}  // namespace asis
