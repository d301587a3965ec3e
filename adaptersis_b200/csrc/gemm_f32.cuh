// gemm_f32.cuh -- internal interface of the FFMA (fp32) strided/batched GEMM used by the fp32
// parity mode (asis_gemm compute=ASIS_F32 and the fp32 attention).
#pragma once
#include "epilogue.cuh"

namespace asis {

struct GemmF32 {
  const float *A;  // element (m, k) at A[m*a_m + k*a_k + b1*a_b1 + b2*a_b2]
  int64_t a_m, a_k, a_b1, a_b2;
  const float *B;  // element (n, k)
  int64_t b_n, b_k, b_b1, b_b2;
  float *C;        // plain (batched) output when epi.kind < 0: C[m*c_m + n + b1*c_b1 + b2*c_b2] = alpha*acc
  int64_t c_m, c_b1, c_b2;
  int M, N, K, nb1, nb2;
  float alpha;
  EpiArgs epi;     // epi.kind >= 0: un-batched, fused epilogue
};

int gemm_f32_launch(const GemmF32 &p, cudaStream_t st);

}  // namespace asis
