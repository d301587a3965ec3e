// gemm_f32.cu -- fp32 FFMA GEMM (parity mode).  64x64x16 tiles, 256 threads, 4x4 micro-tiles,
// arbitrary operand strides and two batch dimensions (so attention's per-(image, head) products
// on the packed qkv tensor need no copies).  Not the performance path: bf16 mode uses gemm_tc.cu.
#include "gemm_f32.cuh"

namespace asis {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmF32 p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int b1 = blockIdx.z / p.nb2, b2 = blockIdx.z % p.nb2;
  const float *A = p.A + b1 * p.a_b1 + b2 * p.a_b2;
  const float *B = p.B + b1 * p.b_b1 + b2 * p.b_b2;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const bool a_kfast = p.a_k == 1, b_kfast = p.b_k == 1;
  // Blocked summation: `acc` holds at most 512 products, then it is folded into `tot`.  A weight gradient of the
  // convolutional stages reduces over up to 10^6 pixels; one running fp32 sum of that length loses ~3 digits
  // (measured against fp64: 2.8e-2 on a stem weight gradient where cuDNN's tree-like order gives 5e-3).
  float acc[4][4], tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = tot[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += TK) {
    if ((k0 & 511) == 0 && k0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r, k;
      if (a_kfast) { k = t & 15; r = (t >> 4) + 16 * i; } else { r = t & 63; k = (t >> 6) + 4 * i; }
      const int gm = m0 + r, gk = k0 + k;
      As[k][r] = (gm < p.M && gk < p.K) ? A[gm * p.a_m + gk * p.a_k] : 0.f;
      if (b_kfast) { k = t & 15; r = (t >> 4) + 16 * i; } else { r = t & 63; k = (t >> 6) + 4 * i; }
      const int gn = n0 + r;
      const int gk2 = k0 + k;
      Bs[k][r] = (gn < p.N && gk2 < p.K) ? B[gn * p.b_n + gk2 * p.b_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += tot[i][j];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      if (p.epi.kind >= 0)
        epi_scalar(p.epi, m, n, acc[i][j]);
      else
        p.C[m * p.c_m + n + b1 * p.c_b1 + b2 * p.c_b2] = p.alpha * acc[i][j];
    }
  }
}

int gemm_f32_launch(const GemmF32 &p, cudaStream_t st) {
  ASIS_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && p.nb1 > 0 && p.nb2 > 0, "gemm_f32: non-positive dimension");
  ASIS_REQUIRE((int64_t)p.nb1 * p.nb2 <= 65535, "gemm_f32: too many batches");
  dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, p.nb1 * p.nb2);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(p);
  ASIS_LAUNCHED();
  return ASIS_OK;
}

}  // namespace asis
