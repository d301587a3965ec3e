// The exponential phase of the attention softmax as an instruction mix (per element: FFMA, MUFU.EX2, FADD, half an
// F2FP pack), 128 elements per "row", run with 1 / 2 / 4 warps per SM sub-partition: clocks per row per warp.
// nvcc -arch=sm_100a -o softmax_mix softmax_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int POLY>      // POLY of every 4 elements go to the FMA-pipe polynomial
__global__ void __launch_bounds__(512) k(const float *in, uint32_t *out, float *lout, long long *clk, int iters) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
  const float scale = in[0], m = in[1];
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 128; i += 2) {
      float e[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float x = fmaf(s[i + u], scale, -m);
        if (((i + u) & 3) < POLY) {
          const float t = x + 12582912.f;
          const float n = t - 12582912.f;
          const float f = x - n;
          float p = fmaf(f, 0.0555f, 0.2402f);
          p = fmaf(p, f, 0.6931f);
          p = fmaf(p, f, 1.0f);
          e[u] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
        } else {
          e[u] = ex2(x);
        }
      }
      l4[(i >> 1) & 3] += e[0] + e[1];
      __nv_bfloat162 h = __floats2bfloat162_rn(e[0], e[1]);
      acc ^= *reinterpret_cast<uint32_t *>(&h);
      s[i] += 1e-3f;       // keep the loop from being hoisted
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  lout[blockIdx.x * blockDim.x + threadIdx.x] = l4[0] + l4[1] + l4[2] + l4[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int POLY>
void run(int threads) {
  float *in, *lout;
  uint32_t *out;
  long long *clk, h;
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&lout, 148 * 512 * 4);
  cudaMalloc(&clk, 8);
  const int iters = 2000;
  k<POLY><<<148, threads>>>(in, out, lout, clk, 10);
  k<POLY><<<148, threads>>>(in, out, lout, clk, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("poly %d/4  %2d warps/SMSP: %7.1f clk per 128-element row per warp  (%.1f clk per row per SMSP)\n", POLY, threads / 128,
         (double)h / iters, (double)h / iters / (threads / 128));
}

int main() {
  for (int t : {128, 256, 512}) run<0>(t);
  for (int t : {128, 256, 512}) run<1>(t);
  for (int t : {128, 256, 512}) run<2>(t);
  return 0;
}
