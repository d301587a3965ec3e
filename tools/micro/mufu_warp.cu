// How fast can ONE warp issue MUFU.EX2 (independent operands), alone on its scheduler and with FFMA filler?
// nvcc -arch=sm_100a -o mufu_warp mufu_warp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int FILL>
__global__ void __launch_bounds__(512) k(float *out, long long *clk, int iters, float seed) {
  float x[16], y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = seed * (threadIdx.x + i) * 1e-3f - 1.f; y[i] = seed + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int f = 0; f < FILL; ++f) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[(i + f) & 15]) : "f"(seed));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int FILL>
void run(int threads) {
  float *out;
  long long *clk, h;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&clk, 8);
  const int iters = 4000;
  k<FILL><<<148, threads>>>(out, clk, 10, 1.f);
  k<FILL><<<148, threads>>>(out, clk, iters, 1.f);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("fill %d  %d warp(s)/SMSP: %6.2f clk per MUFU per warp, %6.2f clk per MUFU per SMSP\n", FILL, threads / 128, (double)h / iters / 16,
         (double)h / iters / 16 / (threads / 128));
}

int main() {
  for (int t : {128, 256, 512}) run<0>(t);
  for (int t : {128, 256, 512}) run<3>(t);
  for (int t : {128, 256, 512}) run<7>(t);
  return 0;
}
