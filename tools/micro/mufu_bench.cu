// Throughput of the exponential variants on one SM sub-partition set: ex2.approx.ftz.f32, ex2.approx.ftz.bf16x2,
// ex2.approx.f16x2, and the Cody-Waite polynomial on the FMA pipe.  nvcc -arch=sm_100a -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + i) * 1e-3f - 1.f;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = 0xbf80bf80u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      } else if (MODE == 3) {          // polynomial: 2^x = 2^n * p(f), magic-number rounding on the FMA pipe
        const float t = x[i] + 12582912.f;               // 1.5 * 2^23: integer part in the mantissa
        const float n = t - 12582912.f;
        const float f = x[i] - n;
        float p = fmaf(f, 0.0555f, 0.2402f);
        p = fmaf(p, f, 0.6931f);
        p = fmaf(p, f, 1.0f);
        x[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
      } else if (MODE == 4) {          // cvt.rn.bf16x2.f32 pack
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        x[i] = __uint_as_float(u[i]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int per_instr) {
  float *out;
  cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE><<<148 * 4, 256>>>(out, 100, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148 * 4, 256>>>(out, iters, 1.f);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)148 * 4 * 256 * iters * 8 * per_instr;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %7.1f Gop/s  = %5.2f results/clk/SM at %d MHz nominal\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("poly (fma pipe) 2^x", 1);
  run<4>("cvt.rn.bf16x2.f32", 1);
  return 0;
}
