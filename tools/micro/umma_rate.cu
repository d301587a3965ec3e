// Issue / execution rate of tcgen05.mma.cta_group::1.kind::f16 by instruction shape (M = 128, K = 16, N = 64 / 128 / 256),
// operands in shared memory (SS) or A in tensor memory (TS): clocks per instruction for a long back-to-back stream from
// one elected thread, one CTA per SM.  The attention backward issues 16 M128 x N64 x K16 instructions per 64-query tile.
//   nvcc -std=c++17 -arch=sm_100a -I adaptersis_b200/csrc -I include -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "tc_common.cuh"

using namespace asis;
using namespace asis::tc;

// LOAD > 0: warps 2.. keep reading (and, LOAD == 2, also writing) tensor memory the way the softmax warps of the attention
// backward do, while warp 1 issues the MMA stream
template <int N, bool TS, bool B_MN, int LOAD = 0>
__global__ void __launch_bounds__(LOAD ? 576 : 128, 1) k(long long *clk, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(128, N, 0, B_MN ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = a + 32768;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t db = B_MN ? smem_desc(b + kk * 2048, 8192, 1024) : smem_desc(b + kk * 32, 16, 1024);
          if (TS) umma_bf16_ts(tmem + 256, tmem + kk * 8, db, idesc, 1u);
          else umma_bf16(tmem + 256, smem_desc(a + kk * 32, 16, 1024), db, idesc, 1u);
        }
      }
      umma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (t0 && blockIdx.x == 0) {
      clk[0] = t1 - t0;      // issue time
      clk[1] = t2 - t0;      // until the last one has completed
    }
  }
  if (LOAD && warp >= 2) {
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = ((warp - 2) >> 2) * 32 % 256;        // score-buffer columns [0, 256): not the MMA's accumulator
    float v[32];
    uint32_t pk[16];
    float acc = 0.f;
    for (int it = 0; it < iters * 2; ++it) {
      tmem_ld32(tmem + lane_addr + col, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
      if (LOAD == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = __float_as_uint(acc) + i;
        tmem_st16u(tmem + lane_addr + col, pk);
        tmem_st_wait();
      }
    }
    if (acc == 123.456f) clk[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, bool TS, bool B_MN, int LOAD = 0>
void run(const char *name) {
  long long *clk, h[2];
  cudaMalloc(&clk, 32);
  const int iters = 2000, smem = 96 * 1024;
  cudaFuncSetAttribute(k<N, TS, B_MN, LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS, B_MN, LOAD><<<148, LOAD ? 576 : 128, smem>>>(clk, 10);
  k<N, TS, B_MN, LOAD><<<148, LOAD ? 576 : 128, smem>>>(clk, iters);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
  const double n = 4.0 * iters;
  printf("%-34s issue %6.1f clk / instr, complete %6.1f clk / instr  (tensor time at peak: %d clk)  %s\n", name, h[0] / n, h[1] / n,
         128 * N * 16 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(clk);
}

int main() {
  run<64, false, false>("M128 N64  K16  SS  B K-major");
  run<128, false, false>("M128 N128 K16  SS  B K-major");
  run<256, false, false>("M128 N256 K16  SS  B K-major");
  run<64, true, false>("M128 N64  K16  TS  B K-major");
  run<64, true, true>("M128 N64  K16  TS  B MN-major");
  run<128, true, false>("M128 N128 K16  TS  B K-major");
  run<64, true, false, 1>("M128 N64 TS + 16 warps tcgen05.ld");
  run<64, true, false, 2>("M128 N64 TS + 16 warps ld + st");
  run<128, true, false, 2>("M128 N128 TS + 16 warps ld + st");
  return 0;
}
