// attention_tc.cu -- flash-style self-attention on tcgen05 for sm_100a, head_dim 64, bf16 operands,
// fp32 softmax statistics and accumulation.  softmax((q * hd^-0.5) k^T) v over the packed qkv tensor
// [B, T, 3, H, 64] exactly as dinov2/layers/attention.py:56-69 / :73-89 compute it (no mask, no
// dropout); T = 1765 is not a multiple of the 128-row tiles: TMA zero-fills rows >= T and the
// softmax masks key columns >= T.
//
// Forward  (grid: q-blocks x H x B):  warp 0 TMA, warp 1 MMA issuer, warps 2-5 softmax (1 thread =
//   1 query row).  S = Q K^T into TMEM (double buffered) -> registers -> online softmax -> P (bf16)
//   written to 128B-swizzled smem -> O_j = P V (V consumed MN-major straight from the TMA tile) ->
//   registers, rescaled and accumulated in fp32.
// Backward (two atomic-free kernels, both recompute P from the saved log-sum-exp):
//   dK/dV kernel, one CTA per key block:   S^T = K Q^T, dP^T = V dO^T  -> P^T, dS^T -> smem ->
//                                          dV += P^T dO, dK += dS^T Q   (accumulated in TMEM)
//   dQ kernel,    one CTA per query block: S = Q K^T, dP = dO V^T -> dS -> smem -> dQ += dS K
// Every gradient element is produced by exactly one CTA: deterministic, no atomics.
#include "tc_common.cuh"

namespace asis {

using namespace tc;

constexpr int ATT_THREADS = 192;
constexpr int TILE = 128;      // query / key rows per tile
constexpr int HD = 64;
constexpr int T16K = TILE * HD * 2;           // one 128 x 64 bf16 tile
constexpr int T32K = TILE * TILE * 2;         // one 128 x 128 bf16 tile (two 128 x 64 sub-tiles)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// write 32 consecutive bf16 of row `row` (columns [col0, col0+32), col0 % 32 == 0) into a K-major,
// 128B-swizzled [128 x 64*n] tile (64-column sub-tiles of 16 KB each)
__device__ __forceinline__ void store_row32_sw128(uint8_t *tile, int row, int col0, const float (&v)[32]) {
  uint8_t *sub = tile + (col0 >> 6) * T16K + (row >> 3) * 1024 + (row & 7) * 128;
  const int ch0 = (col0 & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(sub + (((ch0 + c) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ void store_out64(bf16 *dst, const float (&o)[64], float scale) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(o[8 * c + 2 * i] * scale, o[8 * c + 2 * i + 1] * scale);
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    reinterpret_cast<uint4 *>(dst)[c] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// descriptors for the tile shapes used here
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kstep) {       // [rows x 64] sub-tiles, K = columns
  return smem_desc(tile + (kstep >> 2) * T16K + (kstep & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kstep) {      // [K rows x 64], N = columns
  return smem_desc(tile + kstep * 2048, T16K, 1024);
}

struct AttnParams {
  int B, T, H;
  float scale_log2;  // hd^-0.5 * log2(e)
  float scale;
  bf16 *out;         // fwd: [B, T, H*64]
  float *lse;        // [B, H, T]   natural-log log-sum-exp of the scaled scores
  const float *dvec; // bwd: [B, H, T]  rowsum(dO * O)
  bf16 *dqkv;        // bwd: [B, T, 3, H, 64]
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int FWD_SMEM = T16K /*Q*/ + 2 * T16K /*K*/ + 2 * T16K /*V*/ + 2 * T32K /*P*/ + 1024 + 256;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sQ = smem;
  uint8_t *sK = sQ + T16K;
  uint8_t *sV = sK + 2 * T16K;
  uint8_t *sP = sV + 2 * T16K;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * T32K);
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *s_empty = bars + 7,
           *p_full = bars + 9, *o_full = bars + 11, *o_empty = bars + 13;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int C = p.H * HD;
  const int nkv = (p.T + TILE - 1) / TILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(s_empty + i, 4);
      mbar_init(p_full + i, 4);
      mbar_init(o_full + i, 1);
      mbar_init(o_empty + i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, T16K);
      tma_load_3d(&tmQKV, q_full, sQ, h * HD, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(kv_empty + s, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + s, 2 * T16K);
        tma_load_3d(&tmQKV, kv_full + s, sK + s * T16K, C + h * HD, j * TILE, b);
        tma_load_3d(&tmQKV, kv_full + s, sV + s * T16K, 2 * C + h * HD, j * TILE, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(TILE, TILE, 0, 0);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc(TILE, HD, 0, 1);     // O = P V   : V is MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(kv_full + s, (j >> 1) & 1);
        if (j >= 2) mbar_wait(s_empty + s, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tS + s * TILE, desc_kmajor(aQ, k), desc_kmajor(aK + s * T16K, k), idesc_s, k > 0);
        umma_commit(s_full + s);
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        if (j + 1 < nkv) issue_s(j + 1);
        mbar_wait(p_full + s, (j >> 1) & 1);
        if (j >= 2) mbar_wait(o_empty + s, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tO + s * HD, desc_kmajor(aP + s * T32K, k), desc_mnmajor(aV + s * T16K, k), idesc_o, k > 0);
        umma_commit(o_full + s);
        umma_commit(kv_empty + s);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;            // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(s_full + s, (j >> 1) & 1);
      tc_fence_after();
      const int kv0 = j * TILE;
      // pass 1: row maximum of the scaled scores
      float m_new = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tS + lane_addr + s * TILE + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float t = (kv0 + c * 32 + i < p.T) ? v[i] * p.scale_log2 : -INFINITY;
          m_new = fmaxf(m_new, t);
        }
      }
      const float alpha = fast_exp2(m_run - m_new);
      // pass 2: probabilities -> smem (bf16), row sum
      float l_add = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tS + lane_addr + s * TILE + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float t = (kv0 + c * 32 + i < p.T) ? fast_exp2(v[i] * p.scale_log2 - m_new) : 0.f;
          v[i] = t;
          l_add += t;
        }
        store_row32_sw128(sP + s * T32K, row, c * 32, v);
      }
      tc_fence_before();
      fence_proxy_async();   // make the generic-proxy smem writes visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty + s);
        mbar_arrive(p_full + s);
      }
      // fold in the previous block's P V, then rescale to the new maximum
      if (j > 0) {
        const int sp = (j - 1) & 1;
        mbar_wait(o_full + sp, ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(tO + lane_addr + sp * HD + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += v[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty + sp);
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] *= alpha;
      l_run = l_run * alpha + l_add;
      m_run = m_new;
    }
    {
      const int sp = (nkv - 1) & 1;
      mbar_wait(o_full + sp, ((nkv - 1) >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(tO + lane_addr + sp * HD + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] += v[i];
      }
    }
    const int t = q0 + row;
    if (t < p.T) {
      store_out64(p.out + ((size_t)b * p.T + t) * C + h * HD, o, 1.f / l_run);
      p.lse[((size_t)b * p.H + h) * p.T + t] = (m_run + log2f(l_run)) * LN2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// backward prep: D[b,h,t] = sum_e dO[b,t,h,e] * O[b,t,h,e]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16 *__restrict__ out, const bf16 *__restrict__ dout,
                                                            float *__restrict__ dvec, int B, int T, int H) {
  // one 8-lane group per (b, t, h): 8 lanes x 8 bf16 = 64 elements
  const size_t gidx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const size_t total = (size_t)B * T * H;
  float acc = 0.f;
  if (gidx < total) {
    const uint4 a = *reinterpret_cast<const uint4 *>(out + gidx * HD + sub * 8);
    const uint4 d = *reinterpret_cast<const uint4 *>(dout + gidx * HD + sub * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162 *>(&aw[i]);
      const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162 *>(&dw[i]);
      acc += __low2float(x) * __low2float(y) + __high2float(x) * __high2float(y);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (gidx < total && sub == 0) {
    const int h = (int)(gidx % H);
    const size_t bt = gidx / H;
    const int t = (int)(bt % T);
    const int b = (int)(bt / T);
    dvec[((size_t)b * H + h) * T + t] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// backward, dK / dV: one CTA per (key block, head, image); loops over query blocks
// ------------------------------------------------------------------------------------------------
constexpr int BWD_SMEM = 2 * T16K /*resident pair*/ + 2 * 2 * T16K /*streamed pair x 2 stages*/ + 2 * T32K + 2 * TILE * 4 + 1024 + 256;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sK = smem;
  uint8_t *sV = sK + T16K;
  uint8_t *sQ = sV + T16K;            // 2 stages
  uint8_t *sDO = sQ + 2 * T16K;       // 2 stages
  uint8_t *sPT = sDO + 2 * T16K;      // P^T   [keys x queries]
  uint8_t *sDST = sPT + T32K;         // dS^T
  float *sLse = reinterpret_cast<float *>(sDST + T32K);
  float *sD = sLse + TILE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sD + TILE);
  uint64_t *kv_full = bars, *q_full = bars + 1, *q_empty = bars + 3, *s_full = bars + 5, *s_empty = bars + 6,
           *p_full = bars + 7, *acc_full = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int C = p.H * HD;
  const int nq = (p.T + TILE - 1) / TILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(q_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 320;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * T16K);
      tma_load_3d(&tmQKV, kv_full, sK, C + h * HD, k0, b);
      tma_load_3d(&tmQKV, kv_full, sV, 2 * C + h * HD, k0, b);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        if (i >= 2) mbar_wait(q_empty + s, ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(q_full + s, 2 * T16K);
        tma_load_3d(&tmQKV, q_full + s, sQ + s * T16K, h * HD, i * TILE, b);
        tma_load_3d(&tmDO, q_full + s, sDO + s * T16K, h * HD, i * TILE, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(TILE, TILE, 0, 0);
      constexpr uint32_t idesc_g = make_idesc(TILE, HD, 0, 1);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aDO = smem_u32(sDO),
                     aPT = smem_u32(sPT), aDST = smem_u32(sDST);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        mbar_wait(q_full + s, (i >> 1) & 1);
        if (i >= 1) mbar_wait(s_empty, (i - 1) & 1);   // softmax finished reading S^T / dP^T of block i-1
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S^T = K Q^T
          umma_bf16(tST, desc_kmajor(aK, k), desc_kmajor(aQ + s * T16K, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dP^T = V dO^T
          umma_bf16(tDPT, desc_kmajor(aV, k), desc_kmajor(aDO + s * T16K, k), idesc_s, k > 0);
        umma_commit(s_full);
        mbar_wait(p_full, i & 1);     // P^T and dS^T are in smem
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV += P^T dO
          umma_bf16(tDV, desc_kmajor(aPT, k), desc_mnmajor(aDO + s * T16K, k), idesc_g, (i > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK += dS^T Q
          umma_bf16(tDK, desc_kmajor(aDST, k), desc_mnmajor(aQ + s * T16K, k), idesc_g, (i > 0 || k > 0));
        umma_commit(q_empty + s);     // Q/dO stage and the P^T/dS^T buffers are free once these finish
      }
      umma_commit(acc_full);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // key row inside the tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int tid = threadIdx.x - 64;        // 0..127
    const bool key_ok = k0 + row < p.T;
    const float *lse_b = p.lse + ((size_t)b * p.H + h) * p.T;
    const float *d_b = p.dvec + ((size_t)b * p.H + h) * p.T;
    for (int i = 0; i < nq; ++i) {
      const int qi = i * TILE + tid;
      // previous iteration's readers of sLse/sD are done (they passed the barrier below and the
      // MMA of this iteration cannot complete before p_full of the previous one)
      named_bar_sync(1, 128);
      sLse[tid] = qi < p.T ? lse_b[qi] * LOG2E : 0.f;
      sD[tid] = qi < p.T ? d_b[qi] : 0.f;
      named_bar_sync(1, 128);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      // the P^T / dS^T buffers were last read by the MMAs of block i-1; they are complete because
      // s_full(i) was committed after them (in-order tensor pipe)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float st[32], dp[32];
        tmem_ld32(tST + lane_addr + c * 32, st);
        tmem_ld32(tDPT + lane_addr + c * 32, dp);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int col = c * 32 + q;
          const bool ok = key_ok && (i * TILE + col < p.T);
          const float pr = ok ? fast_exp2(st[q] * p.scale_log2 - sLse[col]) : 0.f;
          st[q] = pr;
          dp[q] = ok ? pr * (dp[q] - sD[col]) : 0.f;
        }
        store_row32_sw128(sPT, row, c * 32, st);
        store_row32_sw128(sDST, row, c * 32, dp);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int t = k0 + row;
    float acc[64];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tDV + lane_addr + c * 32, v);
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[c * 32 + q] = v[q];
    }
    if (t < p.T) store_out64(p.dqkv + ((size_t)b * p.T + t) * 3 * C + 2 * C + h * HD, acc, 1.f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tDK + lane_addr + c * 32, v);
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[c * 32 + q] = v[q];
    }
    if (t < p.T) store_out64(p.dqkv + ((size_t)b * p.T + t) * 3 * C + C + h * HD, acc, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, dQ: one CTA per (query block, head, image); loops over key blocks
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sQ = smem;
  uint8_t *sDO = sQ + T16K;
  uint8_t *sK = sDO + T16K;           // 2 stages
  uint8_t *sV = sK + 2 * T16K;        // 2 stages
  uint8_t *sDS = sV + 2 * T16K;       // dS [queries x keys]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sDS + 2 * T32K + 2 * TILE * 4);
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *s_empty = bars + 6,
           *p_full = bars + 7, *acc_full = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int C = p.H * HD;
  const int nkv = (p.T + TILE - 1) / TILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDQ = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * T16K);
      tma_load_3d(&tmQKV, q_full, sQ, h * HD, q0, b);
      tma_load_3d(&tmDO, q_full, sDO, h * HD, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(kv_empty + s, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + s, 2 * T16K);
        tma_load_3d(&tmQKV, kv_full + s, sK + s * T16K, C + h * HD, j * TILE, b);
        tma_load_3d(&tmQKV, kv_full + s, sV + s * T16K, 2 * C + h * HD, j * TILE, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(TILE, TILE, 0, 0);
      constexpr uint32_t idesc_g = make_idesc(TILE, HD, 0, 1);
      const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aK = smem_u32(sK), aV = smem_u32(sV), aDS = smem_u32(sDS);
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_full + s, (j >> 1) & 1);
        if (j >= 1) mbar_wait(s_empty, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S = Q K^T
          umma_bf16(tS, desc_kmajor(aQ, k), desc_kmajor(aK + s * T16K, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dP = dO V^T
          umma_bf16(tDP, desc_kmajor(aDO, k), desc_kmajor(aV + s * T16K, k), idesc_s, k > 0);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ += dS K
          umma_bf16(tDQ, desc_kmajor(aDS, k), desc_mnmajor(aK + s * T16K, k), idesc_g, (j > 0 || k > 0));
        umma_commit(kv_empty + s);
      }
      umma_commit(acc_full);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int t = q0 + row;
    const bool row_ok = t < p.T;
    const float lse2 = row_ok ? p.lse[((size_t)b * p.H + h) * p.T + t] * LOG2E : 0.f;
    const float dsum = row_ok ? p.dvec[((size_t)b * p.H + h) * p.T + t] : 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float sv[32], dp[32];
        tmem_ld32(tS + lane_addr + c * 32, sv);
        tmem_ld32(tDP + lane_addr + c * 32, dp);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const bool ok = row_ok && (j * TILE + c * 32 + q < p.T);
          const float pr = ok ? fast_exp2(sv[q] * p.scale_log2 - lse2) : 0.f;
          dp[q] = pr * (dp[q] - dsum);
        }
        store_row32_sw128(sDS, row, c * 32, dp);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float acc[64];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tDQ + lane_addr + c * 32, v);
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[c * 32 + q] = v[q];
    }
    if (row_ok) store_out64(p.dqkv + ((size_t)b * p.T + t) * 3 * C + h * HD, acc, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int set_smem(const void *fn, int bytes) {
  ASIS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return ASIS_OK;
}

int attention_tc_forward(const void *qkv, void *out, float *lse, int B, int T, int H, int hd, cudaStream_t st) {
  ASIS_REQUIRE(hd == HD, "attention(bf16): head_dim must be 64");
  ASIS_REQUIRE(aligned16(qkv) && aligned16(out), "attention(bf16): pointers must be 16-byte aligned");
  const int C = H * HD;
  CUtensorMap tq;
  if (int rc = make_tmap_3d(&tq, qkv, 3 * C, T, B, 3 * C, (uint64_t)T * 3 * C, HD, TILE, 1)) return rc;
  static bool configured = false;
  if (!configured) {
    if (int rc = set_smem((const void *)attn_fwd_kernel, FWD_SMEM)) return rc;
    configured = true;
  }
  AttnParams p{};
  p.B = B; p.T = T; p.H = H;
  p.scale = 1.0f / sqrtf((float)HD);
  p.scale_log2 = p.scale * LOG2E;
  p.out = (bf16 *)out;
  p.lse = lse;
  dim3 grid((T + TILE - 1) / TILE, H, B);
  attn_fwd_kernel<<<grid, ATT_THREADS, FWD_SMEM, st>>>(tq, p);
  ASIS_LAUNCHED();
  return ASIS_OK;
}

size_t attention_tc_bwd_ws(int B, int T, int H, int hd) {
  (void)hd;
  return align_up((size_t)B * H * T * sizeof(float), 256);
}

int attention_tc_backward(const void *qkv, const void *out, const float *lse, const void *dout, void *dqkv, int B,
                          int T, int H, int hd, void *ws, cudaStream_t st) {
  ASIS_REQUIRE(hd == HD, "attention(bf16): head_dim must be 64");
  ASIS_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv) && ws, "attention(bf16): pointers must be 16-byte aligned");
  const int C = H * HD;
  CUtensorMap tq, tdo;
  if (int rc = make_tmap_3d(&tq, qkv, 3 * C, T, B, 3 * C, (uint64_t)T * 3 * C, HD, TILE, 1)) return rc;
  if (int rc = make_tmap_3d(&tdo, dout, C, T, B, C, (uint64_t)T * C, HD, TILE, 1)) return rc;
  static bool configured = false;
  if (!configured) {
    if (int rc = set_smem((const void *)attn_bwd_dkdv_kernel, BWD_SMEM)) return rc;
    if (int rc = set_smem((const void *)attn_bwd_dq_kernel, BWD_SMEM)) return rc;
    configured = true;
  }
  float *dvec = (float *)ws;
  {
    const size_t groups = (size_t)B * T * H;
    const unsigned blocks = (unsigned)((groups * 8 + 255) / 256);
    attn_bwd_prep_kernel<<<blocks, 256, 0, st>>>((const bf16 *)out, (const bf16 *)dout, dvec, B, T, H);
    ASIS_LAUNCHED();
  }
  AttnParams p{};
  p.B = B; p.T = T; p.H = H;
  p.scale = 1.0f / sqrtf((float)HD);
  p.scale_log2 = p.scale * LOG2E;
  p.lse = const_cast<float *>(lse);
  p.dvec = dvec;
  p.dqkv = (bf16 *)dqkv;
  dim3 grid((T + TILE - 1) / TILE, H, B);
  attn_bwd_dkdv_kernel<<<grid, ATT_THREADS, BWD_SMEM, st>>>(tq, tdo, p);
  ASIS_LAUNCHED();
  attn_bwd_dq_kernel<<<grid, ATT_THREADS, BWD_SMEM, st>>>(tq, tdo, p);
  ASIS_LAUNCHED();
  return ASIS_OK;
}

}  // namespace asis
