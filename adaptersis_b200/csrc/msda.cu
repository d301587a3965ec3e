// msda.cu -- multi-scale deformable attention for sm_100a: bilinear sampling gather (forward),
// deterministic backward without floating-point atomics, and the fused softmax / sampling-location prep.
//
// Replaces ms_deform_attn_core_pytorch (reference backbones/ops/modules/ms_deform_attn.py:33-54)
// and the arithmetic of MSDeformAttn.forward :156-171.  Bandwidth-bound integer/float gather work (in
// practice the L1 data pipe and the L2->SM path rather than HBM: DESIGN.md 5.1): no tensor cores; the
// design rules are coalesced 16-byte value loads, per-level spatial shapes in
// shared memory, one thread group (16 bytes of channels per lane) per (query, head) so that a bilinear
// corner is one coalesced segment, and grids laid out (queries fastest, then head, then image) so the
// value slice of one (image, head) stays L2/L1 resident while it is being gathered.
//
// Every kernel first builds a per-block table of sampling points in shared memory (4 corner
// offsets + 4 weights per point, computed ONCE per (query, head, level, point) instead of once per
// lane): the first version recomputed the footprint in all D/4 lanes and was issue-bound at 1550
// warp instructions per (query, head) for 48 corner loads (profiles/r1a_msda_ncu_full.md).
//
// Backward, grad_value: contributions (query, level, point, corner) -> value pixel are a many-to-one
// scatter.  Instead of floating-point atomics the kernels build a CSR index of the contributions by
// pixel: (1) count per (image, head, query chunk, pixel) [integer counters, inside the grad_loc /
// grad_attn kernel], (2) scan -> start of every (pixel, chunk) bucket, pixel-major, (3) fill: one warp
// per (chunk, level, image, head) writes (grad_out row offset, weight) entries to their final positions
// in a fixed order, using cursors in shared memory that only it advances (integer adds), (4) gather: one thread
// group per (image, pixel, head) walks the pixel's contiguous entry range and accumulates -> every
// grad_value element is written exactly once, in a fixed summation order: run-to-run deterministic,
// no float atomics, no sort.
#include "common.cuh"

namespace asis {

constexpr int kMaxLevels = 8;

struct LevelInfo {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];
};

__device__ __forceinline__ void load_levels(LevelInfo &s, const int64_t *ss, const int64_t *lsi, int L) {
  if (threadIdx.x < L) {
    s.H[threadIdx.x] = (int)ss[2 * threadIdx.x];
    s.W[threadIdx.x] = (int)ss[2 * threadIdx.x + 1];
    s.start[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  __syncthreads();
}

// Bilinear footprint of one sampling point.  Pixel coordinate follows ATen's
// grid_sampler_unnormalize(align_corners=False) applied to grid = 2*loc-1 (reference :39,47-49).
struct Footprint {
  int x0, y0;
  float fx, fy;
  bool any;  // at least one corner inside the map
};
__device__ __forceinline__ Footprint footprint(float lx, float ly, int H, int W) {
  Footprint f;
  // explicit round-to-nearest steps (no FMA contraction): floor() of this value decides the
  // bilinear cell, so it is kept bit-identical to the plain IEEE evaluation of the formula
  const float x = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(2.0f, lx), -1.0f), 1.0f), (float)W), -1.0f), 0.5f);
  const float y = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(2.0f, ly), -1.0f), 1.0f), (float)H), -1.0f), 0.5f);
  f.any = (x > -1.0f) && (x < (float)W) && (y > -1.0f) && (y < (float)H);
  const float xf = floorf(x), yf = floorf(y);
  f.x0 = f.any ? (int)xf : 0;
  f.y0 = f.any ? (int)yf : 0;
  f.fx = f.any ? x - xf : 0.f;
  f.fy = f.any ? y - yf : 0.f;
  return f;
}

// ---------------------------------------------------------------------------------------------
// 16-byte channel vectors: CPL channels per lane and load (4 x f32, 8 x bf16; bf16 head dims that
// are not a multiple of 8 fall back to 4-channel / 8-byte vectors)
// ---------------------------------------------------------------------------------------------
template <int CPL, typename T> struct VecIO;
template <> struct VecIO<4, float> {
  typedef float4 Raw;
  static __device__ __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ Raw load_raw(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
  static __device__ __forceinline__ void unpack(const Raw &t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { store4(p, v); }
};
template <> struct VecIO<4, bf16> {
  typedef uint2 Raw;
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ Raw load_raw(const bf16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
  static __device__ __forceinline__ void unpack(const Raw &t, float (&v)[4]) {   // bf16 -> f32 is a 16-bit shift
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(bf16 *p, const float (&v)[4]) { store4(p, v); }
};
template <> struct VecIO<8, bf16> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ Raw load_raw(const bf16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
  static __device__ __forceinline__ void unpack(const Raw &t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(bf16 *p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
// asm volatile loads stay in program order, so a run of them is issued back to back (the compiler
// otherwise interleaves each load with the arithmetic on the previous one to save registers)
__device__ __forceinline__ float4 ldg_v(const float4 *p) {
  float4 t;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
  return t;
}
__device__ __forceinline__ uint4 ldg_v(const uint4 *p) {
  uint4 t;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
  return t;
}
__device__ __forceinline__ uint2 ldg_v(const uint2 *p) {
  uint2 t;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(p));
  return t;
}
template <int CPL, typename T>
__device__ __forceinline__ typename VecIO<CPL, T>::Raw load_raw_v(const T *p) {
  return ldg_v(reinterpret_cast<const typename VecIO<CPL, T>::Raw *>(p));
}
template <int CPL, typename T>
__device__ __forceinline__ void vec_load(const T *p, float (&v)[CPL]) { VecIO<CPL, T>::unpack(VecIO<CPL, T>::load_raw(p), v); }

// ---------------------------------------------------------------------------------------------
// forward: one group of GP lanes per (n, q, m); lane g owns channels [CPL*g, CPL*g + CPL)
// grid = (ceil(Lq / items_per_block), M, N); dynamic smem = items_per_block * L*P * 32 bytes
// ---------------------------------------------------------------------------------------------
struct __align__(16) PointFwd {
  unsigned off[4];  // byte offsets (pixel * M*D * sizeof) of the 4 corners inside the image's value tensor, clamped
  float w[4];     // bilinear weight * attention weight, 0 for corners outside the map
};

__device__ __forceinline__ void make_point(PointFwd &pt, const LevelInfo &lv, int l, float2 xy, float a, unsigned MD) {
  const int H = lv.H[l], W = lv.W[l];
  const Footprint f = footprint(xy.x, xy.y, H, W);
  const bool xl = f.any && f.x0 >= 0, xr = f.any && f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
  const int xa = max(f.x0, 0), xb = min(f.x0 + 1, W - 1), ya = max(f.y0, 0), yc = min(f.y0 + 1, H - 1);
  const int base = lv.start[l];
  pt.off[0] = (unsigned)(base + ya * W + xa) * MD;
  pt.off[1] = (unsigned)(base + ya * W + xb) * MD;
  pt.off[2] = (unsigned)(base + yc * W + xa) * MD;
  pt.off[3] = (unsigned)(base + yc * W + xb) * MD;
  const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
  pt.w[0] = (yt && xl) ? ofx * ofy * a : 0.f;
  pt.w[1] = (yt && xr) ? f.fx * ofy * a : 0.f;
  pt.w[2] = (yb && xl) ? ofx * f.fy * a : 0.f;
  pt.w[3] = (yb && xr) ? f.fx * f.fy * a : 0.f;
}

// tuning knobs (A/B builds: tools/msda_variants.sh): minimum resident blocks per SM the register allocation is held to, and
// sampling points per loop iteration (4 corner loads each)
// Measured (tools/msda_variants.sh, profiles/r2y_msda_variants.jsonl; outputs bit-identical in every variant): holding the
// forward to 6 blocks per SM (40 registers instead of 43) is 5-14 % faster on the two real shapes; 7 blocks (32 registers,
// 16 bytes of spills) gains another 4-6 % with bf16 values and loses 5-8 % with f32 ones, hence the dtype-dependent bound.
#ifndef ASIS_MSDA_FWD_MINB
#define ASIS_MSDA_FWD_MINB (sizeof(VT) == 2 ? 7 : 6)
#endif
#ifndef ASIS_MSDA_FWD_UNROLL
#define ASIS_MSDA_FWD_UNROLL 2
#endif
#define ASIS_MSDA_PRAGMA(x) _Pragma(#x)
#define ASIS_MSDA_UNROLL(n) ASIS_MSDA_PRAGMA(unroll n)
template <typename VT, int CPL, int GP>
__global__ void __launch_bounds__(256, ASIS_MSDA_FWD_MINB) msda_fwd_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                       const int64_t *__restrict__ lsi,
                                                       const float *__restrict__ loc, const float *__restrict__ aw,
                                                       VT *__restrict__ out, int S, int M, int D, int Lq, int L,
                                                       int P) {
  extern __shared__ __align__(16) uint8_t msda_smem[];
  PointFwd *tab = reinterpret_cast<PointFwd *>(msda_smem);
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const int LP = L * P;
  const int ipb = blockDim.x / GP;                 // (query, head) items per block
  const int q_first = blockIdx.x * ipb;
  const int m = blockIdx.y, n = blockIdx.z;
  const int MD = M * D;
  // phase 1: the block's sampling points, one thread per point
  for (int t = threadIdx.x; t < ipb * LP; t += blockDim.x) {
    const int it = t / LP, lp = t - it * LP;
    const int q = min(q_first + it, Lq - 1);
    const size_t item = ((size_t)n * Lq + q) * M + m;
    make_point(tab[t], lv, lp / P, __ldg(reinterpret_cast<const float2 *>(loc) + item * LP + lp),
               __ldg(aw + item * LP + lp), (unsigned)(MD * sizeof(VT)));
  }
  __syncthreads();
  // phase 2: gather
  const int G = D / CPL;
  const int g = threadIdx.x % GP, it = threadIdx.x / GP;
  const int q = q_first + it;
  if (q >= Lq || g >= G) return;
  // 64-bit base + 32-bit byte offset: two integer instructions per load address
  const char *vb = reinterpret_cast<const char *>(value + (size_t)n * S * MD + (size_t)m * D + CPL * g);
  const PointFwd *mine = tab + it * LP;
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
  ASIS_MSDA_UNROLL(ASIS_MSDA_FWD_UNROLL)
  for (int lp = 0; lp < LP; ++lp) {
    const uint4 o = *reinterpret_cast<const uint4 *>(mine[lp].off);
    const float4 w = *reinterpret_cast<const float4 *>(mine[lp].w);
    float v00[CPL], v01[CPL], v10[CPL], v11[CPL];
    vec_load<CPL>(reinterpret_cast<const VT *>(vb + o.x), v00);
    vec_load<CPL>(reinterpret_cast<const VT *>(vb + o.y), v01);
    vec_load<CPL>(reinterpret_cast<const VT *>(vb + o.z), v10);
    vec_load<CPL>(reinterpret_cast<const VT *>(vb + o.w), v11);
#pragma unroll
    for (int i = 0; i < CPL; ++i)
      acc[i] = fmaf(w.w, v11[i], fmaf(w.z, v10[i], fmaf(w.y, v01[i], fmaf(w.x, v00[i], acc[i]))));
  }
  VecIO<CPL, VT>::store(out + (((size_t)n * Lq + q) * M + m) * (size_t)D + CPL * g, acc);
}

// ---------------------------------------------------------------------------------------------
// backward A (locaw): grad_loc / grad_attn, and the contribution counts for the grad_value index.
//
// One group of GP lanes per (n, q, m); each lane owns 16 channels (NVL = 16/CPL vectors, vector
// index g + GP*k), so D = 128 needs 8 lanes and a warp works on 4 items at once.  Per sampling point
// the lane forms the four dot products t_c = <grad_out, value[corner c]> over its channels; the
// three gradients are linear in (t00, t01, t10, t11):
//     d attn = ofy (ofx t00 + fx t01) + fy (ofx t10 + fx t11)
//     d x    = ofy (t01 - t00) + fy (t11 - t10)            d y = (ofx t10 + fx t11) - (ofx t00 + fx t01)
// (corners outside the map have t = 0), so the per-channel work is 4 FMAs and the reduction over D is
// three shuffle trees of log2(GP) steps per point.
// ---------------------------------------------------------------------------------------------
template <int GP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, GP);
  return v;
}
template <>
__device__ __forceinline__ float group_sum<1>(float v) { return v; }

struct __align__(16) PointBwd {
  unsigned off[4]; // clamped corner byte offsets (as PointFwd)
  float mask[4];   // 1 for corners inside the map, else 0
  float fx, fy, aw_w, aw_h;   // fractions; attention weight * level width / height (d pixel / d loc)
};

#ifndef ASIS_MSDA_LOCAW_MINB
#define ASIS_MSDA_LOCAW_MINB 1      // (measured: 5 / 6 blocks per SM spill 24-132 bytes and are 2-20 % slower)
#endif
template <typename VT, int CPL, int GP, bool FULL>   // FULL: D == 16 * GP, every lane vector is live
__global__ void __launch_bounds__(256, ASIS_MSDA_LOCAW_MINB) msda_bwd_locaw_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                             const int64_t *__restrict__ lsi,
                                                             const float *__restrict__ loc,
                                                             const float *__restrict__ aw,
                                                             const VT *__restrict__ gout, float *__restrict__ gloc,
                                                             float *__restrict__ gaw, int *__restrict__ counts,
                                                             int S, int M, int D, int Lq, int L, int P, int KC,
                                                             int qchunk) {
  constexpr int NVL = 16 / CPL;                // vectors per lane
  extern __shared__ __align__(16) uint8_t msda_smem[];
  PointBwd *tab = reinterpret_cast<PointBwd *>(msda_smem);
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const int LP = L * P;
  const int ipb = blockDim.x / GP;
  const int q_first = blockIdx.x * ipb;
  const int m = blockIdx.y, n = blockIdx.z;
  const int MD = M * D;
  // counts[(n*M + m)][query chunk][pixel]
  int *cnt = counts + ((size_t)n * M + m) * KC * S;
  for (int t = threadIdx.x; t < ipb * LP; t += blockDim.x) {
    const int it = t / LP, lp = t - it * LP;
    const int q = q_first + it;
    const int qc = min(q, Lq - 1);
    const size_t item = ((size_t)n * Lq + qc) * M + m;
    const int l = lp / P;
    const int H = lv.H[l], W = lv.W[l];
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + item * LP + lp);
    const Footprint f = footprint(xy.x, xy.y, H, W);
    const bool xl = f.any && f.x0 >= 0, xr = f.any && f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
    const int xa = max(f.x0, 0), xb = min(f.x0 + 1, W - 1), ya = max(f.y0, 0), yc = min(f.y0 + 1, H - 1);
    const int base = lv.start[l];
    PointBwd &pt = tab[t];
    const int p00 = base + ya * W + xa, p01 = base + ya * W + xb, p10 = base + yc * W + xa, p11 = base + yc * W + xb;
    const unsigned MDB = (unsigned)(MD * sizeof(VT));
    pt.off[0] = p00 * MDB; pt.off[1] = p01 * MDB; pt.off[2] = p10 * MDB; pt.off[3] = p11 * MDB;
    pt.mask[0] = (yt && xl) ? 1.f : 0.f;
    pt.mask[1] = (yt && xr) ? 1.f : 0.f;
    pt.mask[2] = (yb && xl) ? 1.f : 0.f;
    pt.mask[3] = (yb && xr) ? 1.f : 0.f;
    const float a = __ldg(aw + item * LP + lp);
    pt.fx = f.fx; pt.fy = f.fy; pt.aw_w = a * (float)W; pt.aw_h = a * (float)H;
    if (q < Lq) {
      int *c = cnt + (size_t)(q / qchunk) * S;      // integer counters: order-independent
      if (yt && xl) atomicAdd(c + p00, 1);
      if (yt && xr) atomicAdd(c + p01, 1);
      if (yb && xl) atomicAdd(c + p10, 1);
      if (yb && xr) atomicAdd(c + p11, 1);
    }
  }
  __syncthreads();

  const int NV = D / CPL;                     // vectors per (pixel, head)
  const int g = threadIdx.x % GP, it = threadIdx.x / GP;
  int q = q_first + it;
  const bool live = q < Lq;
  if (!live) q = Lq - 1;
  const size_t item = ((size_t)n * Lq + q) * M + m;
  const char *vb = reinterpret_cast<const char *>(value + (size_t)n * S * MD + (size_t)m * D);
  float go[NVL][CPL];
#pragma unroll
  for (int k = 0; k < NVL; ++k) {
    const int v = g + GP * k;
#pragma unroll
    for (int i = 0; i < CPL; ++i) go[k][i] = 0.f;
    if (v < NV) vec_load<CPL>(gout + item * (size_t)D + CPL * v, go[k]);
  }
  const PointBwd *mine = tab + it * LP;
  const size_t obase = item * (size_t)LP;
  for (int lp = 0; lp < LP; ++lp) {
    const uint4 o = *reinterpret_cast<const uint4 *>(mine[lp].off);
    const float4 mk = *reinterpret_cast<const float4 *>(mine[lp].mask);
    const float4 fr = *reinterpret_cast<const float4 *>(&mine[lp].fx);
    float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
    typedef typename VecIO<CPL, VT>::Raw Raw;
    Raw r00[NVL], r01[NVL], r10[NVL], r11[NVL];
    // all corner loads of the point first (16 x 16 bytes in flight per lane), then the dot products
#pragma unroll
    for (int k = 0; k < NVL; ++k) {
      const int v = g + GP * k;
      if (FULL || v < NV) {
        const unsigned vo = (unsigned)(CPL * sizeof(VT)) * v;
        r00[k] = load_raw_v<CPL>(reinterpret_cast<const VT *>(vb + (o.x + vo)));
        r01[k] = load_raw_v<CPL>(reinterpret_cast<const VT *>(vb + (o.y + vo)));
        r10[k] = load_raw_v<CPL>(reinterpret_cast<const VT *>(vb + (o.z + vo)));
        r11[k] = load_raw_v<CPL>(reinterpret_cast<const VT *>(vb + (o.w + vo)));
      } else {
        r00[k] = r01[k] = r10[k] = r11[k] = VecIO<CPL, VT>::zero();
      }
    }
#pragma unroll
    for (int k = 0; k < NVL; ++k) {
      float v00[CPL], v01[CPL], v10[CPL], v11[CPL];
      VecIO<CPL, VT>::unpack(r00[k], v00);
      VecIO<CPL, VT>::unpack(r01[k], v01);
      VecIO<CPL, VT>::unpack(r10[k], v10);
      VecIO<CPL, VT>::unpack(r11[k], v11);
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        t00 = fmaf(go[k][i], v00[i], t00);
        t01 = fmaf(go[k][i], v01[i], t01);
        t10 = fmaf(go[k][i], v10[i], t10);
        t11 = fmaf(go[k][i], v11[i], t11);
      }
    }
    t00 *= mk.x; t01 *= mk.y; t10 *= mk.z; t11 *= mk.w;
    const float ofx = 1.f - fr.x, ofy = 1.f - fr.y;
    const float top = ofx * t00 + fr.x * t01, bot = ofx * t10 + fr.x * t11;
    float d_a = ofy * top + fr.y * bot;
    float d_x = ofy * (t01 - t00) + fr.y * (t11 - t10);
    float d_y = bot - top;
    d_a = group_sum<GP>(d_a);
    d_x = group_sum<GP>(d_x);
    d_y = group_sum<GP>(d_y);
    if (g == 0 && live) {
      gaw[obase + lp] = d_a;
      // d pixel / d loc = size (x_pix = loc * W - 0.5)
      reinterpret_cast<float2 *>(gloc)[obase + lp] = make_float2(fr.z * d_x, fr.w * d_y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward B (scan): counts[nm][chunk][pixel] -> start of every (pixel, chunk) bucket, pixel-major
// (all chunks of a pixel are contiguous, in chunk order), written in place; pixptr[nm][s] = start of
// pixel s, pixptr[nm][S] = end.  Entries of (n, m) start at (n*M+m) * cap, cap = Lq*L*P*4.
// One block per (n, m); a thread owns kScanTiles pixels (stride 1024) so that all count loads of a
// pass are in flight together.
// ---------------------------------------------------------------------------------------------
constexpr int kScanTiles = 8;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__global__ void __launch_bounds__(1024) msda_bwd_scan_kernel(int *__restrict__ table, int *__restrict__ pixptr,
                                                             int S, int KC, int cap) {
  __shared__ int warp_tot[kScanTiles][32];
  __shared__ int carry_s;
  const int nm = blockIdx.x;
  int *tb = table + (size_t)nm * KC * S;
  int *pp = pixptr + (size_t)nm * (S + 1);
  if (threadIdx.x == 0) carry_s = nm * cap;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < S; base += 1024 * kScanTiles) {
    int v[kScanTiles], incl[kScanTiles];
#pragma unroll
    for (int j = 0; j < kScanTiles; ++j) {
      const int s = base + j * 1024 + threadIdx.x;
      v[j] = 0;
      if (s < S) {
        for (int k0 = 0; k0 < KC; k0 += 8) {      // 8 independent loads in flight
          int c[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) c[u] = (k0 + u < KC) ? tb[(size_t)(k0 + u) * S + s] : 0;
#pragma unroll
          for (int u = 0; u < 8; ++u) v[j] += c[u];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kScanTiles; ++j) {
      incl[j] = warp_incl_scan(v[j], lane);
      if (lane == 31) warp_tot[j][wid] = incl[j];
    }
    __syncthreads();
    if (wid == 0) {
      int carry = carry_s;
#pragma unroll
      for (int j = 0; j < kScanTiles; ++j) {
        const int t = warp_tot[j][lane];
        const int ti = warp_incl_scan(t, lane);
        warp_tot[j][lane] = carry + ti - t;   // exclusive prefix of warp `lane` of tile j
        carry += __shfl_sync(0xffffffffu, ti, 31);
      }
      if (lane == 0) carry_s = carry;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kScanTiles; ++j) {
      const int s = base + j * 1024 + threadIdx.x;
      if (s < S) {
        int run = warp_tot[j][wid] + incl[j] - v[j];
        pp[s] = run;
        for (int k0 = 0; k0 < KC; k0 += 8) {
          int c[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) c[u] = (k0 + u < KC) ? tb[(size_t)(k0 + u) * S + s] : 0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (k0 + u < KC) tb[(size_t)(k0 + u) * S + s] = run;
            run += c[u];
          }
        }
        if (s == S - 1) pp[S] = run;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward C (fill): one warp per (query chunk, level, (n, m)) writes the entries (grad_out row
// offset, weight) of its contributions to their final positions.  The position inside a (pixel,
// chunk) bucket is a cursor in shared memory that only this warp advances, with integer adds issued
// in program order: a lane owns one sampling point and adds its four corners one after the other,
// 32 points per round.  The entry order of a bucket is therefore (round, corner, lane) -- fixed by
// the program and, for lanes of one instruction that hit the same pixel, by the lane arbitration of
// the shared-memory atomic unit -- so the gather below sums every pixel in the same order on every
// run (tests/test_gpu_msda.py checks bit-identical results).  No float atomics, no sort.
// CAVEAT: the order in which the lanes of ONE shared-memory atomic instruction that hit the same address are served is
// not specified by CUDA; it has been stable on every B200 run (the test would catch a change), and only the ORDER of
// the summands of such a pixel depends on it, never the set -- a different arbitration would change low-order bits of
// grad_value at pixels that two lanes of one round hit at once (rare), not correctness.
// PACK16: two 16-bit cursors per word (a bucket then holds at most 65535 entries); levels with more
// pixels than `wmax` are processed window by window.
// ---------------------------------------------------------------------------------------------
template <bool PACK16>
__global__ void __launch_bounds__(32) msda_bwd_fill_kernel(const int64_t *__restrict__ ss, const int64_t *__restrict__ lsi,
                                                           const float *__restrict__ loc,
                                                           const float *__restrict__ aw, const int *__restrict__ table,
                                                           int2 *__restrict__ entries, int S, int M, int Lq,
                                                           int L, int P, int KC, int qchunk, int wmax,
                                                           int row_bytes) {
  extern __shared__ __align__(16) uint8_t msda_smem[];
  unsigned *cur = reinterpret_cast<unsigned *>(msda_smem);
  const int kc = blockIdx.x, l = blockIdx.y, nm = blockIdx.z;
  const int n = nm / M, m = nm - n * M;
  const int lane = threadIdx.x;
  const int H = (int)ss[2 * l], W = (int)ss[2 * l + 1], start = (int)lsi[l];
  const int Sl = H * W;
  const int q0 = kc * qchunk, nq = min(Lq, q0 + qchunk) - q0;
  if (nq <= 0) return;
  const int total = nq * P, LP = L * P;          // sampling points of this (chunk, level)
  const int *rp = table + ((size_t)nm * KC + kc) * S + start;
  const float2 *loc2 = reinterpret_cast<const float2 *>(loc) + (((size_t)n * Lq + q0) * M + m) * LP + l * P;
  const float *awp = aw + (((size_t)n * Lq + q0) * M + m) * LP + l * P;
  const size_t qstride = (size_t)M * LP;

  for (int w0 = 0; w0 < Sl; w0 += wmax) {
    const int wn = min(wmax, Sl - w0);
    const int words = PACK16 ? (wn + 1) >> 1 : wn;
    for (int i = lane; i < words; i += 32) cur[i] = 0u;
    __syncwarp();
    // Software pipeline over rounds of 32 points (point t = q_local * P + p, one per lane):
    //   stage A (two rounds ahead)  load loc / attention weight
    //   stage B (one round ahead)   footprint, load the bucket starts of the 4 corners
    //   stage C                     cursor adds (all four first), then the entry stores
    float2 axy = make_float2(0.f, 0.f);
    float aa = 0.f;
    int aql = -1;
    auto stage_a = [&](int t0) {
      aql = -1;
      axy = make_float2(0.f, 0.f);
      aa = 0.f;
      if (t0 + lane < total) {
        aql = (t0 + lane) / P;
        const int p = t0 + lane - aql * P;
        axy = __ldg(loc2 + aql * qstride + p);
        aa = __ldg(awp + aql * qstride + p);
      }
    };
    int bpix[4], bstart[4], bqoff = 0;
    float bwgt[4];
    auto stage_b = [&]() {
      const Footprint fp = footprint(axy.x, axy.y, H, W);
      const bool ptok = aql >= 0 && fp.any;
      bqoff = (q0 + aql) * row_bytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cx = fp.x0 + (c & 1), cy = fp.y0 + (c >> 1);
        const int px = cy * W + cx - w0;
        const bool ok = ptok && cx >= 0 && cx < W && cy >= 0 && cy < H && px >= 0 && px < wn;
        bpix[c] = ok ? px : -1;
        bstart[c] = ok ? __ldg(rp + w0 + px) : 0;
        bwgt[c] = ((c & 1) ? fp.fx : 1.f - fp.fx) * ((c >> 1) ? fp.fy : 1.f - fp.fy) * aa;
      }
    };
    stage_a(0);
    stage_b();
    stage_a(32);
    for (int t0 = 0; t0 < total; t0 += 32) {
      int pix[4], start4[4];
      float wgt[4];
      const int qoff = bqoff;
#pragma unroll
      for (int c = 0; c < 4; ++c) { pix[c] = bpix[c]; start4[c] = bstart[c]; wgt[c] = bwgt[c]; }
      if (t0 + 32 < total) {
        stage_b();              // consumes the loads of stage A issued one round ago
        stage_a(t0 + 64);
      }
      unsigned old[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        old[c] = 0u;
        if (pix[c] >= 0) {
          if (PACK16) old[c] = atomicAdd(cur + (pix[c] >> 1), 1u << ((pix[c] & 1) * 16));
          else old[c] = atomicAdd(cur + pix[c], 1u);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (pix[c] >= 0) {
          const unsigned o = PACK16 ? ((old[c] >> ((pix[c] & 1) * 16)) & 0xffffu) : old[c];
          entries[start4[c] + (int)o] = make_int2(qoff, __float_as_int(wgt[c]));
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward D (gather): one group of GP lanes per (n, pixel, m) walks the pixel's entries in their
// stored order and accumulates weight * grad_out[query] over its CPL channels; every grad_value
// element is written exactly once.
// ---------------------------------------------------------------------------------------------
#ifndef ASIS_MSDA_GATHER_U
#define ASIS_MSDA_GATHER_U 4        // grad_out rows in flight per lane
#endif
#ifndef ASIS_MSDA_GATHER_WIDE
#define ASIS_MSDA_GATHER_WIDE 0     // > 0: long entry lists first run in unmasked steps of this many rows
#endif
// UU entries starting at e: entry loads are uniform over the group (one broadcast transaction each); entries past
// the end are clamped to the last one and masked, so all UU row loads are unconditional.  The entries are accumulated
// in their stored order whatever UU is: the result does not depend on the step width.
template <typename GT, int CPL, int UU>
__device__ __forceinline__ void gather_rows(const int2 *__restrict__ entries, const char *gb, int e, int end, float (&acc)[CPL]) {
  typedef typename VecIO<CPL, GT>::Raw Raw;
  int2 t[UU];
  Raw r[UU];
#pragma unroll
  for (int u = 0; u < UU; ++u) t[u] = __ldg(entries + min(e + u, end - 1));
#pragma unroll
  for (int u = 0; u < UU; ++u) r[u] = load_raw_v<CPL>(reinterpret_cast<const GT *>(gb + (unsigned)t[u].x));
#pragma unroll
  for (int u = 0; u < UU; ++u) {
    float v[CPL];
    VecIO<CPL, GT>::unpack(r[u], v);
    const float w = (e + u < end) ? __int_as_float(t[u].y) : 0.f;
    if (e + u < end) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) acc[i] = fmaf(w, v[i], acc[i]);
    }
  }
}

// (measured: 6 blocks per SM -- 40 registers instead of 50 / 72, 8-28 bytes of spills -- backward 7-16 % faster; 8 rows in flight
//  per lane, with or without an unmasked wide loop for long lists, 5-17 % slower: occupancy, not loads in flight, is what
//  this kernel lacks)
#ifndef ASIS_MSDA_GATHER_MINB
#define ASIS_MSDA_GATHER_MINB 6
#endif
template <typename GT, int CPL, int GP>
__global__ void __launch_bounds__(256, ASIS_MSDA_GATHER_MINB) msda_bwd_gather_kernel(const GT *__restrict__ gout, const int *__restrict__ pixptr,
                                                              const int2 *__restrict__ entries,
                                                              GT *__restrict__ gvalue, int S, int M, int D, int Lq) {
  constexpr int U = ASIS_MSDA_GATHER_U;
  const int G = D / CPL;
  const int g = threadIdx.x % GP;
  const int s = blockIdx.x * (blockDim.x / GP) + threadIdx.x / GP;
  const int m = blockIdx.y, n = blockIdx.z;
  const bool lane_on = s < S && g < G;
  if (!lane_on) return;
  const int *pp = pixptr + ((size_t)n * M + m) * (S + 1) + s;
  int e = __ldg(pp);
  const int end = __ldg(pp + 1);
  const size_t MD = (size_t)M * D;
  const char *gb = reinterpret_cast<const char *>(gout + (size_t)n * Lq * MD + (size_t)m * D + CPL * g);
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
#if ASIS_MSDA_GATHER_WIDE > 0
  for (; e + ASIS_MSDA_GATHER_WIDE <= end; e += ASIS_MSDA_GATHER_WIDE) gather_rows<GT, CPL, ASIS_MSDA_GATHER_WIDE>(entries, gb, e, end, acc);
#endif
  for (; e < end; e += U) gather_rows<GT, CPL, U>(entries, gb, e, end, acc);
  if (lane_on) VecIO<CPL, GT>::store(gvalue + ((size_t)n * S + s) * MD + (size_t)m * D + CPL * g, acc);
}

// ---------------------------------------------------------------------------------------------
// prep: softmax over L*P + sampling locations (MSDeformAttn.forward :156-171) and its backward
// one thread per (row, head)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) msda_prep_fwd_kernel(const T *__restrict__ off, const T *__restrict__ logits,
                                                            const float *__restrict__ ref, int ref_rows, int ref_levels,
                                                            int ref_dim, const int64_t *__restrict__ ss,
                                                            float *__restrict__ loc, float *__restrict__ attn,
                                                            int R, int Lq, int M, int L, int P) {
  __shared__ int sH[kMaxLevels], sW[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)ss[2 * threadIdx.x];
    sW[threadIdx.x] = (int)ss[2 * threadIdx.x + 1];
  }
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)R * M) return;
  const int m = (int)(t % M);
  const size_t r = t / M;
  const int LP = L * P;
  const T *lg = logits + t * LP;
  float mx = -INFINITY;
  for (int i = 0; i < LP; ++i) mx = fmaxf(mx, to_f(lg[i]));
  float sum = 0.f;
  for (int i = 0; i < LP; ++i) sum += expf(to_f(lg[i]) - mx);
  const float inv = 1.f / sum;
  float *ap = attn + t * LP;
  for (int i = 0; i < LP; ++i) ap[i] = expf(to_f(lg[i]) - mx) * inv;

  const size_t rr = ref_rows == R ? r : (r % Lq);
  const T *op = off + t * LP * 2;
  float *lp = loc + t * LP * 2;
  for (int l = 0; l < L; ++l) {
    const float *rp = ref + (rr * ref_levels + (ref_levels == 1 ? 0 : l)) * ref_dim;
    const float rx = rp[0], ry = rp[1];
    for (int p = 0; p < P; ++p) {
      const int i = l * P + p;
      if (ref_dim == 2) {
        lp[2 * i] = rx + to_f(op[2 * i]) / (float)sW[l];
        lp[2 * i + 1] = ry + to_f(op[2 * i + 1]) / (float)sH[l];
      } else {
        lp[2 * i] = rx + to_f(op[2 * i]) / (float)P * rp[2] * 0.5f;
        lp[2 * i + 1] = ry + to_f(op[2 * i + 1]) / (float)P * rp[3] * 0.5f;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) msda_prep_bwd_kernel(const float *__restrict__ gloc, const float *__restrict__ gattn,
                                                            const float *__restrict__ attn,
                                                            const float *__restrict__ ref, int ref_rows, int ref_levels,
                                                            int ref_dim, const int64_t *__restrict__ ss,
                                                            T *__restrict__ goff, T *__restrict__ glogits, int R,
                                                            int Lq, int M, int L, int P) {
  __shared__ int sH[kMaxLevels], sW[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)ss[2 * threadIdx.x];
    sW[threadIdx.x] = (int)ss[2 * threadIdx.x + 1];
  }
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)R * M) return;
  const size_t r = t / M;
  const int LP = L * P;
  const float *ga = gattn + t * LP;
  const float *ap = attn + t * LP;
  float dot = 0.f;
  for (int i = 0; i < LP; ++i) dot += ga[i] * ap[i];
  T *gl = glogits + t * LP;
  for (int i = 0; i < LP; ++i) gl[i] = from_f<T>(ap[i] * (ga[i] - dot));

  const size_t rr = ref_rows == R ? r : (r % Lq);
  const float *glp = gloc + t * LP * 2;
  T *go = goff + t * LP * 2;
  for (int l = 0; l < L; ++l) {
    const float *rp = ref + (rr * ref_levels + (ref_levels == 1 ? 0 : l)) * ref_dim;
    for (int p = 0; p < P; ++p) {
      const int i = l * P + p;
      if (ref_dim == 2) {
        go[2 * i] = from_f<T>(glp[2 * i] / (float)sW[l]);
        go[2 * i + 1] = from_f<T>(glp[2 * i + 1] / (float)sH[l]);
      } else {
        go[2 * i] = from_f<T>(glp[2 * i] / (float)P * rp[2] * 0.5f);
        go[2 * i + 1] = from_f<T>(glp[2 * i + 1] / (float)P * rp[3] * 0.5f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int group_pad(int G) {
  int gp = 1;
  while (gp < G) gp <<= 1;
  return gp;
}

static int check_msda_dims(int N, int S, int M, int D, int Lq, int L, int P) {
  ASIS_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && L > 0 && P > 0, "msda: non-positive dimension");
  ASIS_REQUIRE(L <= kMaxLevels, "msda: n_levels %d > %d", L, kMaxLevels);
  ASIS_REQUIRE(L * P <= 64, "msda: n_levels * n_points = %d > 64", L * P);
  ASIS_REQUIRE(D % 4 == 0 && D <= 128, "msda: head dim %d must be a multiple of 4 and <= 128", D);
  ASIS_REQUIRE(M <= 65535 && N <= 65535, "msda: n_heads / batch exceed grid limits");
  ASIS_REQUIRE((size_t)S * M * D < ((size_t)1 << 30), "msda: one image's value tensor must have < 2^30 elements");
  return ASIS_OK;
}

// items (query, head pairs) per block so that the point table fits comfortably in shared memory
static int items_per_block(int gp, int LP, int entry_bytes) {
  int ipb = 256 / gp;
  // at least one full warp per block (the kernels use full-mask shuffles)
  while (ipb * gp > 32 && ipb * LP * entry_bytes > 40 * 1024) ipb >>= 1;
  return ipb;
}

template <typename K>
static int allow_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) ASIS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return ASIS_OK;
}

// Query chunks of the grad_value index: one warp fills a (chunk, level) sequentially, so chunks are
// about 256 queries; the (chunk x pixel) cursor table is capped at 256 MB.
static void query_chunks(int N, int M, int S, int Lq, int &KC, int &qchunk) {
  long long kc = ((long long)Lq + 255) / 256;
  const long long per_chunk = (long long)N * M * S * (long long)sizeof(int);
  long long cap = ((long long)256 << 20) / per_chunk;
  if (cap < 1) cap = 1;
  if (kc > cap) kc = cap;
  qchunk = (int)((Lq + kc - 1) / kc);
  KC = (Lq + qchunk - 1) / qchunk;
}

constexpr int kFillBytes = 160 * 1024;   // shared memory one fill warp may use for its pixel cursors

struct BwdWorkspace {
  size_t table, pixptr, entries, total;
};
static BwdWorkspace bwd_workspace(int N, int S, int M, int Lq, int L, int P) {
  int KC, qchunk;
  query_chunks(N, M, S, Lq, KC, qchunk);
  BwdWorkspace w;
  w.table = align_up((size_t)N * M * KC * S * sizeof(int), 256);
  w.pixptr = align_up((size_t)N * M * ((size_t)S + 1) * sizeof(int), 256);
  w.entries = align_up((size_t)N * M * Lq * L * P * 4 * sizeof(int2), 256);
  w.total = w.table + w.pixptr + w.entries;
  return w;
}

// vector width: 16-byte loads (4 x f32, 8 x bf16); bf16 head dims that are not a multiple of 8 use 8-byte loads
template <typename VT> static int cpl_of(int D) { return (sizeof(VT) == 2 && D % 8 == 0) ? 8 : 4; }

#define ASIS_GP_CASES(MACRO, VT, CPL, gp)            \
  switch (gp) {                                      \
    case 4: { MACRO(VT, CPL, 4); } break;            \
    case 8: { MACRO(VT, CPL, 8); } break;            \
    case 16: { MACRO(VT, CPL, 16); } break;          \
    default: { MACRO(VT, CPL, 32); } break;          \
  }

template <typename VT>
static int launch_fwd(const void *value, const int64_t *ss, const int64_t *lsi, const float *loc, const float *aw,
                      void *out, int N, int S, int M, int D, int Lq, int L, int P, cudaStream_t st) {
  const int cpl = cpl_of<VT>(D);
  const int gp = group_pad(D / cpl) < 4 ? 4 : group_pad(D / cpl);
  const int ipb = items_per_block(gp, L * P, (int)sizeof(PointFwd));
  const int threads = ipb * gp;
  const size_t smem = (size_t)ipb * L * P * sizeof(PointFwd);
  dim3 grid((Lq + ipb - 1) / ipb, M, N);
#define ASIS_FWD(VTT, CPLV, GPV) \
  msda_fwd_kernel<VTT, CPLV, GPV><<<grid, threads, smem, st>>>((const VTT *)value, ss, lsi, loc, aw, (VTT *)out, S, M, D, Lq, L, P)
  if constexpr (sizeof(VT) == 2) {
    if (cpl == 8) {
      switch (gp) {
        case 4: ASIS_FWD(VT, 8, 4); break;
        case 8: ASIS_FWD(VT, 8, 8); break;
        default: ASIS_FWD(VT, 8, 16); break;
      }
    } else {
      ASIS_GP_CASES(ASIS_FWD, VT, 4, gp)
    }
  } else {
    ASIS_GP_CASES(ASIS_FWD, VT, 4, gp)
  }
#undef ASIS_FWD
  ASIS_LAUNCHED();
  return ASIS_OK;
}

template <typename VT>
static int launch_bwd(const void *value, const int64_t *ss, const int64_t *lsi, const float *loc, const float *aw,
                      const void *gout, void *gvalue, float *gloc, float *gaw, int N, int S, int M, int D, int Lq,
                      int L, int P, void *workspace, cudaStream_t st) {
  int KC, qchunk;
  query_chunks(N, M, S, Lq, KC, qchunk);
  const BwdWorkspace w = bwd_workspace(N, S, M, Lq, L, P);
  int *table = (int *)workspace;
  int *pixptr = (int *)((char *)workspace + w.table);
  int2 *entries = (int2 *)((char *)workspace + w.table + w.pixptr);
  const int cpl = cpl_of<VT>(D);
  ASIS_CUDA(cudaMemsetAsync(table, 0, (size_t)N * M * KC * S * sizeof(int), st));
  {
    // 16 channels per lane: D/16 lanes per (query, head)
    const int gp = group_pad((D + 15) / 16);
    const int ipb = items_per_block(gp, L * P, (int)sizeof(PointBwd));
    const int threads = ipb * gp;
    const size_t smem = (size_t)ipb * L * P * sizeof(PointBwd);
    dim3 grid((Lq + ipb - 1) / ipb, M, N);
#define ASIS_LOCAW_F(VTT, CPLV, GPV, FULLV)                                                                          \
  if (int rc = allow_smem(msda_bwd_locaw_kernel<VTT, CPLV, GPV, FULLV>, smem)) return rc;                            \
  msda_bwd_locaw_kernel<VTT, CPLV, GPV, FULLV><<<grid, threads, smem, st>>>((const VTT *)value, ss, lsi, loc, aw,    \
                                                                             (const VTT *)gout, gloc, gaw, table, S, \
                                                                             M, D, Lq, L, P, KC, qchunk)
#define ASIS_LOCAW(VTT, CPLV, GPV)        \
  if (D == 16 * GPV) {                    \
    ASIS_LOCAW_F(VTT, CPLV, GPV, true);   \
  } else {                                \
    ASIS_LOCAW_F(VTT, CPLV, GPV, false);  \
  }
#define ASIS_LOCAW_GP(VTT, CPLV)                 \
  switch (gp) {                                  \
    case 1: { ASIS_LOCAW(VTT, CPLV, 1); } break; \
    case 2: { ASIS_LOCAW(VTT, CPLV, 2); } break; \
    case 4: { ASIS_LOCAW(VTT, CPLV, 4); } break; \
    default: { ASIS_LOCAW(VTT, CPLV, 8); } break; \
  }
    if constexpr (sizeof(VT) == 2) {
      if (cpl == 8) { ASIS_LOCAW_GP(VT, 8) } else { ASIS_LOCAW_GP(VT, 4) }
    } else {
      ASIS_LOCAW_GP(VT, 4)
    }
#undef ASIS_LOCAW_GP
#undef ASIS_LOCAW
#undef ASIS_LOCAW_F
    ASIS_LAUNCHED();
  }
  const int cap = Lq * L * P * 4;
  msda_bwd_scan_kernel<<<N * M, 1024, 0, st>>>(table, pixptr, S, KC, cap);
  ASIS_LAUNCHED();
  {
    ASIS_REQUIRE((long long)N * M <= 65535 && L <= 65535, "msda_backward: batch * heads exceeds grid limits");
    dim3 grid(KC, L, N * M);
    // a pixel receives at most one contribution per (query, point): 16-bit cursors while qchunk * P fits
    if ((long long)qchunk * P <= 65535) {
      const int wmax = S < kFillBytes / 2 ? S : kFillBytes / 2;
      const size_t smem = align_up((size_t)wmax * sizeof(uint16_t), 16);
      if (int rc = allow_smem(msda_bwd_fill_kernel<true>, smem)) return rc;
      msda_bwd_fill_kernel<true><<<grid, 32, smem, st>>>(ss, lsi, loc, aw, table, entries, S, M, Lq, L, P, KC, qchunk, wmax, (int)(M * D * sizeof(VT)));
    } else {
      const int wmax = S < kFillBytes / 4 ? S : kFillBytes / 4;
      const size_t smem = (size_t)wmax * sizeof(int);
      if (int rc = allow_smem(msda_bwd_fill_kernel<false>, smem)) return rc;
      msda_bwd_fill_kernel<false><<<grid, 32, smem, st>>>(ss, lsi, loc, aw, table, entries, S, M, Lq, L, P, KC, qchunk, wmax, (int)(M * D * sizeof(VT)));
    }
    ASIS_LAUNCHED();
  }
  {
    const int gp = group_pad(D / cpl) < 4 ? 4 : group_pad(D / cpl);
    const int ipb = 256 / gp;
    dim3 grid((S + ipb - 1) / ipb, M, N);
#define ASIS_GATHER(VTT, CPLV, GPV) \
  msda_bwd_gather_kernel<VTT, CPLV, GPV><<<grid, 256, 0, st>>>((const VTT *)gout, pixptr, entries, (VTT *)gvalue, S, M, D, Lq)
    if constexpr (sizeof(VT) == 2) {
      if (cpl == 8) {
        switch (gp) {
          case 4: ASIS_GATHER(VT, 8, 4); break;
          case 8: ASIS_GATHER(VT, 8, 8); break;
          default: ASIS_GATHER(VT, 8, 16); break;
        }
      } else {
        ASIS_GP_CASES(ASIS_GATHER, VT, 4, gp)
      }
    } else {
      ASIS_GP_CASES(ASIS_GATHER, VT, 4, gp)
    }
#undef ASIS_GATHER
    ASIS_LAUNCHED();
  }
  return ASIS_OK;
}

}  // namespace asis

using namespace asis;

extern "C" int asis_msda_forward(const void *value, int value_dtype, const int64_t *spatial_shapes,
                                 const int64_t *level_start, const float *sampling_loc, const float *attn_weight,
                                 void *out, int out_dtype, int N, int S, int M, int D, int Lq, int L, int P,
                                 void *stream) {
  if (int rc = check_msda_dims(N, S, M, D, Lq, L, P)) return rc;
  ASIS_REQUIRE(value && spatial_shapes && level_start && sampling_loc && attn_weight && out, "msda_forward: null pointer");
  ASIS_REQUIRE(dtype_ok(value_dtype) && value_dtype == out_dtype, "msda_forward: value/out dtype must match (f32 or bf16)");
  ASIS_REQUIRE(aligned16(value) && aligned16(out) && aligned16(sampling_loc), "msda_forward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (value_dtype == ASIS_F32)
    return launch_fwd<float>(value, spatial_shapes, level_start, sampling_loc, attn_weight, out, N, S, M, D, Lq, L, P, st);
  return launch_fwd<bf16>(value, spatial_shapes, level_start, sampling_loc, attn_weight, out, N, S, M, D, Lq, L, P, st);
}

extern "C" size_t asis_msda_backward_workspace_bytes(int N, int S, int M, int D, int Lq, int L, int P) {
  (void)D;
  if (N <= 0 || S <= 0 || M <= 0 || Lq <= 0 || L <= 0 || P <= 0) return 0;
  return bwd_workspace(N, S, M, Lq, L, P).total;
}

extern "C" int asis_msda_backward(const void *value, int value_dtype, const int64_t *spatial_shapes,
                                  const int64_t *level_start, const float *sampling_loc, const float *attn_weight,
                                  const void *grad_out, int gdtype, void *grad_value, float *grad_loc,
                                  float *grad_attn, int N, int S, int M, int D, int Lq, int L, int P,
                                  void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = check_msda_dims(N, S, M, D, Lq, L, P)) return rc;
  ASIS_REQUIRE(value && spatial_shapes && level_start && sampling_loc && attn_weight && grad_out && grad_value &&
                   grad_loc && grad_attn && workspace,
               "msda_backward: null pointer");
  ASIS_REQUIRE(dtype_ok(value_dtype) && value_dtype == gdtype, "msda_backward: value/grad dtype must match (f32 or bf16)");
  ASIS_REQUIRE((size_t)N * M * Lq * L * P * 4 < ((size_t)1 << 31), "msda_backward: bucket index exceeds int32");
  ASIS_REQUIRE((size_t)Lq * M * D * 4 < ((size_t)1 << 31), "msda_backward: one image's grad_out must be < 2 GB");
  const size_t need = asis_msda_backward_workspace_bytes(N, S, M, D, Lq, L, P);
  if (workspace_bytes < need) ASIS_FAIL(ASIS_ERR_WORKSPACE, "msda_backward: workspace %zu < %zu bytes", workspace_bytes, need);
  ASIS_REQUIRE(aligned16(workspace) && aligned16(value) && aligned16(grad_out) && aligned16(grad_value), "msda_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (value_dtype == ASIS_F32)
    return launch_bwd<float>(value, spatial_shapes, level_start, sampling_loc, attn_weight, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L, P, workspace, st);
  return launch_bwd<bf16>(value, spatial_shapes, level_start, sampling_loc, attn_weight, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L, P, workspace, st);
}

extern "C" int asis_msda_prep_forward(const void *offsets, const void *logits, int in_dtype, const float *ref,
                                      int ref_rows, int ref_levels, int ref_dim, const int64_t *spatial_shapes,
                                      float *loc, float *attn, int R, int Lq, int M, int L, int P, void *stream) {
  ASIS_REQUIRE(offsets && logits && ref && spatial_shapes && loc && attn, "msda_prep_forward: null pointer");
  ASIS_REQUIRE(dtype_ok(in_dtype), "msda_prep_forward: bad dtype");
  ASIS_REQUIRE(ref_dim == 2 || ref_dim == 4, "Last dim of reference_points must be 2 or 4, but get %d instead.", ref_dim);
  ASIS_REQUIRE(R > 0 && Lq > 0 && R % Lq == 0 && (ref_rows == R || ref_rows == Lq), "msda_prep_forward: reference rows %d must be %d or %d", ref_rows, R, Lq);
  ASIS_REQUIRE((ref_levels == L || ref_levels == 1) && L <= kMaxLevels, "msda_prep_forward: reference levels %d vs %d", ref_levels, L);
  const size_t total = (size_t)R * M;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(in_dtype, T, (msda_prep_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T *)offsets, (const T *)logits, ref, ref_rows, ref_levels, ref_dim, spatial_shapes, loc, attn, R, Lq, M, L, P)));
  ASIS_LAUNCHED();
  return ASIS_OK;
}

extern "C" int asis_msda_prep_backward(const float *grad_loc, const float *grad_attn, const float *attn,
                                       const float *ref, int ref_rows, int ref_levels, int ref_dim,
                                       const int64_t *spatial_shapes, void *grad_offsets, void *grad_logits,
                                       int out_dtype, int R, int Lq, int M, int L, int P, void *stream) {
  ASIS_REQUIRE(grad_loc && grad_attn && attn && ref && spatial_shapes && grad_offsets && grad_logits, "msda_prep_backward: null pointer");
  ASIS_REQUIRE(dtype_ok(out_dtype), "msda_prep_backward: bad dtype");
  ASIS_REQUIRE(ref_dim == 2 || ref_dim == 4, "Last dim of reference_points must be 2 or 4, but get %d instead.", ref_dim);
  ASIS_REQUIRE(R > 0 && Lq > 0 && R % Lq == 0 && (ref_rows == R || ref_rows == Lq), "msda_prep_backward: reference rows");
  ASIS_REQUIRE((ref_levels == L || ref_levels == 1) && L <= kMaxLevels, "msda_prep_backward: reference levels");
  const size_t total = (size_t)R * M;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(out_dtype, T, (msda_prep_bwd_kernel<T><<<blocks, 256, 0, st>>>(grad_loc, grad_attn, attn, ref, ref_rows, ref_levels, ref_dim, spatial_shapes, (T *)grad_offsets, (T *)grad_logits, R, Lq, M, L, P)));
  ASIS_LAUNCHED();
  return ASIS_OK;
}
