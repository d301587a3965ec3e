// msda.cu -- multi-scale deformable attention for sm_100a: bilinear sampling gather (forward),
// atomic-free deterministic backward, and the fused softmax / sampling-location prep.
//
// Replaces ms_deform_attn_core_pytorch (reference backbones/ops/modules/ms_deform_attn.py:33-54)
// and the arithmetic of MSDeformAttn.forward :156-171.  HBM-bound integer/float gather work: no
// tensor cores; the design rules are coalesced 16-byte value loads, per-level spatial shapes in
// shared memory, one thread group (D/4 lanes) per (query, head) so that a bilinear corner is one
// coalesced 4*D-byte segment, and grids laid out (queries fastest, then head, then image) so the
// value slice of one (image, head) stays L2/L1 resident while it is being gathered.
//
// Every kernel first builds a per-block table of sampling points in shared memory (4 corner
// offsets + 4 weights per point, computed ONCE per (query, head, level, point) instead of once per
// lane): the first version recomputed the footprint in all D/4 lanes and was issue-bound at 1550
// warp instructions per (query, head) for 48 corner loads (profiles/r1a_msda_ncu_full.md).
//
// Backward, grad_value: contributions (query, level, point, corner) -> value pixel are a many-to-one
// scatter.  Instead of floating-point atomics: (1) count contributions per (image, head, pixel,
// query chunk) [integer counters], (2) exclusive scan per (image, head), (3) fill (id, weight)
// entries into the bucket, (4) one thread group per (image, pixel, head) walks its buckets in chunk
// order, orders each by contribution id (rank sort in registers) and reduces it -> every grad_value
// element is written exactly once, in a fixed summation order: run-to-run deterministic, no float
// atomics.  Query chunks keep buckets short (about 16 entries) whatever Lq / S is.
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
// forward: one group of GP lanes per (n, q, m); lane g owns channels [4g, 4g+4)
// grid = (ceil(Lq / items_per_block), M, N); dynamic smem = items_per_block * L*P * 32 bytes
// ---------------------------------------------------------------------------------------------
struct __align__(16) PointFwd {
  int off[4];     // element offsets (pixel * M*D) of the 4 corners inside the image's value tensor, clamped
  float w[4];     // bilinear weight * attention weight, 0 for corners outside the map
};

__device__ __forceinline__ void make_point(PointFwd &pt, const LevelInfo &lv, int l, float2 xy, float a, int MD) {
  const int H = lv.H[l], W = lv.W[l];
  const Footprint f = footprint(xy.x, xy.y, H, W);
  const bool xl = f.any && f.x0 >= 0, xr = f.any && f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
  const int xa = max(f.x0, 0), xb = min(f.x0 + 1, W - 1), ya = max(f.y0, 0), yc = min(f.y0 + 1, H - 1);
  const int base = lv.start[l];
  pt.off[0] = (base + ya * W + xa) * MD;
  pt.off[1] = (base + ya * W + xb) * MD;
  pt.off[2] = (base + yc * W + xa) * MD;
  pt.off[3] = (base + yc * W + xb) * MD;
  const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
  pt.w[0] = (yt && xl) ? ofx * ofy * a : 0.f;
  pt.w[1] = (yt && xr) ? f.fx * ofy * a : 0.f;
  pt.w[2] = (yb && xl) ? ofx * f.fy * a : 0.f;
  pt.w[3] = (yb && xr) ? f.fx * f.fy * a : 0.f;
}

template <typename VT, typename OT, int GP>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                       const int64_t *__restrict__ lsi,
                                                       const float *__restrict__ loc, const float *__restrict__ aw,
                                                       OT *__restrict__ out, int S, int M, int D, int Lq, int L,
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
               __ldg(aw + item * LP + lp), MD);
  }
  __syncthreads();
  // phase 2: gather
  const int G = D >> 2;
  const int g = threadIdx.x % GP, it = threadIdx.x / GP;
  const int q = q_first + it;
  if (q >= Lq || g >= G) return;
  const VT *vb = value + (size_t)n * S * MD + (size_t)m * D + 4 * g;
  const PointFwd *mine = tab + it * LP;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int lp = 0; lp < LP; ++lp) {
    const int4 o = *reinterpret_cast<const int4 *>(mine[lp].off);
    const float4 w = *reinterpret_cast<const float4 *>(mine[lp].w);
    float v00[4], v01[4], v10[4], v11[4];
    load4(vb + o.x, v00);
    load4(vb + o.y, v01);
    load4(vb + o.z, v10);
    load4(vb + o.w, v11);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += w.x * v00[i] + w.y * v01[i] + w.z * v10[i] + w.w * v11[i];
  }
  store4(out + (((size_t)n * Lq + q) * M + m) * (size_t)D + 4 * g, acc);
}

// ---------------------------------------------------------------------------------------------
// backward A: grad_loc / grad_attn (reduction over D by warp shuffles inside the group) and the
// per-pixel contribution counts for the grad_value buckets.
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
  int off[4];      // clamped corner offsets (as PointFwd)
  float mask[4];   // 1 for corners inside the map, else 0
  float fx, fy, a, pad;
};

// One group of GP lanes per (n, q, m); lane g owns the float4 channel vectors {g + GP*k, k < 4}
// (16 channels per lane when D = 128: 8 lanes per item, so the three per-point reductions over D
// cost 3 shuffle steps instead of 5 and a warp works on 4 items at once).
// Also counts the contributions per (pixel, query chunk) for the grad_value buckets.
template <typename VT, typename GT, int GP>
__global__ void __launch_bounds__(256) msda_bwd_locaw_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                             const int64_t *__restrict__ lsi,
                                                             const float *__restrict__ loc,
                                                             const float *__restrict__ aw,
                                                             const GT *__restrict__ gout, float *__restrict__ gloc,
                                                             float *__restrict__ gaw, int *__restrict__ counts,
                                                             int S, int M, int D, int Lq, int L, int P, int KC,
                                                             int qchunk) {
  extern __shared__ __align__(16) uint8_t msda_smem[];
  PointBwd *tab = reinterpret_cast<PointBwd *>(msda_smem);
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const int LP = L * P;
  const int ipb = blockDim.x / GP;
  const int q_first = blockIdx.x * ipb;
  const int m = blockIdx.y, n = blockIdx.z;
  const int MD = M * D;
  int *cnt = counts + ((size_t)n * M + m) * S * KC;
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
    pt.off[0] = p00 * MD; pt.off[1] = p01 * MD; pt.off[2] = p10 * MD; pt.off[3] = p11 * MD;
    pt.mask[0] = (yt && xl) ? 1.f : 0.f;
    pt.mask[1] = (yt && xr) ? 1.f : 0.f;
    pt.mask[2] = (yb && xl) ? 1.f : 0.f;
    pt.mask[3] = (yb && xr) ? 1.f : 0.f;
    pt.fx = f.fx; pt.fy = f.fy; pt.a = __ldg(aw + item * LP + lp); pt.pad = 0.f;
    if (q < Lq) {
      const int ch = q / qchunk;
      if (yt && xl) atomicAdd(cnt + (size_t)p00 * KC + ch, 1);
      if (yt && xr) atomicAdd(cnt + (size_t)p01 * KC + ch, 1);
      if (yb && xl) atomicAdd(cnt + (size_t)p10 * KC + ch, 1);
      if (yb && xr) atomicAdd(cnt + (size_t)p11 * KC + ch, 1);
    }
  }
  __syncthreads();

  const int NV = D >> 2;                      // float4 vectors per (pixel, head)
  const int g = threadIdx.x % GP, it = threadIdx.x / GP;
  int q = q_first + it;
  const bool live = q < Lq;
  if (!live) q = Lq - 1;
  const size_t item = ((size_t)n * Lq + q) * M + m;
  const VT *vb = value + (size_t)n * S * MD + (size_t)m * D;
  float go[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int v = g + GP * k;
#pragma unroll
    for (int i = 0; i < 4; ++i) go[k][i] = 0.f;
    if (v < NV) load4(gout + item * (size_t)D + 4 * v, go[k]);
  }
  const PointBwd *mine = tab + it * LP;
  for (int lp = 0; lp < LP; ++lp) {
    const int4 o = *reinterpret_cast<const int4 *>(mine[lp].off);
    const float4 mk = *reinterpret_cast<const float4 *>(mine[lp].mask);
    const float4 fr = *reinterpret_cast<const float4 *>(&mine[lp].fx);
    const float ofx = 1.f - fr.x, ofy = 1.f - fr.y;
    float d_a = 0.f, d_x = 0.f, d_y = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = g + GP * k;
      if (v < NV) {
        float v00[4], v01[4], v10[4], v11[4];
        load4(vb + o.x + 4 * v, v00);
        load4(vb + o.y + 4 * v, v01);
        load4(vb + o.z + 4 * v, v10);
        load4(vb + o.w + 4 * v, v11);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a00 = v00[i] * mk.x, a01 = v01[i] * mk.y, a10 = v10[i] * mk.z, a11 = v11[i] * mk.w;
          const float top = ofx * a00 + fr.x * a01;
          const float bot = ofx * a10 + fr.x * a11;
          d_a += go[k][i] * (ofy * top + fr.y * bot);
          d_x += go[k][i] * (ofy * (a01 - a00) + fr.y * (a11 - a10));
          d_y += go[k][i] * (bot - top);
        }
      }
    }
    d_a = group_sum<GP>(d_a);
    d_x = group_sum<GP>(d_x);
    d_y = group_sum<GP>(d_y);
    if (g == 0 && live) {
      const int l = lp / P;
      const size_t oidx = item * (size_t)LP + lp;
      gaw[oidx] = d_a;
      // d pixel / d loc = size (x_pix = loc * W - 0.5)
      reinterpret_cast<float2 *>(gloc)[oidx] = make_float2(fr.z * d_x * (float)lv.W[l], fr.z * d_y * (float)lv.H[l]);
    }
  }
}

// exclusive scan of the S counters of one (n, m); bucket storage of (n, m) starts at
// (n*M+m) * cap where cap = Lq*L*P*4 (the most contributions one (n, m) can have).
__global__ void __launch_bounds__(1024) msda_bwd_scan_kernel(const int *counts, int *__restrict__ rowptr,
                                                             int *cursor, int S /* = pixels * query chunks */, int cap) {
  __shared__ int warp_tot[32];
  __shared__ int chunk_tot;
  __shared__ int carry_s;
  const int nm = blockIdx.x;
  const int *c = counts + (size_t)nm * S;
  int *rp = rowptr + (size_t)nm * S;
  int *cu = cursor + (size_t)nm * S;  // may alias counts: each element is read before it is written
  if (threadIdx.x == 0) carry_s = nm * cap;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < S; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < S ? c[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int t = warp_tot[lane];
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;  // exclusive prefix over the warps of this chunk
      if (lane == 31) chunk_tot = ti;
    }
    __syncthreads();
    const int excl = carry_s + warp_tot[wid] + incl - v;
    if (i < S) {
      rp[i] = excl;
      cu[i] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_tot;
    __syncthreads();
  }
}

// fill: one thread per (n, q, m, l, p); entry = (contribution id within (n, m), weight).
// id = (q << idshift) + (l*P + p)*4 + corner with 2^idshift >= L*P*4: unique per (n, m), increasing
// in q, and the query index comes back with one shift.  Bucket = (pixel, query chunk).
__global__ void __launch_bounds__(256) msda_bwd_fill_kernel(const int64_t *__restrict__ ss, const int64_t *__restrict__ lsi,
                                                            const float *__restrict__ loc,
                                                            const float *__restrict__ aw, int *__restrict__ cursor,
                                                            int2 *__restrict__ entries, int S, int M, int Lq,
                                                            int L, int P, int KC, int qchunk, int idshift,
                                                            size_t total) {
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  // t = (((n*Lq + q)*M + m)*L + l)*P + p
  const int p = (int)(t % P);
  size_t r = t / P;
  const int l = (int)(r % L);
  r /= L;
  const int m = (int)(r % M);
  r /= M;
  const int q = (int)(r % Lq);
  const int n = (int)(r / Lq);
  const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + t);
  const float a = __ldg(aw + t);
  const int H = lv.H[l], W = lv.W[l];
  const Footprint f = footprint(xy.x, xy.y, H, W);
  if (!f.any) return;
  const bool xl = f.x0 >= 0, xr = f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
  int *cu = cursor + ((size_t)n * M + m) * S * KC + q / qchunk;
  const int base = lv.start[l] + f.y0 * W + f.x0;
  const int id = (q << idshift) + (l * P + p) * 4;
  const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
  if (yt && xl) entries[atomicAdd(cu + (size_t)base * KC, 1)] = make_int2(id + 0, __float_as_int(ofx * ofy * a));
  if (yt && xr) entries[atomicAdd(cu + (size_t)(base + 1) * KC, 1)] = make_int2(id + 1, __float_as_int(f.fx * ofy * a));
  if (yb && xl) entries[atomicAdd(cu + (size_t)(base + W) * KC, 1)] = make_int2(id + 2, __float_as_int(ofx * f.fy * a));
  if (yb && xr) entries[atomicAdd(cu + (size_t)(base + W + 1) * KC, 1)] = make_int2(id + 3, __float_as_int(f.fx * f.fy * a));
}

// Bring one bucket into registers (NS entries per lane), rank every entry by id (one shuffle per
// bucket entry, NS compares), and write it to its rank position in the group's smem strip.
template <int NS, int GP>
__device__ __forceinline__ void bucket_to_strip(const int2 *__restrict__ entries, int beg, int len, int wlen, int g,
                                                unsigned gbase, int2 *mystrip) {
  int id[NS], wbits[NS], rank[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const int e = k * GP + g;
    int2 t = make_int2(0x7fffffff, 0);
    if (e < len) t = entries[beg + e];
    id[k] = t.x;
    wbits[k] = t.y;
    rank[k] = 0;
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    for (int j = 0; j < GP && k * GP + j < wlen; ++j) {
      const int other = __shfl_sync(0xffffffffu, id[k], gbase + j);
#pragma unroll
      for (int kk = 0; kk < NS; ++kk) rank[kk] += (other < id[kk]) ? 1 : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k)
    if (k * GP + g < len) mystrip[rank[k]] = make_int2(id[k], wbits[k]);
}

// gather: one group of GP lanes per (n, s, m).  The pixel's KC buckets are walked in chunk order;
// each bucket [rowptr, cursor) is brought into registers (kSlots entries per lane), every entry is
// ranked by id with one shuffle per entry, written to its rank position in a per-group
// shared-memory strip, then consumed in rank order.
template <typename GT, int GP>
__global__ void __launch_bounds__(256) msda_bwd_gather_kernel(const GT *__restrict__ gout, const int *__restrict__ rowptr,
                                                              const int *__restrict__ cursor,
                                                              const int2 *__restrict__ entries,
                                                              GT *__restrict__ gvalue, int S, int M, int D, int Lq,
                                                              int KC, int idshift) {
  constexpr int kSlots = 4;
  __shared__ int2 strip[256 * kSlots];
  const int G = D >> 2;
  const int g = threadIdx.x % GP;
  int s = blockIdx.x * (blockDim.x / GP) + threadIdx.x / GP;
  const int m = blockIdx.y, n = blockIdx.z;
  const bool live = s < S;
  if (!live) s = S - 1;
  const bool lane_on = g < G;
  const size_t seg0 = (((size_t)n * M + m) * S + s) * KC;
  const GT *gb = gout + (size_t)n * Lq * M * D + (size_t)m * D + 4 * g;
  const size_t MD = (size_t)M * D;
  float acc[4] = {0, 0, 0, 0};
  const unsigned lane = threadIdx.x & 31;
  const unsigned gbase = lane & ~(unsigned)(GP - 1);
  int2 *mystrip = strip + (threadIdx.x / GP) * (GP * kSlots);

  for (int kc = 0; kc < KC; ++kc) {
    const int beg = rowptr[seg0 + kc];
    const int len = live ? cursor[seg0 + kc] - beg : 0;
    // the longest bucket in the warp decides which path every group of the warp takes, so that
    // the full-mask shuffles below stay convergent
    int wlen = len;
#pragma unroll
    for (int o = 16; o >= GP; o >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, o));
    if (wlen == 0) continue;

    if (wlen <= kSlots * GP) {
      __syncwarp();   // previous bucket's strip reads are done
      // number of register slots per lane this warp needs (warp-uniform): the common case is 1
      if (wlen <= GP) bucket_to_strip<1, GP>(entries, beg, len, wlen, g, gbase, mystrip);
      else if (wlen <= 2 * GP) bucket_to_strip<2, GP>(entries, beg, len, wlen, g, gbase, mystrip);
      else bucket_to_strip<kSlots, GP>(entries, beg, len, wlen, g, gbase, mystrip);
      __syncwarp();
      if (lane_on) {
#pragma unroll 4
        for (int r = 0; r < len; ++r) {
          const int2 t = mystrip[r];
          float v[4];
          load4(gb + (size_t)(t.x >> idshift) * MD, v);
          const float w = __int_as_float(t.y);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += w * v[i];
        }
      }
    } else {
      // long bucket (pathological sampling patterns): selection by successive minima, still in id
      // order, O(len^2 / GP) but correct for any length
      int last = -1;
      for (int r = 0; r < wlen; ++r) {
        int best = 0x7fffffff;
        float bw = 0.f;
        for (int e = g; e < len; e += GP) {
          const int2 t = entries[beg + e];
          if (t.x > last && t.x < best) {
            best = t.x;
            bw = __int_as_float(t.y);
          }
        }
#pragma unroll
        for (int o = GP / 2; o > 0; o >>= 1) {
          const int ob = __shfl_xor_sync(0xffffffffu, best, o, GP);
          const float ow = __shfl_xor_sync(0xffffffffu, bw, o, GP);
          if (ob < best) {
            best = ob;
            bw = ow;
          }
        }
        if (best != 0x7fffffff) {
          if (lane_on) {
            float v[4];
            load4(gb + (size_t)(best >> idshift) * MD, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += bw * v[i];
          }
          last = best;
        }
      }
    }
  }
  if (live && lane_on) store4(gvalue + ((size_t)n * S + s) * MD + (size_t)m * D + 4 * g, acc);
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
  ASIS_REQUIRE((size_t)S * M * D < ((size_t)1 << 31), "msda: one image's value tensor must have < 2^31 elements");
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

// query chunks per pixel bucket: about 16 contributions per bucket on average
static int query_chunks(int S, int Lq, int L, int P) {
  const long long per_pixel = ((long long)Lq * L * P * 4 + S - 1) / S;
  long long kc = (per_pixel + 15) / 16;
  if (kc < 1) kc = 1;
  if (kc > 64) kc = 64;
  if (kc > Lq) kc = Lq;
  return (int)kc;
}

static int id_shift(int L, int P) {
  int sh = 0;
  while ((1 << sh) < L * P * 4) ++sh;
  return sh;
}

template <typename VT, typename OT>
static int launch_fwd(const void *value, const int64_t *ss, const int64_t *lsi, const float *loc, const float *aw,
                      void *out, int N, int S, int M, int D, int Lq, int L, int P, cudaStream_t st) {
  const int gp = group_pad((D + 3) / 4) < 4 ? 4 : group_pad((D + 3) / 4);
  const int ipb = items_per_block(gp, L * P, (int)sizeof(PointFwd));
  const int threads = ipb * gp;
  const size_t smem = (size_t)ipb * L * P * sizeof(PointFwd);
  dim3 grid((Lq + ipb - 1) / ipb, M, N);
#define ASIS_FWD(GPV) \
  msda_fwd_kernel<VT, OT, GPV><<<grid, threads, smem, st>>>((const VT *)value, ss, lsi, loc, aw, (OT *)out, S, M, D, Lq, L, P)
  switch (gp) {
    case 4: ASIS_FWD(4); break;
    case 8: ASIS_FWD(8); break;
    case 16: ASIS_FWD(16); break;
    default: ASIS_FWD(32); break;
  }
#undef ASIS_FWD
  ASIS_LAUNCHED();
  return ASIS_OK;
}

template <typename VT, typename GT>
static int launch_bwd(const void *value, const int64_t *ss, const int64_t *lsi, const float *loc, const float *aw,
                      const void *gout, void *gvalue, float *gloc, float *gaw, int N, int S, int M, int D, int Lq,
                      int L, int P, int *counts, int *rowptr, int2 *entries, cudaStream_t st) {
  const int KC = query_chunks(S, Lq, L, P);
  const int qchunk = (Lq + KC - 1) / KC;
  const int idshift = id_shift(L, P);
  ASIS_REQUIRE(((long long)Lq << idshift) < (1LL << 31), "msda_backward: contribution id exceeds int32");
  const size_t nbuckets = (size_t)N * M * S * KC;
  ASIS_CUDA(cudaMemsetAsync(counts, 0, nbuckets * sizeof(int), st));
  {
    // 16 channels per lane: D/16 lanes per (query, head)
    const int gp = group_pad((D + 15) / 16);
    const int ipb = items_per_block(gp, L * P, (int)sizeof(PointBwd));
    const int threads = ipb * gp;
    const size_t smem = (size_t)ipb * L * P * sizeof(PointBwd);
    dim3 grid((Lq + ipb - 1) / ipb, M, N);
#define ASIS_LOCAW(GPV)                                                                                            \
  if (int rc = allow_smem(msda_bwd_locaw_kernel<VT, GT, GPV>, smem)) return rc;                                    \
  msda_bwd_locaw_kernel<VT, GT, GPV><<<grid, threads, smem, st>>>((const VT *)value, ss, lsi, loc, aw, (const GT *)gout, \
                                                                   gloc, gaw, counts, S, M, D, Lq, L, P, KC, qchunk)
    switch (gp) {
      case 1: { ASIS_LOCAW(1); } break;
      case 2: { ASIS_LOCAW(2); } break;
      case 4: { ASIS_LOCAW(4); } break;
      default: { ASIS_LOCAW(8); } break;
    }
#undef ASIS_LOCAW
    ASIS_LAUNCHED();
  }
  const int cap = Lq * L * P * 4;
  // counts -> rowptr (start) ; the counts buffer is then reused as the fill cursor
  msda_bwd_scan_kernel<<<N * M, 1024, 0, st>>>(counts, rowptr, counts, S * KC, cap);
  ASIS_LAUNCHED();
  {
    const size_t total = (size_t)N * Lq * M * L * P;
    msda_bwd_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ss, lsi, loc, aw, counts, entries, S, M, Lq, L, P, KC, qchunk, idshift, total);
    ASIS_LAUNCHED();
  }
  {
    const int gp = group_pad((D + 3) / 4) < 4 ? 4 : group_pad((D + 3) / 4);
    const int ipb = 256 / gp;
    dim3 grid((S + ipb - 1) / ipb, M, N);
    switch (gp) {
      case 4: msda_bwd_gather_kernel<GT, 4><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, KC, idshift); break;
      case 8: msda_bwd_gather_kernel<GT, 8><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, KC, idshift); break;
      case 16: msda_bwd_gather_kernel<GT, 16><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, KC, idshift); break;
      default: msda_bwd_gather_kernel<GT, 32><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, KC, idshift); break;
    }
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
    return launch_fwd<float, float>(value, spatial_shapes, level_start, sampling_loc, attn_weight, out, N, S, M, D, Lq, L, P, st);
  return launch_fwd<bf16, bf16>(value, spatial_shapes, level_start, sampling_loc, attn_weight, out, N, S, M, D, Lq, L, P, st);
}

extern "C" size_t asis_msda_backward_workspace_bytes(int N, int S, int M, int D, int Lq, int L, int P) {
  (void)D;
  if (N <= 0 || S <= 0 || M <= 0 || Lq <= 0 || L <= 0 || P <= 0) return 0;
  const size_t nms = align_up((size_t)N * M * S * query_chunks(S, Lq, L, P) * sizeof(int), 256);
  const size_t ent = align_up((size_t)N * M * Lq * L * P * 4 * sizeof(int2), 256);
  return 2 * nms + ent;
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
  const size_t need = asis_msda_backward_workspace_bytes(N, S, M, D, Lq, L, P);
  if (workspace_bytes < need) ASIS_FAIL(ASIS_ERR_WORKSPACE, "msda_backward: workspace %zu < %zu bytes", workspace_bytes, need);
  ASIS_REQUIRE(aligned16(workspace) && aligned16(value) && aligned16(grad_out) && aligned16(grad_value), "msda_backward: pointers must be 16-byte aligned");
  const size_t nms = align_up((size_t)N * M * S * query_chunks(S, Lq, L, P) * sizeof(int), 256);
  int *counts = (int *)workspace;
  int *rowptr = (int *)((char *)workspace + nms);
  int2 *entries = (int2 *)((char *)workspace + 2 * nms);
  cudaStream_t st = (cudaStream_t)stream;
  if (value_dtype == ASIS_F32)
    return launch_bwd<float, float>(value, spatial_shapes, level_start, sampling_loc, attn_weight, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L, P, counts, rowptr, entries, st);
  return launch_bwd<bf16, bf16>(value, spatial_shapes, level_start, sampling_loc, attn_weight, grad_out, grad_value, grad_loc, grad_attn, N, S, M, D, Lq, L, P, counts, rowptr, entries, st);
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
