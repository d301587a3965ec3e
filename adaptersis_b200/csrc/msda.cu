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
// Backward, grad_value: contributions (query, level, point, corner) -> value pixel are a many-to-one
// scatter.  Instead of floating-point atomics: (1) count contributions per (image, head, pixel)
// [integer counters], (2) exclusive scan per (image, head), (3) fill (weight, id) entries into the
// pixel's bucket, (4) one thread group per (image, pixel, head) orders its bucket by contribution
// id (rank sort in registers) and reduces it -> every grad_value element is written exactly once,
// in a fixed summation order: run-to-run deterministic, no float atomics.
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
// grid = (ceil(Lq / groups_per_block), M, N)
// ---------------------------------------------------------------------------------------------
template <typename VT>
__device__ __forceinline__ void fwd_point(const VT *__restrict__ vb, size_t MD, int H, int W, float2 xy, float a,
                                          float (&acc)[4]) {
  // branch-free: out-of-range corners read a clamped (valid) address with weight 0, so the four
  // 16-byte loads of every point are unconditional and can all be in flight together
  const Footprint f = footprint(xy.x, xy.y, H, W);
  const bool xl = f.any && f.x0 >= 0, xr = f.any && f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
  const int xa = max(f.x0, 0), xb = min(f.x0 + 1, W - 1), ya = max(f.y0, 0), yc = min(f.y0 + 1, H - 1);
  float v00[4], v01[4], v10[4], v11[4];
  load4(vb + ((size_t)ya * W + xa) * MD, v00);
  load4(vb + ((size_t)ya * W + xb) * MD, v01);
  load4(vb + ((size_t)yc * W + xa) * MD, v10);
  load4(vb + ((size_t)yc * W + xb) * MD, v11);
  const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
  const float w00 = (yt && xl) ? ofx * ofy * a : 0.f, w01 = (yt && xr) ? f.fx * ofy * a : 0.f;
  const float w10 = (yb && xl) ? ofx * f.fy * a : 0.f, w11 = (yb && xr) ? f.fx * f.fy * a : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] += w00 * v00[i] + w01 * v01[i] + w10 * v10[i] + w11 * v11[i];
}

template <typename VT, typename OT, int GP, int PT>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                       const int64_t *__restrict__ lsi,
                                                       const float *__restrict__ loc, const float *__restrict__ aw,
                                                       OT *__restrict__ out, int S, int M, int D, int Lq, int L,
                                                       int Prt) {
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const int P = PT > 0 ? PT : Prt;
  const int G = D >> 2;
  const int g = threadIdx.x % GP;
  const int q = blockIdx.x * (blockDim.x / GP) + threadIdx.x / GP;
  const int m = blockIdx.y, n = blockIdx.z;
  if (q >= Lq || g >= G) return;

  const size_t item = ((size_t)n * Lq + q) * M + m;
  const float2 *locp = reinterpret_cast<const float2 *>(loc) + item * (size_t)(L * P);
  const float *awp = aw + item * (size_t)(L * P);
  const size_t MD = (size_t)M * D;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};

  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const VT *vb = value + ((size_t)n * S + lv.start[l]) * MD + (size_t)m * D + 4 * g;
    if (PT > 0) {
#pragma unroll
      for (int p = 0; p < PT; ++p) fwd_point(vb, MD, H, W, __ldg(locp + l * PT + p), __ldg(awp + l * PT + p), acc);
    } else {
      for (int p = 0; p < P; ++p) fwd_point(vb, MD, H, W, __ldg(locp + l * P + p), __ldg(awp + l * P + p), acc);
    }
  }
  store4(out + item * (size_t)D + 4 * g, acc);
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

template <typename VT, typename GT, int GP>
__global__ void __launch_bounds__(256) msda_bwd_locaw_kernel(const VT *__restrict__ value, const int64_t *__restrict__ ss,
                                                             const int64_t *__restrict__ lsi,
                                                             const float *__restrict__ loc,
                                                             const float *__restrict__ aw,
                                                             const GT *__restrict__ gout, float *__restrict__ gloc,
                                                             float *__restrict__ gaw, int *__restrict__ counts,
                                                             int S, int M, int D, int Lq, int L, int P) {
  __shared__ LevelInfo lv;
  load_levels(lv, ss, lsi, L);
  const int G = D >> 2;
  const int g = threadIdx.x % GP;
  int q = blockIdx.x * (blockDim.x / GP) + threadIdx.x / GP;
  const int m = blockIdx.y, n = blockIdx.z;
  // whole groups past the end still take part in the shuffles below (a warp may hold several
  // groups); they work on a clamped query and skip every store
  const bool live = q < Lq;
  if (!live) q = Lq - 1;
  const bool lane_on = g < G;

  const size_t item = ((size_t)n * Lq + q) * M + m;
  const float *locp = loc + item * (size_t)(L * P * 2);
  const float *awp = aw + item * (size_t)(L * P);
  const size_t MD = (size_t)M * D;
  float go[4] = {0, 0, 0, 0};
  if (lane_on) load4(gout + item * (size_t)D + 4 * g, go);
  int *cnt = counts + ((size_t)n * M + m) * S;

  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const VT *vb = value + ((size_t)n * S + lv.start[l]) * MD + (size_t)m * D + 4 * g;
    for (int p = 0; p < P; ++p) {
      const float2 xy = __ldg(reinterpret_cast<const float2 *>(locp) + l * P + p);
      const float a = __ldg(awp + l * P + p);
      const Footprint f = footprint(xy.x, xy.y, H, W);
      float d_a = 0.f, d_x = 0.f, d_y = 0.f;
      if (f.any) {  // uniform across the group
        const bool xl = f.x0 >= 0, xr = f.x0 + 1 < W, yt = f.y0 >= 0, yb = f.y0 + 1 < H;
        float v00[4] = {0, 0, 0, 0}, v01[4] = {0, 0, 0, 0}, v10[4] = {0, 0, 0, 0}, v11[4] = {0, 0, 0, 0};
        const VT *r0 = vb + ((size_t)f.y0 * W + f.x0) * MD;
        if (lane_on) {
          if (yt && xl) load4(r0, v00);
          if (yt && xr) load4(r0 + MD, v01);
          if (yb && xl) load4(r0 + (size_t)W * MD, v10);
          if (yb && xr) load4(r0 + (size_t)W * MD + MD, v11);
        }
        const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float top = ofx * v00[i] + f.fx * v01[i];
          const float bot = ofx * v10[i] + f.fx * v11[i];
          d_a += go[i] * (ofy * top + f.fy * bot);
          d_x += go[i] * (ofy * (v01[i] - v00[i]) + f.fy * (v11[i] - v10[i]));
          d_y += go[i] * (bot - top);
        }
        if (g == 0 && live) {
          const int base = lv.start[l] + f.y0 * W + f.x0;
          if (yt && xl) atomicAdd(cnt + base, 1);
          if (yt && xr) atomicAdd(cnt + base + 1, 1);
          if (yb && xl) atomicAdd(cnt + base + W, 1);
          if (yb && xr) atomicAdd(cnt + base + W + 1, 1);
        }
      }
      d_a = group_sum<GP>(d_a);
      d_x = group_sum<GP>(d_x);
      d_y = group_sum<GP>(d_y);
      if (g == 0 && live) {
        const size_t o = item * (size_t)(L * P) + l * P + p;
        gaw[o] = d_a;
        // d pixel / d loc = size (x_pix = loc * W - 0.5)
        reinterpret_cast<float2 *>(gloc)[o] = make_float2(a * d_x * (float)W, a * d_y * (float)H);
      }
    }
  }
}

// exclusive scan of the S counters of one (n, m); bucket storage of (n, m) starts at
// (n*M+m) * cap where cap = Lq*L*P*4 (the most contributions one (n, m) can have).
__global__ void __launch_bounds__(1024) msda_bwd_scan_kernel(const int *counts, int *__restrict__ rowptr,
                                                             int *cursor, int S, int cap) {
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
// id = ((q*L + l)*P + p)*4 + corner: unique per (n, m), so ordering a bucket by id is total.
__global__ void __launch_bounds__(256) msda_bwd_fill_kernel(const int64_t *__restrict__ ss, const int64_t *__restrict__ lsi,
                                                            const float *__restrict__ loc,
                                                            const float *__restrict__ aw, int *__restrict__ cursor,
                                                            int2 *__restrict__ entries, int S, int M, int Lq,
                                                            int L, int P, size_t total) {
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
  int *cu = cursor + ((size_t)n * M + m) * S;
  const int base = lv.start[l] + f.y0 * W + f.x0;
  const int id = ((q * L + l) * P + p) * 4;
  const float ofx = 1.f - f.fx, ofy = 1.f - f.fy;
  if (yt && xl) entries[atomicAdd(cu + base, 1)] = make_int2(id + 0, __float_as_int(ofx * ofy * a));
  if (yt && xr) entries[atomicAdd(cu + base + 1, 1)] = make_int2(id + 1, __float_as_int(f.fx * ofy * a));
  if (yb && xl) entries[atomicAdd(cu + base + W, 1)] = make_int2(id + 2, __float_as_int(ofx * f.fy * a));
  if (yb && xr) entries[atomicAdd(cu + base + W + 1, 1)] = make_int2(id + 3, __float_as_int(f.fx * f.fy * a));
}

// gather: one group of GP lanes per (n, s, m).  The bucket [rowptr, cursor) is brought into
// registers (kSlots entries per lane), every entry is ranked by id with one shuffle per entry,
// written to its rank position in a per-group shared-memory strip, then consumed in rank order.
template <typename GT, int GP>
__global__ void __launch_bounds__(256) msda_bwd_gather_kernel(const GT *__restrict__ gout, const int *__restrict__ rowptr,
                                                              const int *__restrict__ cursor,
                                                              const int2 *__restrict__ entries,
                                                              GT *__restrict__ gvalue, int S, int M, int D, int Lq,
                                                              int LP4) {
  constexpr int kSlots = 4;
  __shared__ int2 strip[256 * kSlots];
  const int G = D >> 2;
  const int g = threadIdx.x % GP;
  int s = blockIdx.x * (blockDim.x / GP) + threadIdx.x / GP;
  const int m = blockIdx.y, n = blockIdx.z;
  const bool live = s < S;
  if (!live) s = S - 1;
  const bool lane_on = g < G;
  const size_t seg = ((size_t)n * M + m) * S + s;
  const int beg = rowptr[seg];
  const int len = live ? cursor[seg] - beg : 0;
  const GT *gb = gout + (size_t)n * Lq * M * D + (size_t)m * D + 4 * g;
  const size_t MD = (size_t)M * D;
  float acc[4] = {0, 0, 0, 0};
  const unsigned lane = threadIdx.x & 31;
  const unsigned gbase = lane & ~(unsigned)(GP - 1);
  int2 *mystrip = strip + (threadIdx.x / GP) * (GP * kSlots);
  // the longest bucket in the warp decides which path every group of the warp takes, so that the
  // full-mask shuffles below stay convergent
  int wlen = len;
#pragma unroll
  for (int o = 16; o >= GP; o >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, o));

  if (wlen <= kSlots * GP) {
    int id[kSlots], wbits[kSlots], rank[kSlots];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
      const int e = k * GP + g;
      int2 t = make_int2(0x7fffffff, 0);
      if (e < len) t = entries[beg + e];
      id[k] = t.x;
      wbits[k] = t.y;
      rank[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
      for (int j = 0; j < GP && k * GP + j < wlen; ++j) {
        const int other = __shfl_sync(0xffffffffu, id[k], gbase + j);
#pragma unroll
        for (int kk = 0; kk < kSlots; ++kk) rank[kk] += (other < id[kk]) ? 1 : 0;
      }
    }
#pragma unroll
    for (int k = 0; k < kSlots; ++k)
      if (k * GP + g < len) mystrip[rank[k]] = make_int2(id[k], wbits[k]);
    __syncwarp();
    if (lane_on) {
      for (int r = 0; r < len; ++r) {
        const int2 t = mystrip[r];
        const int q = t.x / LP4;
        float v[4];
        load4(gb + (size_t)q * MD, v);
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
          const int q = best / LP4;
          float v[4];
          load4(gb + (size_t)q * MD, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += bw * v[i];
        }
        last = best;
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
  int gp = 4;
  while (gp < G) gp <<= 1;
  return gp;
}

static int check_msda_dims(int N, int S, int M, int D, int Lq, int L, int P) {
  ASIS_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && L > 0 && P > 0, "msda: non-positive dimension");
  ASIS_REQUIRE(L <= kMaxLevels, "msda: n_levels %d > %d", L, kMaxLevels);
  ASIS_REQUIRE(D % 4 == 0 && D <= 128, "msda: head dim %d must be a multiple of 4 and <= 128", D);
  ASIS_REQUIRE(M <= 65535 && N <= 65535, "msda: n_heads / batch exceed grid limits");
  ASIS_REQUIRE((size_t)Lq * L * P * 4 < ((size_t)1 << 31), "msda: too many sampling points per image");
  return ASIS_OK;
}

template <typename VT, typename OT>
static int launch_fwd(const void *value, const int64_t *ss, const int64_t *lsi, const float *loc, const float *aw,
                      void *out, int N, int S, int M, int D, int Lq, int L, int P, cudaStream_t st) {
  const int gp = group_pad(D / 4);
  const int qpb = 256 / gp;
  dim3 grid((Lq + qpb - 1) / qpb, M, N);
#define ASIS_FWD(GPV)                                                                                             \
  do {                                                                                                            \
    if (P == 4)                                                                                                   \
      msda_fwd_kernel<VT, OT, GPV, 4><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (OT *)out, S, M, \
                                                            D, Lq, L, P);                                         \
    else                                                                                                          \
      msda_fwd_kernel<VT, OT, GPV, 0><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (OT *)out, S, M, \
                                                            D, Lq, L, P);                                         \
  } while (0)
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
  const int gp = group_pad(D / 4);
  const int ipb = 256 / gp;
  const size_t nms = (size_t)N * M * S;
  ASIS_CUDA(cudaMemsetAsync(counts, 0, nms * sizeof(int), st));
  {
    dim3 grid((Lq + ipb - 1) / ipb, M, N);
    switch (gp) {
      case 4: msda_bwd_locaw_kernel<VT, GT, 4><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (const GT *)gout, gloc, gaw, counts, S, M, D, Lq, L, P); break;
      case 8: msda_bwd_locaw_kernel<VT, GT, 8><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (const GT *)gout, gloc, gaw, counts, S, M, D, Lq, L, P); break;
      case 16: msda_bwd_locaw_kernel<VT, GT, 16><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (const GT *)gout, gloc, gaw, counts, S, M, D, Lq, L, P); break;
      default: msda_bwd_locaw_kernel<VT, GT, 32><<<grid, 256, 0, st>>>((const VT *)value, ss, lsi, loc, aw, (const GT *)gout, gloc, gaw, counts, S, M, D, Lq, L, P); break;
    }
    ASIS_LAUNCHED();
  }
  const int cap = Lq * L * P * 4;
  // counts -> rowptr (start) ; counts buffer is then reused as the fill cursor
  msda_bwd_scan_kernel<<<N * M, 1024, 0, st>>>(counts, rowptr, counts, S, cap);
  ASIS_LAUNCHED();
  {
    const size_t total = (size_t)N * Lq * M * L * P;
    msda_bwd_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ss, lsi, loc, aw, counts, entries, S, M, Lq, L, P, total);
    ASIS_LAUNCHED();
  }
  {
    dim3 grid((S + ipb - 1) / ipb, M, N);
    switch (gp) {
      case 4: msda_bwd_gather_kernel<GT, 4><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, L * P * 4); break;
      case 8: msda_bwd_gather_kernel<GT, 8><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, L * P * 4); break;
      case 16: msda_bwd_gather_kernel<GT, 16><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, L * P * 4); break;
      default: msda_bwd_gather_kernel<GT, 32><<<grid, 256, 0, st>>>((const GT *)gout, rowptr, counts, entries, (GT *)gvalue, S, M, D, Lq, L * P * 4); break;
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
  const size_t nms = align_up((size_t)N * M * S * sizeof(int), 256);
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
  const size_t nms = align_up((size_t)N * M * S * sizeof(int), 256);
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
