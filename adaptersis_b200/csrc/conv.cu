// conv.cu -- the convolutional stages either side of the hot path, channels-last (token-major) throughout:
//   * spatial prior module  FeatureEncoder  (backbones/encoders.py:4-74):  3x3 convolutions (+ SyncBatchNorm + ReLU),
//     3x3 / stride-2 max pooling;
//   * FeatureDecoder (backbones/decoders.py:92-164): 3x3 convolutions + BatchNorm + ReLU, the final 64 -> n_classes
//     3x3 convolution.
// A convolution is  im2col (here)  +  asis_gemm  (tcgen05 in bf16 mode):  rows = output pixels, K = (ky, kx, c) with the
// channel fastest, so a tap of a pixel is one contiguous C-vector and the GEMM consumes the matrix K-major with the
// weight viewed as [Cout, ky, kx, Cin].  The input gradient is the transposed GEMM followed by a gather over the taps
// that touch an input pixel (col2im without atomics); the weight gradient is the MN-major GEMM over the same matrix.
// BatchNorm in training mode is two kernels each way: per-channel shifted sums (fixed-order partials, no atomics) and a
// fused normalise (+ReLU) / gradient pass.  Everything here is HBM-bound and written as 16-byte vector passes.
//
// A map is stored as [B, H + 2*sp, W + 2*sp, C] with the logical H x W image at offset (sp, sp) ("storage padding";
// sp = 0 is the plain layout).  Writers fill the storage border with zeros.
#include <algorithm>

#include "common.cuh"

namespace asis {

template <typename T> struct CVec;
template <> struct CVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { load4(p, v); }
  static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { store4(p, v); }
};
template <> struct CVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16 *p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
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

struct ConvGeom {
  int B, H, W, C, sp;        // input map (logical size, channels, storage padding)
  int k, stride, pad;        // square kernel
  int Ho, Wo;
  int64_t ldk;               // row pitch of the column matrix (elements), >= k*k*C
};

// ---- im2col ------------------------------------------------------------------------------------------
// vector path: C % N == 0 (one thread per (output pixel, tap, channel vector))
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) im2col_vec_kernel(const TI *__restrict__ x, TO *__restrict__ cols, const ConvGeom g) {
  constexpr int N = CVec<TO>::N;          // output vector width; the input is read with the same element count
  const int CV = g.C / N, taps = g.k * g.k;
  const size_t total = (size_t)g.B * g.Ho * g.Wo * taps * CV;
  const int Hs = g.H + 2 * g.sp, Ws = g.W + 2 * g.sp;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    size_t r = t / CV;
    const int tap = (int)(r % taps);
    r /= taps;                                  // output pixel index (b, oy, ox)
    const int ox = (int)(r % g.Wo);
    const size_t r2 = r / g.Wo;
    const int oy = (int)(r2 % g.Ho);
    const int b = (int)(r2 / g.Ho);
    const int iy = oy * g.stride - g.pad + tap / g.k, ix = ox * g.stride - g.pad + tap % g.k;
    float v[N];
    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
      const TI *src = x + (((size_t)b * Hs + iy + g.sp) * Ws + ix + g.sp) * g.C + (size_t)cv * N;
      if (sizeof(TI) == sizeof(TO)) {
        CVec<TO>::load(reinterpret_cast<const TO *>(src), v);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = to_f(src[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = 0.f;
    }
    CVec<TO>::store(cols + r * g.ldk + (size_t)tap * g.C + (size_t)cv * N, v);
  }
}

// scalar path (C = 3: the image); also zero-fills the K padding columns [k*k*C, ldk)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) im2col_scalar_kernel(const TI *__restrict__ x, TO *__restrict__ cols, const ConvGeom g) {
  const int taps = g.k * g.k;
  const size_t total = (size_t)g.B * g.Ho * g.Wo * g.ldk;
  const int Hs = g.H + 2 * g.sp, Ws = g.W + 2 * g.sp;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(t % g.ldk);
    const size_t r = t / g.ldk;
    float v = 0.f;
    if (kk < taps * g.C) {
      const int tap = kk / g.C, c = kk % g.C;
      const int ox = (int)(r % g.Wo);
      const size_t r2 = r / g.Wo;
      const int oy = (int)(r2 % g.Ho);
      const int b = (int)(r2 / g.Ho);
      const int iy = oy * g.stride - g.pad + tap / g.k, ix = ox * g.stride - g.pad + tap % g.k;
      if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = to_f(x[(((size_t)b * Hs + iy + g.sp) * Ws + ix + g.sp) * g.C + c]);
    }
    cols[t] = from_f<TO>(v);
  }
}

// ---- col2im: dx[b, iy, ix, :] = sum over the (ky, kx) whose output pixel exists of dcols[(b, oy, ox), (ky, kx), :] -----
template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T *__restrict__ dcols, T *__restrict__ dx, const ConvGeom g) {
  constexpr int N = CVec<T>::N;
  const int CV = g.C / N;
  const int Hs = g.H + 2 * g.sp, Ws = g.W + 2 * g.sp;
  const size_t total = (size_t)g.B * Hs * Ws * CV;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    size_t r = t / CV;
    const int sx = (int)(r % Ws);
    r /= Ws;
    const int sy = (int)(r % Hs);
    const int b = (int)(r / Hs);
    const int iy = sy - g.sp, ix = sx - g.sp;
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
      for (int ky = 0; ky < g.k; ++ky) {
        const int ny = iy + g.pad - ky;
        if (ny < 0 || ny % g.stride) continue;
        const int oy = ny / g.stride;
        if (oy >= g.Ho) continue;
        for (int kx = 0; kx < g.k; ++kx) {
          const int nx = ix + g.pad - kx;
          if (nx < 0 || nx % g.stride) continue;
          const int ox = nx / g.stride;
          if (ox >= g.Wo) continue;
          float v[N];
          CVec<T>::load(dcols + (((size_t)b * g.Ho + oy) * g.Wo + ox) * g.ldk + (size_t)(ky * g.k + kx) * g.C + (size_t)cv * N, v);
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i] += v[i];
        }
      }
    }
    CVec<T>::store(dx + t * N, acc);          // border of a padded map: zeros
  }
}

// ---- per-channel statistics -------------------------------------------------------------------------
// block (by) owns a strip of pixels; thread (tx) owns one channel vector and strides over the strip's pixels with
// ty (PL = 256 / cl pixel lanes): partial[(by*PL + ty), 0:C] = sum (x - shift), partial[.., C:2C] = sum (x - shift)^2 over the logical pixels.
// mode 1 (backward): sums of dz and dz * xhat with dz = dy * relu'(a (x - mean) + b), xhat = (x - mean) * rstd.
struct StatArgs {
  int B, H, W, C, sp, sp_dy;              // storage padding of x and of dy
  const float *shift;                     // mode 0
  const float *a, *b, *mean, *rstd;       // mode 1 (a, b: the forward's per-channel scale / offset)
  int relu;
};

template <typename T, int MODE>
__global__ void __launch_bounds__(256) chan_stats_kernel(const T *__restrict__ x, const T *__restrict__ dy, float *__restrict__ partial,
                                                         const StatArgs s, int pix_per_block, int cl) {
  constexpr int N = CVec<T>::N;
  const int CV = s.C / N;
  // cl channel-vector lanes x (256 / cl) pixel lanes: a 64-channel bf16 map has only 8 channel vectors, so a fixed 32 x 8
  // split left three quarters of the threads idle on the largest maps of the step (stem, decoder_4)
  const int tx = threadIdx.x % cl, ty = threadIdx.x / cl, PL = 256 / cl;
  __shared__ float red[256 * 2 * N];
  const int Hs = s.H + 2 * s.sp, Ws = s.W + 2 * s.sp, Hd = s.H + 2 * s.sp_dy, Wd = s.W + 2 * s.sp_dy;
  const size_t npix = (size_t)s.B * s.H * s.W;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  // (the loop bound is uniform over the block -- idle lanes run it with clamped addresses -- because it contains barriers)
  for (int cv0 = blockIdx.y * cl; cv0 < CV; cv0 += gridDim.y * cl) {
    const int cv_raw = cv0 + tx;
    const bool cv_ok = cv_raw < CV;
    const int cv = cv_ok ? cv_raw : CV - 1;
    float k0[N], k1[N], k2[N], k3[N], s1[N], s2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { s1[i] = s2[i] = 0.f; }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) k0[i] = s.shift[cv * N + i];
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        k0[i] = s.a[cv * N + i]; k1[i] = s.b[cv * N + i]; k2[i] = s.mean[cv * N + i]; k3[i] = s.rstd[cv * N + i];
      }
    }
    for (size_t pa = p0 + ty; pa < p1; pa += 2 * PL) {           // two pixels per iteration: twice the loads in flight
      float v[2][N], g[2][N];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t p = pa + (size_t)u * PL;
        ok[u] = p < p1;
        const size_t pc = ok[u] ? p : pa;
        const int ix = (int)(pc % s.W);
        const size_t q = pc / s.W;
        const int iy = (int)(q % s.H);
        const size_t b = q / s.H;
        CVec<T>::load(x + ((b * Hs + iy + s.sp) * Ws + ix + s.sp) * s.C + (size_t)cv * N, v[u]);
        if (MODE == 1) CVec<T>::load(dy + ((b * Hd + iy + s.sp_dy) * Wd + ix + s.sp_dy) * s.C + (size_t)cv * N, g[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float d = v[u][i] - k0[i];
            s1[i] += d;
            s2[i] = fmaf(d, d, s2[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float dz = (s.relu && fmaf(k0[i], v[u][i] - k2[i], k1[i]) <= 0.f) ? 0.f : g[u][i];
            s1[i] += dz;
            s2[i] = fmaf(dz, (v[u][i] - k2[i]) * k3[i], s2[i]);
          }
        }
      }
    }
    // the pixel lanes of the block are added in shared memory, in a fixed order: ONE partial row per block (with a row
    // per pixel lane the final reduction walked 19 k rows on 4 blocks: 63 us per call)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      red[(ty * cl + tx) * 2 * N + i] = s1[i];
      red[(ty * cl + tx) * 2 * N + N + i] = s2[i];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < cl * 2 * N; o += 256) {        // o = (tx, which, i)
      float a = 0.f;
      for (int y = 0; y < PL; ++y) a += red[y * cl * 2 * N + o];
      const int txo = o / (2 * N), rem = o % (2 * N), which = rem / N, i = rem % N;
      const int cvo = cv0 + txo;
      if (cvo < CV) partial[(size_t)blockIdx.x * 2 * s.C + (size_t)which * s.C + (size_t)cvo * N + i] = a;
    }
  }
}

// fixed-order reduction of the partial rows (same scheme as norm.cu's reduce_partials_kernel)
__global__ void __launch_bounds__(256) conv_reduce_partials_kernel(const float *__restrict__ partial, int nparts, int ncols,
                                                                   float *__restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (c < ncols)
    for (int p = ty; p < nparts; p += 8) t += partial[(size_t)p * ncols + c];
  red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < ncols) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][tx];
    out[c] = a;
  }
}

// ---- BatchNorm apply (+ReLU) and its gradient --------------------------------------------------------
// forward:  y = act(a[c] * (x - mean[c]) + b[c])        (a = weight * rstd, b = bias: the centred form -- a x + (b - mean a)
//           cancels when |mean| >> std and moves the sign of near-zero pre-activations, i.e. the ReLU mask); y may have
//           another storage padding than x; its border is written with zeros
// backward: dx = a[c] * (dz - c1[c] - xhat * c2[c]),  dz = dy * relu'(a (x - mean) + b),  c1 = mean(dz), c2 = mean(dz * xhat)
struct ApplyArgs {
  int B, H, W, C, sp_in, sp_out, sp_dy;   // storage padding of x, of the output (y or dx) and of dy
  const float *a, *b, *mean, *rstd, *c1, *c2;
  int relu;
};

template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(256) bn_apply_kernel(const TI *__restrict__ x, const TI *__restrict__ dy, TO *__restrict__ out,
                                                       const ApplyArgs s) {
  constexpr int N = CVec<TO>::N;
  const int CV = s.C / N;
  const int Hi = s.H + 2 * s.sp_in, Wi = s.W + 2 * s.sp_in, Ho = s.H + 2 * s.sp_out, Wo = s.W + 2 * s.sp_out;
  const size_t total = (size_t)s.B * Ho * Wo * CV;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    size_t r = t / CV;
    const int sx = (int)(r % Wo);
    r /= Wo;
    const int sy = (int)(r % Ho);
    const size_t b = r / Ho;
    const int iy = sy - s.sp_out, ix = sx - s.sp_out;
    float o[N];
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = 0.f;
    if (iy >= 0 && iy < s.H && ix >= 0 && ix < s.W) {
      const size_t off = ((b * Hi + iy + s.sp_in) * Wi + ix + s.sp_in) * s.C + (size_t)cv * N;
      float v[N];
      if (sizeof(TI) == sizeof(TO)) {
        CVec<TO>::load(reinterpret_cast<const TO *>(x + off), v);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = to_f(x[off + i]);
      }
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float y = fmaf(s.a[cv * N + i], v[i] - s.mean[cv * N + i], s.b[cv * N + i]);
          o[i] = s.relu ? fmaxf(y, 0.f) : y;
        }
      } else {
        const size_t offd = ((b * (s.H + 2 * s.sp_dy) + iy + s.sp_dy) * (s.W + 2 * s.sp_dy) + ix + s.sp_dy) * s.C + (size_t)cv * N;
        float g[N];
        if (sizeof(TI) == sizeof(TO)) {
          CVec<TO>::load(reinterpret_cast<const TO *>(dy + offd), g);
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) g[i] = to_f(dy[offd + i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const int c = cv * N + i;
          const float a = s.a[c];
          const float dz = (s.relu && fmaf(a, v[i] - s.mean[c], s.b[c]) <= 0.f) ? 0.f : g[i];
          const float xhat = (v[i] - s.mean[c]) * s.rstd[c];
          o[i] = a * (dz - s.c1[c] - xhat * s.c2[c]);
        }
      }
    }
    CVec<TO>::store(out + t * N, o);
  }
}

// ---- 3x3 / stride-2 / pad-1 max pooling (encoders.py:20) ----------------------------------------------
// forward stores, per output element, which of the 9 taps won (first maximum in row-major window order, as ATen);
// backward gathers: an input pixel collects the gradient of the <= 4 windows that selected it.
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, uint8_t *__restrict__ idx,
                                                          int B, int H, int W, int C, int sp_in, int Ho, int Wo, int sp_out) {
  constexpr int N = CVec<T>::N;
  const int CV = C / N;
  const int Hi = H + 2 * sp_in, Wi = W + 2 * sp_in, Hs = Ho + 2 * sp_out, Ws = Wo + 2 * sp_out;
  const size_t total = (size_t)B * Hs * Ws * CV;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    size_t r = t / CV;
    const int sx = (int)(r % Ws);
    r /= Ws;
    const int sy = (int)(r % Hs);
    const size_t b = r / Hs;
    const int oy = sy - sp_out, ox = sx - sp_out;
    float best[N];
    uint8_t bi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { best[i] = 0.f; bi[i] = 255; }
    if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
#pragma unroll
      for (int i = 0; i < N; ++i) best[i] = -INFINITY;
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        float v[N];
        CVec<T>::load(x + ((b * Hi + iy + sp_in) * Wi + ix + sp_in) * C + (size_t)cv * N, v);
#pragma unroll
        for (int i = 0; i < N; ++i)
          if (v[i] > best[i] || v[i] != v[i]) { best[i] = v[i]; bi[i] = (uint8_t)tap; }
      }
    }
    CVec<T>::store(y + t * N, best);
#pragma unroll
    for (int i = 0; i < N; ++i) idx[t * N + i] = bi[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T *__restrict__ gy, const uint8_t *__restrict__ idx, T *__restrict__ gx,
                                                          int B, int H, int W, int C, int sp_in, int Ho, int Wo, int sp_out) {
  constexpr int N = CVec<T>::N;
  const int CV = C / N;
  const int Hi = H + 2 * sp_in, Wi = W + 2 * sp_in, Hs = Ho + 2 * sp_out, Ws = Wo + 2 * sp_out;
  const size_t total = (size_t)B * Hi * Wi * CV;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    size_t r = t / CV;
    const int sx = (int)(r % Wi);
    r /= Wi;
    const int sy = (int)(r % Hi);
    const size_t b = r / Hi;
    const int iy = sy - sp_in, ix = sx - sp_in;
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      for (int ky = 0; ky < 3; ++ky) {
        const int ny = iy + 1 - ky;
        if (ny < 0 || (ny & 1)) continue;
        const int oy = ny >> 1;
        if (oy >= Ho) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int nx = ix + 1 - kx;
          if (nx < 0 || (nx & 1)) continue;
          const int ox = nx >> 1;
          if (ox >= Wo) continue;
          const size_t o = (((b * Hs + oy + sp_out) * Ws + ox + sp_out) * C) + (size_t)cv * N;
          float v[N];
          CVec<T>::load(gy + o, v);
          const uint8_t tap = (uint8_t)(ky * 3 + kx);
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (idx[o + i] == tap) acc[i] += v[i];
        }
      }
    }
    CVec<T>::store(gx + t * N, acc);
  }
}

// ---- segmentation head: nn.Upsample(x2, bilinear, align_corners=True) -> 3x3 conv to n_classes <= 4 -----------------
// (decoders.py:125-129 -- the last resize of decoder_4 followed by final_out.)  Both are linear and the resize acts
// per channel, so  conv(up(z)) = sum_tap shift_tap( up( z . w_tap ) ):  the 64 -> n_classes contraction is done FIRST, at
// the LOW resolution (U[pixel, tap, class], 18 numbers per pixel for 2 classes), and the high-resolution pass only
// interpolates and adds 9 shifted taps.  The 12 x 672 x 672 x 64 upsampled activation (694 MB in bf16) and its gradient
// are never materialised; the contraction runs on a quarter of the pixels.
struct HeadSrc {
  int i0, i1;
  float l;
};
__device__ __forceinline__ HeadSrc head_src(int o, float scale, int in) {
  const float sf = scale * (float)o;
  HeadSrc r;
  r.i0 = min((int)sf, in - 1);
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l = sf - (float)r.i0;
  return r;
}
__device__ __forceinline__ float head_weight(int o, int i, float scale, int in) {
  const HeadSrc sr = head_src(o, scale, in);
  return (sr.i0 == i ? 1.f - sr.l : 0.f) + (sr.i1 == i ? sr.l : 0.f);
}

// U[p, t] = sum_c z[p, c] * w2[t, c];  one thread per low-resolution pixel, weights broadcast from shared memory
template <typename T, int NT>
__global__ void __launch_bounds__(128) head_project_fwd_kernel(const T *__restrict__ z, const float *__restrict__ w2, float *__restrict__ U,
                                                               size_t R, int C) {
  constexpr int N = CVec<T>::N;
  extern __shared__ float sw[];          // [NT][C]
  for (int i = threadIdx.x; i < NT * C; i += blockDim.x) sw[i] = w2[i];
  __syncthreads();
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < R; p += (size_t)gridDim.x * blockDim.x) {
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    for (int c = 0; c < C; c += N) {
      float v[N];
      CVec<T>::load(z + p * C + c, v);
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < N; ++i) acc[t] = fmaf(v[i], sw[t * C + c + i], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) U[p * NT + t] = acc[t];
  }
}

// y[b, oy, ox, co] = bias[co] + sum_tap up(U[.., tap, co])(oy - 1 + ky, ox - 1 + kx)   (zero outside the 2H x 2W map)
template <int CO>
__global__ void __launch_bounds__(256) head_gather_fwd_kernel(const float *__restrict__ U, const float *__restrict__ bias, float *__restrict__ y,
                                                              int B, int H, int W, float sy, float sx) {
  constexpr int NT = 9 * CO;
  const int OH = 2 * H, OW = 2 * W;
  const size_t total = (size_t)B * OH * OW;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(t % OW);
    const size_t q = t / OW;
    const int oy = (int)(q % OH);
    const size_t b = q / OH;
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = bias ? bias[co] : 0.f;
    const float *Ub = U + b * H * W * NT;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = oy - 1 + ky;
      if (yy < 0 || yy >= OH) continue;
      const HeadSrc ys = head_src(yy, sy, H);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = ox - 1 + kx;
        if (xx < 0 || xx >= OW) continue;
        const HeadSrc xs = head_src(xx, sx, W);
        const int tap = ky * 3 + kx;
        const float w00 = (1.f - ys.l) * (1.f - xs.l), w01 = (1.f - ys.l) * xs.l, w10 = ys.l * (1.f - xs.l), w11 = ys.l * xs.l;
        const float *u00 = Ub + ((size_t)ys.i0 * W + xs.i0) * NT + tap * CO, *u01 = Ub + ((size_t)ys.i0 * W + xs.i1) * NT + tap * CO;
        const float *u10 = Ub + ((size_t)ys.i1 * W + xs.i0) * NT + tap * CO, *u11 = Ub + ((size_t)ys.i1 * W + xs.i1) * NT + tap * CO;
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co] += w00 * u00[co] + w01 * u01[co] + w10 * u10[co] + w11 * u11[co];
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) y[t * CO + co] = acc[co];
  }
}

// dU[b, iy, ix, tap, co] = sum over the upsampled positions (uy, ux) that read (iy, ix) of
//                          wy(uy) * wx(ux) * gy[b, uy + 1 - ky, ux + 1 - kx, co]        (gather, no atomics)
// One thread per low-resolution pixel, all nine taps at once, separably: the upsampled rows / columns that read source
// index i lie in [2i - 2, 2i + 3]; for every logit row oy of the 8-row window the three column sums
// tx[kx] = sum_ox wx(ox - 1 + kx) g[oy, ox] are formed once and added to the three taps ky with weight wy(oy - 1 + ky).
// (The first version looped over an 8 x 8 window per (pixel, tap) and re-derived the interpolation weights inside:
// 5.4 ms; this one is ~0.2 ms.)
template <int CO>
__global__ void __launch_bounds__(128) head_gather_bwd_kernel(const float *__restrict__ gy, float *__restrict__ dU, int B, int H, int W,
                                                              float sy, float sx) {
  constexpr int NT = 9 * CO;
  const int OH = 2 * H, OW = 2 * W;
  const size_t total = (size_t)B * H * W;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(t % W);
    const size_t q = t / W;
    const int iy = (int)(q % H);
    const size_t b = q / H;
    // interpolation weights of the upsampled rows u = 2i - 4 + j, j = 0..9 (zero outside the map / when i is not read)
    float wyw[10], wxw[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int uy = 2 * iy - 4 + j, ux = 2 * ix - 4 + j;
      wyw[j] = (uy >= 0 && uy < OH) ? head_weight(uy, iy, sy, H) : 0.f;
      wxw[j] = (ux >= 0 && ux < OW) ? head_weight(ux, ix, sx, W) : 0.f;
    }
    float acc[9][CO];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int co = 0; co < CO; ++co) acc[k][co] = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {                  // logit row oy = 2 iy - 3 + r; its tap ky reads upsampled row oy - 1 + ky
      const int oy = 2 * iy - 3 + r;
      if (oy < 0 || oy >= OH) continue;
      float tx[3][CO];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int co = 0; co < CO; ++co) tx[kx][co] = 0.f;
      const float *grow = gy + ((b * OH + oy) * OW) * CO;
#pragma unroll
      for (int c = 0; c < 8; ++c) {                // logit column ox = 2 ix - 3 + c; tap kx reads upsampled column ox - 1 + kx
        const int ox = 2 * ix - 3 + c;
        if (ox < 0 || ox >= OW) continue;
        float g[CO];
#pragma unroll
        for (int co = 0; co < CO; ++co) g[co] = grow[(size_t)ox * CO + co];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float w = wxw[c + kx];             // u = ox - 1 + kx = 2 ix - 4 + (c + kx)
#pragma unroll
          for (int co = 0; co < CO; ++co) tx[kx][co] = fmaf(w, g[co], tx[kx][co]);
        }
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float w = wyw[r + ky];               // u = oy - 1 + ky = 2 iy - 4 + (r + ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int co = 0; co < CO; ++co) acc[ky * 3 + kx][co] = fmaf(w, tx[kx][co], acc[ky * 3 + kx][co]);
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int co = 0; co < CO; ++co) dU[t * NT + k * CO + co] = acc[k][co];
  }
}

// dz[p, c] = sum_t dU[p, t] * w2[t, c]
template <typename T, int NT>
__global__ void __launch_bounds__(128) head_project_dgrad_kernel(const float *__restrict__ dU, const float *__restrict__ w2, T *__restrict__ dz,
                                                                 size_t R, int C) {
  constexpr int N = CVec<T>::N;
  extern __shared__ float sw[];          // [NT][C]
  for (int i = threadIdx.x; i < NT * C; i += blockDim.x) sw[i] = w2[i];
  __syncthreads();
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < R; p += (size_t)gridDim.x * blockDim.x) {
    float d[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) d[t] = dU[p * NT + t];
    for (int c = 0; c < C; c += N) {
      float o[N];
#pragma unroll
      for (int i = 0; i < N; ++i) o[i] = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = fmaf(d[t], sw[t * C + c + i], o[i]);
      CVec<T>::store(dz + p * C + c, o);
    }
  }
}

// dw2[t, c] = sum_p dU[p, t] * z[p, c]: a block owns a strip of pixels, a WARP one pixel at a time (4 in flight), a lane
// CPL = C / 32 adjacent channels with all NT tap accumulators in registers: per pixel one coalesced row load of z and NT
// warp-uniform (broadcast) loads of dU feed NT * CPL FMAs per lane.  Warps are combined through shared memory in a fixed
// order, blocks through partial[block, NT*C] + conv_reduce_partials_kernel.  (The first version gave every thread five
// scattered (t, c) outputs and two dependent loads per FMA: 5.0 ms at the reference's sizes, 3.5 % of the step.)
template <typename T, int NT, int CPL>
__global__ void __launch_bounds__(256) head_project_wgrad_kernel(const float *__restrict__ dU, const T *__restrict__ z, float *__restrict__ partial,
                                                                 size_t R, int pix_per_block) {
  constexpr int C = 32 * CPL;
  __shared__ float red[8][CPL][33];                 // one tap at a time: [warp][i][lane] (+1: conflict-free column reads)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[NT][CPL];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[t][i] = 0.f;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block, p1 = min(R, p0 + pix_per_block);
  constexpr int UN = 4;
  for (size_t pb = p0 + warp * UN; pb < p1; pb += 8 * UN) {
    float zv[UN][CPL], du[UN][NT];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const size_t p = pb + u;
      const bool ok = p < p1;
      const size_t pc = ok ? p : p0;                 // (clamped address; the contribution is zeroed below)
      const T *zp = z + pc * C + lane * CPL;
#pragma unroll
      for (int i = 0; i < CPL; ++i) zv[u][i] = ok ? to_f(zp[i]) : 0.f;
      const float *dp = dU + pc * NT;
#pragma unroll
      for (int t = 0; t < NT; ++t) du[u][t] = __ldg(dp + t);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < CPL; ++i) acc[t][i] = fmaf(du[u][t], zv[u][i], acc[t][i]);
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) red[warp][i][lane] = acc[t][i];
    __syncthreads();
    if (threadIdx.x < C) {                           // channel c = l * CPL + i, warps added in a fixed order
      const int l = threadIdx.x / CPL, i = threadIdx.x % CPL;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[w][i][l];
      partial[(size_t)blockIdx.x * (NT * C) + t * C + threadIdx.x] = a;
    }
    __syncthreads();
  }
}

// generic shape (any C): a block owns a strip of pixels, a thread up to HEAD_WG outputs (t, c), c fastest across the threads
constexpr int HEAD_WG = 20;
template <typename T>
__global__ void __launch_bounds__(256) head_project_wgrad_generic_kernel(const float *__restrict__ dU, const T *__restrict__ z,
                                                                         float *__restrict__ partial, size_t R, int C, int NT, int pix_per_block) {
  const int nout = NT * C;
  float acc[HEAD_WG];
  int oc[HEAD_WG], ot[HEAD_WG];
#pragma unroll
  for (int j = 0; j < HEAD_WG; ++j) {
    const int o = threadIdx.x + j * 256;
    acc[j] = 0.f;
    oc[j] = o < nout ? o % C : 0;
    ot[j] = o < nout ? o / C : -1;
  }
  const size_t p0 = (size_t)blockIdx.x * pix_per_block, p1 = min(R, p0 + pix_per_block);
  for (size_t p = p0; p < p1; ++p) {
#pragma unroll
    for (int j = 0; j < HEAD_WG; ++j)
      if (ot[j] >= 0) acc[j] = fmaf(dU[p * NT + ot[j]], to_f(z[p * C + oc[j]]), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < HEAD_WG; ++j) {
    const int o = threadIdx.x + j * 256;
    if (o < nout) partial[(size_t)blockIdx.x * nout + o] = acc[j];
  }
}

static unsigned grid_for(size_t total) { return (unsigned)std::min<size_t>((total + 255) / 256, (size_t)148 * 32); }

}  // namespace asis

using namespace asis;

static int conv_geom(ConvGeom &g, int B, int H, int W, int C, int sp, int k, int stride, int pad, int64_t ldk) {
  ASIS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && sp >= 0 && k >= 1 && k <= 7 && stride >= 1 && pad >= 0, "conv: bad geometry");
  g.B = B; g.H = H; g.W = W; g.C = C; g.sp = sp; g.k = k; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - k) / stride + 1;
  g.Wo = (W + 2 * pad - k) / stride + 1;
  g.ldk = ldk;
  ASIS_REQUIRE(g.Ho > 0 && g.Wo > 0, "conv: empty output");
  ASIS_REQUIRE(ldk >= (int64_t)k * k * C, "conv: ldk=%lld < k*k*C", (long long)ldk);
  return ASIS_OK;
}

extern "C" int asis_im2col(const void *x, int x_dtype, void *cols, int cols_dtype, int B, int H, int W, int C, int storage_pad,
                           int k, int stride, int pad, int64_t ldk, void *stream) {
  ASIS_REQUIRE(x && cols, "im2col: null pointer");
  ASIS_REQUIRE(dtype_ok(x_dtype) && dtype_ok(cols_dtype), "im2col: bad dtype");
  ConvGeom g;
  if (int rc = conv_geom(g, B, H, W, C, storage_pad, k, stride, pad, ldk)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int NV = cols_dtype == ASIS_BF16 ? 8 : 4;
  const bool vec = C % NV == 0 && ldk % NV == 0 && ldk == (int64_t)k * k * C && aligned16(x) && aligned16(cols) &&
                   !(x_dtype == ASIS_BF16 && cols_dtype == ASIS_F32);
  if (vec) {
    const size_t total = (size_t)B * g.Ho * g.Wo * k * k * (C / NV);
    ASIS_DISPATCH_DTYPE(x_dtype, TI, ASIS_DISPATCH_DTYPE(cols_dtype, TO, (im2col_vec_kernel<TI, TO><<<grid_for(total), 256, 0, st>>>((const TI *)x, (TO *)cols, g))));
  } else {
    const size_t total = (size_t)B * g.Ho * g.Wo * ldk;
    ASIS_DISPATCH_DTYPE(x_dtype, TI, ASIS_DISPATCH_DTYPE(cols_dtype, TO, (im2col_scalar_kernel<TI, TO><<<grid_for(total), 256, 0, st>>>((const TI *)x, (TO *)cols, g))));
  }
  ASIS_LAUNCHED();
  return ASIS_OK;
}

extern "C" int asis_col2im(const void *dcols, void *dx, int dtype, int B, int H, int W, int C, int storage_pad, int k, int stride,
                           int pad, int64_t ldk, void *stream) {
  ASIS_REQUIRE(dcols && dx, "col2im: null pointer");
  ASIS_REQUIRE(dtype_ok(dtype), "col2im: bad dtype");
  ConvGeom g;
  if (int rc = conv_geom(g, B, H, W, C, storage_pad, k, stride, pad, ldk)) return rc;
  const int NV = dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && ldk % NV == 0 && aligned16(dcols) && aligned16(dx), "col2im: C=%d and ldk must be multiples of %d", C, NV);
  const size_t total = (size_t)B * (H + 2 * storage_pad) * (W + 2 * storage_pad) * (C / NV);
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(dtype, T, (col2im_kernel<T><<<grid_for(total), 256, 0, st>>>((const T *)dcols, (T *)dx, g)));
  ASIS_LAUNCHED();
  return ASIS_OK;
}

static int stat_blocks(size_t npix, int &ppb) {
  int nb = (int)std::min<size_t>((npix + 255) / 256, 592);
  if (nb < 1) nb = 1;
  ppb = (int)((npix + nb - 1) / nb);
  return nb;
}

extern "C" size_t asis_chan_stats_workspace_bytes(int B, int H, int W, int C) {
  int ppb;
  const int nb = stat_blocks((size_t)B * H * W, ppb);
  return (size_t)nb * 2 * C * sizeof(float);            // one partial row per block
}

// channel-vector lanes of a block: the largest power of two <= min(32, CV), at least 4
static int stat_lanes(int CV) {
  int cl = 32;
  while (cl > 4 && cl > CV) cl >>= 1;
  return cl;
}

// mode 0: s1 = sum (x - shift), s2 = sum (x - shift)^2.   mode 1: s1 = sum dz, s2 = sum dz * xhat (see StatArgs).
extern "C" int asis_chan_stats(int mode, const void *x, const void *dy, int dtype, int B, int H, int W, int C, int storage_pad,
                               int dy_pad, const float *shift, const float *a, const float *b, const float *mean, const float *rstd, int relu,
                               float *s1, float *s2, void *workspace, size_t workspace_bytes, void *stream) {
  ASIS_REQUIRE(x && s1 && s2 && workspace, "chan_stats: null pointer");
  ASIS_REQUIRE(dtype_ok(dtype), "chan_stats: bad dtype");
  ASIS_REQUIRE(mode == 0 ? shift != nullptr : (dy && a && b && mean && rstd), "chan_stats: missing per-channel vectors for mode %d", mode);
  const int NV = dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && aligned16(x), "chan_stats: C=%d must be a multiple of %d", C, NV);
  ASIS_REQUIRE(s2 == s1 + C, "chan_stats: s2 must follow s1 (one [2C] buffer)");
  const size_t need = asis_chan_stats_workspace_bytes(B, H, W, C);
  if (workspace_bytes < need) ASIS_FAIL(ASIS_ERR_WORKSPACE, "chan_stats: workspace %zu < %zu bytes", workspace_bytes, need);
  int ppb;
  const int nb = stat_blocks((size_t)B * H * W, ppb);
  StatArgs s{B, H, W, C, storage_pad, dy_pad, shift, a, b, mean, rstd, relu};
  const int CV = C / NV;
  const int cl = stat_lanes(CV);
  dim3 grid(nb, (CV + cl - 1) / cl > 8 ? 8 : (CV + cl - 1) / cl);
  cudaStream_t st = (cudaStream_t)stream;
  float *partial = (float *)workspace;
  if (mode == 0) {
    ASIS_DISPATCH_DTYPE(dtype, T, (chan_stats_kernel<T, 0><<<grid, 256, 0, st>>>((const T *)x, nullptr, partial, s, ppb, cl)));
  } else {
    ASIS_DISPATCH_DTYPE(dtype, T, (chan_stats_kernel<T, 1><<<grid, 256, 0, st>>>((const T *)x, (const T *)dy, partial, s, ppb, cl)));
  }
  ASIS_LAUNCHED();
  conv_reduce_partials_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(partial, nb, 2 * C, s1);
  ASIS_LAUNCHED();
  return ASIS_OK;
}

// mode 0: y = act(a (x - mean) + b).  mode 1: dx = a (dz - c1 - xhat c2)  (dx in storage pad_out, dy in storage dy_pad).
extern "C" int asis_bn_apply(int mode, const void *x, const void *dy, int in_dtype, void *out, int out_dtype, int B, int H, int W, int C,
                             int pad_in, int pad_out, int dy_pad, const float *a, const float *b, const float *mean,
                             const float *rstd, const float *c1, const float *c2, int relu, void *stream) {
  ASIS_REQUIRE(x && out && a && b && mean, "bn_apply: null pointer");
  ASIS_REQUIRE(dtype_ok(in_dtype) && dtype_ok(out_dtype), "bn_apply: bad dtype");
  ASIS_REQUIRE(mode == 0 || (dy && mean && rstd && c1 && c2), "bn_apply: backward needs dy, mean, rstd, c1, c2");
  const int NV = out_dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && aligned16(x) && aligned16(out), "bn_apply: C=%d must be a multiple of %d", C, NV);
  ASIS_REQUIRE(!(in_dtype == ASIS_BF16 && out_dtype == ASIS_F32), "bn_apply: bf16 -> f32 is not provided");
  ApplyArgs s{B, H, W, C, pad_in, pad_out, dy_pad, a, b, mean, rstd, c1, c2, relu};
  const size_t total = (size_t)B * (H + 2 * pad_out) * (W + 2 * pad_out) * (C / NV);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    ASIS_DISPATCH_DTYPE(in_dtype, TI, ASIS_DISPATCH_DTYPE(out_dtype, TO, (bn_apply_kernel<TI, TO, 0><<<grid_for(total), 256, 0, st>>>((const TI *)x, nullptr, (TO *)out, s))));
  } else {
    ASIS_DISPATCH_DTYPE(in_dtype, TI, ASIS_DISPATCH_DTYPE(out_dtype, TO, (bn_apply_kernel<TI, TO, 1><<<grid_for(total), 256, 0, st>>>((const TI *)x, (const TI *)dy, (TO *)out, s))));
  }
  ASIS_LAUNCHED();
  return ASIS_OK;
}

extern "C" int asis_maxpool3x3s2_forward(const void *x, void *y, uint8_t *idx, int dtype, int B, int H, int W, int C, int pad_in,
                                         int pad_out, void *stream) {
  ASIS_REQUIRE(x && y && idx && dtype_ok(dtype), "maxpool: bad arguments");
  const int NV = dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && aligned16(x) && aligned16(y), "maxpool: C=%d must be a multiple of %d", C, NV);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * (Ho + 2 * pad_out) * (Wo + 2 * pad_out) * (C / NV);
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(dtype, T, (maxpool_fwd_kernel<T><<<grid_for(total), 256, 0, st>>>((const T *)x, (T *)y, idx, B, H, W, C, pad_in, Ho, Wo, pad_out)));
  ASIS_LAUNCHED();
  return ASIS_OK;
}

extern "C" int asis_maxpool3x3s2_backward(const void *gy, const uint8_t *idx, void *gx, int dtype, int B, int H, int W, int C, int pad_in,
                                          int pad_out, void *stream) {
  ASIS_REQUIRE(gy && gx && idx && dtype_ok(dtype), "maxpool: bad arguments");
  const int NV = dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && aligned16(gy) && aligned16(gx), "maxpool: C=%d must be a multiple of %d", C, NV);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * (H + 2 * pad_in) * (W + 2 * pad_in) * (C / NV);
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(dtype, T, (maxpool_bwd_kernel<T><<<grid_for(total), 256, 0, st>>>((const T *)gy, idx, (T *)gx, B, H, W, C, pad_in, Ho, Wo, pad_out)));
  ASIS_LAUNCHED();
  return ASIS_OK;
}

static int head_check(int dtype, int C, int CO) {
  ASIS_REQUIRE(dtype_ok(dtype), "seg_head: bad dtype");
  const int NV = dtype == ASIS_BF16 ? 8 : 4;
  ASIS_REQUIRE(C % NV == 0 && C >= NV, "seg_head: C=%d must be a multiple of %d", C, NV);
  ASIS_REQUIRE(CO >= 1 && CO <= 4, "seg_head: 1..4 classes (got %d); wider heads go through asis_upsample2x + im2col + gemm", CO);
  ASIS_REQUIRE(9 * CO * C <= HEAD_WG * 256 && (size_t)9 * CO * C * sizeof(float) <= 160 * 1024, "seg_head: C * classes too large");
  return ASIS_OK;
}

#define ASIS_HEAD_CO(CO, ...)           \
  do {                                  \
    switch (CO) {                       \
      case 1: { constexpr int CO_ = 1; __VA_ARGS__; } break; \
      case 2: { constexpr int CO_ = 2; __VA_ARGS__; } break; \
      case 3: { constexpr int CO_ = 3; __VA_ARGS__; } break; \
      default: { constexpr int CO_ = 4; __VA_ARGS__; } break; \
    }                                   \
  } while (0)

extern "C" size_t asis_seg_head_workspace_bytes(int B, int H, int W, int C, int CO) {
  // U / dU [B*H*W, 9*CO] f32, followed by the weight-gradient partials of the backward
  const size_t R = (size_t)B * H * W;
  const size_t nb = std::min<size_t>((R + 1023) / 1024, 592);
  return align_up(R * 9 * CO * sizeof(float), 256) + nb * 9 * CO * C * sizeof(float);
}

template <typename T, int CO>
static int seg_head_fwd(const void *z, const float *w2, const float *bias, float *y, float *U, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int NT = 9 * CO;
  const size_t R = (size_t)B * H * W;
  const size_t shb = (size_t)NT * C * sizeof(float);
  static bool configured = false;
  if (!configured) {
    ASIS_CUDA(cudaFuncSetAttribute(head_project_fwd_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  const unsigned pb = (unsigned)std::min<size_t>((R + 127) / 128, (size_t)148 * 16);
  head_project_fwd_kernel<T, NT><<<pb, 128, shb, st>>>((const T *)z, w2, U, R, C);
  ASIS_LAUNCHED();
  const float sy = (float)(H - 1) / (float)(2 * H - 1), sx = (float)(W - 1) / (float)(2 * W - 1);
  head_gather_fwd_kernel<CO><<<grid_for(R * 4), 256, 0, st>>>(U, bias, y, B, H, W, sy, sx);
  ASIS_LAUNCHED();
  return ASIS_OK;
}

// y [B, 2H, 2W, CO] f32 = conv3x3_pad1( upsample2x_bilinear_align_corners(z [B, H, W, C]) ) + bias;
// w2 [9*CO, C] f32 with row index (ky*3 + kx)*CO + co.
extern "C" int asis_seg_head_forward(const void *z, int dtype, const float *w2, const float *bias, float *y, int B, int H, int W, int C,
                                     int CO, void *workspace, size_t workspace_bytes, void *stream) {
  ASIS_REQUIRE(z && w2 && y && workspace, "seg_head: null pointer");
  ASIS_REQUIRE(B > 0 && H > 1 && W > 1, "seg_head: need B > 0, H > 1, W > 1");
  if (int rc = head_check(dtype, C, CO)) return rc;
  if (workspace_bytes < asis_seg_head_workspace_bytes(B, H, W, C, CO)) ASIS_FAIL(ASIS_ERR_WORKSPACE, "seg_head: workspace too small");
  int rc = ASIS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(dtype, T, ASIS_HEAD_CO(CO, rc = (seg_head_fwd<T, CO_>(z, w2, bias, y, (float *)workspace, B, H, W, C, st))));
  return rc;
}

template <typename T, int CO>
static int seg_head_bwd(const void *z, const float *w2, const float *gy, void *gz, float *gw2, float *dU, float *partial, int B, int H,
                        int W, int C, cudaStream_t st) {
  constexpr int NT = 9 * CO;
  const size_t R = (size_t)B * H * W;
  const float sy = (float)(H - 1) / (float)(2 * H - 1), sx = (float)(W - 1) / (float)(2 * W - 1);
  head_gather_bwd_kernel<CO><<<(unsigned)std::min<size_t>((R + 127) / 128, (size_t)148 * 32), 128, 0, st>>>(gy, dU, B, H, W, sy, sx);
  ASIS_LAUNCHED();
  if (gz) {
    const size_t shb = (size_t)NT * C * sizeof(float);
    static bool configured = false;
    if (!configured) {
      ASIS_CUDA(cudaFuncSetAttribute(head_project_dgrad_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured = true;
    }
    const unsigned pb = (unsigned)std::min<size_t>((R + 127) / 128, (size_t)148 * 16);
    head_project_dgrad_kernel<T, NT><<<pb, 128, shb, st>>>(dU, w2, (T *)gz, R, C);
    ASIS_LAUNCHED();
  }
  if (gw2) {
    const int nb = (int)std::min<size_t>((R + 1023) / 1024, 592);
    const int ppb = (int)((R + nb - 1) / nb);
    if (C == 64) head_project_wgrad_kernel<T, NT, 2><<<nb, 256, 0, st>>>(dU, (const T *)z, partial, R, ppb);
    else if (C == 128) head_project_wgrad_kernel<T, NT, 4><<<nb, 256, 0, st>>>(dU, (const T *)z, partial, R, ppb);
    else head_project_wgrad_generic_kernel<T><<<nb, 256, 0, st>>>(dU, (const T *)z, partial, R, C, NT, ppb);
    ASIS_LAUNCHED();
    conv_reduce_partials_kernel<<<(NT * C + 31) / 32, 256, 0, st>>>(partial, nb, NT * C, gw2);
    ASIS_LAUNCHED();
  }
  return ASIS_OK;
}

// gz [B, H, W, C] (z's dtype, optional), gw2 [9*CO, C] f32 (optional); the bias gradient is the sum of gy over pixels
extern "C" int asis_seg_head_backward(const void *z, int dtype, const float *w2, const float *gy, void *gz, float *gw2, int B, int H,
                                      int W, int C, int CO, void *workspace, size_t workspace_bytes, void *stream) {
  ASIS_REQUIRE(z && w2 && gy && workspace, "seg_head: null pointer");
  if (int rc = head_check(dtype, C, CO)) return rc;
  if (workspace_bytes < asis_seg_head_workspace_bytes(B, H, W, C, CO)) ASIS_FAIL(ASIS_ERR_WORKSPACE, "seg_head: workspace too small");
  const size_t R = (size_t)B * H * W;
  float *dU = (float *)workspace;
  float *partial = (float *)((char *)workspace + align_up(R * 9 * CO * sizeof(float), 256));
  int rc = ASIS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ASIS_DISPATCH_DTYPE(dtype, T, ASIS_HEAD_CO(CO, rc = (seg_head_bwd<T, CO_>(z, w2, gy, gz, gw2, dU, partial, B, H, W, C, st))));
  return rc;
}
