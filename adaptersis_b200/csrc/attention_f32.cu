// attention_f32.cu -- fp32 parity-mode self-attention over the packed qkv tensor
// (dinov2/layers/attention.py:56-69): batched FFMA GEMMs on strided views of qkv plus row-wise
// softmax kernels; scores are materialised in a caller-provided workspace.  The bf16
// performance path is the flash-style tcgen05 kernel in attention_tc.cu.
#include "gemm_f32.cuh"

namespace asis {

// in place: row -> softmax(row); lse = max + log(sum)
__global__ void __launch_bounds__(256) softmax_rows_kernel(float *__restrict__ S, float *__restrict__ lse, int T) {
  __shared__ float red[8];
  float *row = S + (size_t)blockIdx.x * T;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < T; i += blockDim.x) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) sum += expf(row[i] - mx);
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < T; i += blockDim.x) row[i] = expf(row[i] - mx) * inv;
  if (threadIdx.x == 0) lse[blockIdx.x] = mx + logf(sum);
}

// in place: S -> P = exp(S - lse[row])
__global__ void __launch_bounds__(256) probs_from_lse_kernel(float *__restrict__ S, const float *__restrict__ lse, int T) {
  float *row = S + (size_t)blockIdx.x * T;
  const float l = lse[blockIdx.x];
  for (int i = threadIdx.x; i < T; i += blockDim.x) row[i] = expf(row[i] - l);
}

// in place on dP: dS = P * (dP - D), D = sum_e dO[t,e] * O[t,e]; row index = (b*H + h)*T + t
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float *__restrict__ Pm, float *__restrict__ dP,
                                                               const float *__restrict__ dO, const float *__restrict__ O,
                                                               int T, int H, int hd) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const int t = r % T, h = (r / T) % H, b = r / (T * H);
  const size_t o = ((size_t)b * T + t) * H * hd + (size_t)h * hd;
  float d = 0.f;
  for (int e = threadIdx.x; e < hd; e += blockDim.x) d += dO[o + e] * O[o + e];
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  d = 0.f;
  for (int w = 0; w < 8; ++w) d += red[w];
  const float *p = Pm + (size_t)r * T;
  float *dp = dP + (size_t)r * T;
  for (int i = threadIdx.x; i < T; i += blockDim.x) dp[i] = p[i] * (dp[i] - d);
}

static GemmF32 batched(int B, int H) {
  GemmF32 g{};
  g.nb1 = B;
  g.nb2 = H;
  g.alpha = 1.f;
  g.epi.kind = -1;
  return g;
}

size_t attention_f32_fwd_ws(int B, int T, int H) { return (size_t)B * H * T * T * sizeof(float); }
size_t attention_f32_bwd_ws(int B, int T, int H) { return 2 * (size_t)B * H * T * T * sizeof(float); }

int attention_f32_forward(const float *qkv, float *out, float *lse, int B, int T, int H, int hd, float *ws,
                          cudaStream_t st) {
  const int64_t C = (int64_t)H * hd, C3 = 3 * C;
  const float scale = 1.0f / sqrtf((float)hd);
  float *S = ws;
  GemmF32 g = batched(B, H);
  // S = scale * Q K^T
  g.A = qkv; g.a_m = C3; g.a_k = 1; g.a_b1 = T * C3; g.a_b2 = hd;
  g.B = qkv + C; g.b_n = C3; g.b_k = 1; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = S; g.c_m = T; g.c_b1 = (int64_t)H * T * T; g.c_b2 = (int64_t)T * T;
  g.M = T; g.N = T; g.K = hd; g.alpha = scale;
  if (int rc = gemm_f32_launch(g, st)) return rc;
  softmax_rows_kernel<<<B * H * T, 256, 0, st>>>(S, lse, T);
  ASIS_LAUNCHED();
  // O = P V
  g = batched(B, H);
  g.A = S; g.a_m = T; g.a_k = 1; g.a_b1 = (int64_t)H * T * T; g.a_b2 = (int64_t)T * T;
  g.B = qkv + 2 * C; g.b_n = 1; g.b_k = C3; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = out; g.c_m = C; g.c_b1 = T * C; g.c_b2 = hd;
  g.M = T; g.N = hd; g.K = T;
  return gemm_f32_launch(g, st);
}

int attention_f32_backward(const float *qkv, const float *out, const float *lse, const float *dout, float *dqkv,
                           int B, int T, int H, int hd, float *ws, cudaStream_t st) {
  const int64_t C = (int64_t)H * hd, C3 = 3 * C;
  const int64_t TT = (int64_t)T * T;
  const float scale = 1.0f / sqrtf((float)hd);
  float *Pm = ws;
  float *dP = ws + (size_t)B * H * TT;
  GemmF32 g = batched(B, H);
  // P = exp(scale * Q K^T - lse)
  g.A = qkv; g.a_m = C3; g.a_k = 1; g.a_b1 = T * C3; g.a_b2 = hd;
  g.B = qkv + C; g.b_n = C3; g.b_k = 1; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = Pm; g.c_m = T; g.c_b1 = H * TT; g.c_b2 = TT;
  g.M = T; g.N = T; g.K = hd; g.alpha = scale;
  if (int rc = gemm_f32_launch(g, st)) return rc;
  probs_from_lse_kernel<<<B * H * T, 256, 0, st>>>(Pm, lse, T);
  ASIS_LAUNCHED();
  // dV = P^T dO
  g = batched(B, H);
  g.A = Pm; g.a_m = 1; g.a_k = T; g.a_b1 = H * TT; g.a_b2 = TT;
  g.B = dout; g.b_n = 1; g.b_k = C; g.b_b1 = T * C; g.b_b2 = hd;
  g.C = dqkv + 2 * C; g.c_m = C3; g.c_b1 = T * C3; g.c_b2 = hd;
  g.M = T; g.N = hd; g.K = T;
  if (int rc = gemm_f32_launch(g, st)) return rc;
  // dP = dO V^T
  g = batched(B, H);
  g.A = dout; g.a_m = C; g.a_k = 1; g.a_b1 = T * C; g.a_b2 = hd;
  g.B = qkv + 2 * C; g.b_n = C3; g.b_k = 1; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = dP; g.c_m = T; g.c_b1 = H * TT; g.c_b2 = TT;
  g.M = T; g.N = T; g.K = hd;
  if (int rc = gemm_f32_launch(g, st)) return rc;
  softmax_bwd_rows_kernel<<<B * H * T, 256, 0, st>>>(Pm, dP, dout, out, T, H, hd);
  ASIS_LAUNCHED();
  // dQ = scale * dS K
  g = batched(B, H);
  g.A = dP; g.a_m = T; g.a_k = 1; g.a_b1 = H * TT; g.a_b2 = TT;
  g.B = qkv + C; g.b_n = 1; g.b_k = C3; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = dqkv; g.c_m = C3; g.c_b1 = T * C3; g.c_b2 = hd;
  g.M = T; g.N = hd; g.K = T; g.alpha = scale;
  if (int rc = gemm_f32_launch(g, st)) return rc;
  // dK = scale * dS^T Q
  g = batched(B, H);
  g.A = dP; g.a_m = 1; g.a_k = T; g.a_b1 = H * TT; g.a_b2 = TT;
  g.B = qkv; g.b_n = 1; g.b_k = C3; g.b_b1 = T * C3; g.b_b2 = hd;
  g.C = dqkv + C; g.c_m = C3; g.c_b1 = T * C3; g.c_b2 = hd;
  g.M = T; g.N = hd; g.K = T; g.alpha = scale;
  return gemm_f32_launch(g, st);
}

}  // namespace asis
