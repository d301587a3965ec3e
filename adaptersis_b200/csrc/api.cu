// api.cu -- library-level entry points and the compute-mode dispatchers of the C ABI
// (include/asis_b200.h).
#include "gemm_f32.cuh"

namespace asis {

std::atomic<uint64_t> g_launches{0};

char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

// tcgen05 paths (gemm_tc.cu / attention_tc.cu)
int gemm_tc_launch(const void *A, int a_major, int64_t lda, const void *B, int b_major, int64_t ldb, int M, int N,
                   int K, const EpiArgs &epi, cudaStream_t st);
int gemm_tc_conv3x3(int op, const void *a, const void *b, void *c, int c_dtype, const float *bias, int B, int H, int W, int Cin,
                    int Cout, cudaStream_t st);
int attention_tc_forward(const void *qkv, void *out, float *lse, int B, int T, int H, int hd, cudaStream_t st);
size_t attention_tc_bwd_ws(int B, int T, int H, int hd);
int attention_tc_backward(const void *qkv, const void *out, const float *lse, const void *dout, void *dqkv, int B,
                          int T, int H, int hd, void *ws, cudaStream_t st);

// fp32 paths (attention_f32.cu)
size_t attention_f32_fwd_ws(int B, int T, int H);
size_t attention_f32_bwd_ws(int B, int T, int H);
int attention_f32_forward(const float *qkv, float *out, float *lse, int B, int T, int H, int hd, float *ws,
                          cudaStream_t st);
int attention_f32_backward(const float *qkv, const float *out, const float *lse, const float *dout, float *dqkv,
                           int B, int T, int H, int hd, float *ws, cudaStream_t st);

}  // namespace asis

using namespace asis;

extern "C" int asis_abi_version(void) { return ASIS_ABI_VERSION; }
extern "C" const char *asis_last_error(void) { return err_buf(); }
extern "C" uint64_t asis_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int asis_gemm(int compute, const void *A, int a_major, int64_t lda, const void *B, int b_major, int64_t ldb,
                         void *C, int c_dtype, int64_t ldc, int M, int N, int K, int epilogue, const float *bias,
                         const float *gamma, const float *residual, void *aux, int aux_dtype, int64_t ldaux,
                         void *stream) {
  ASIS_REQUIRE(A && B && C, "gemm: null pointer");
  ASIS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: non-positive dimension (M=%d N=%d K=%d)", M, N, K);
  ASIS_REQUIRE(dtype_ok(compute) && dtype_ok(c_dtype), "gemm: bad dtype");
  ASIS_REQUIRE((a_major == ASIS_MAJOR_K || a_major == ASIS_MAJOR_MN) && (b_major == ASIS_MAJOR_K || b_major == ASIS_MAJOR_MN), "gemm: bad operand major");
  ASIS_REQUIRE(lda >= (a_major == ASIS_MAJOR_K ? K : M) && ldb >= (b_major == ASIS_MAJOR_K ? K : N) && ldc >= N, "gemm: leading dimension too small");
  if (int rc = check_epilogue(epilogue, gamma, residual, aux, aux_dtype, c_dtype)) return rc;
  if (aux) ASIS_REQUIRE(ldaux >= N, "gemm: ldaux too small");
  EpiArgs e{epilogue, bias, gamma, residual, aux, aux_dtype, ldaux, C, c_dtype, ldc};
  cudaStream_t st = (cudaStream_t)stream;
  if (compute == ASIS_F32) {
    GemmF32 g{};
    g.A = (const float *)A;
    g.a_m = a_major == ASIS_MAJOR_K ? lda : 1;
    g.a_k = a_major == ASIS_MAJOR_K ? 1 : lda;
    g.B = (const float *)B;
    g.b_n = b_major == ASIS_MAJOR_K ? ldb : 1;
    g.b_k = b_major == ASIS_MAJOR_K ? 1 : ldb;
    g.M = M; g.N = N; g.K = K; g.nb1 = 1; g.nb2 = 1; g.alpha = 1.f;
    g.epi = e;
    return gemm_f32_launch(g, st);
  }
  return gemm_tc_launch(A, a_major, lda, B, b_major, ldb, M, N, K, e, st);
}

extern "C" int asis_conv3x3s1_gemm(int op, const void *a, const void *b, void *c, int c_dtype, const float *bias, int B, int H,
                                   int W, int Cin, int Cout, void *stream) {
  ASIS_REQUIRE(a && b && c, "conv3x3s1_gemm: null pointer");
  ASIS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dtype_ok(c_dtype), "conv3x3s1_gemm: bad arguments");
  return gemm_tc_conv3x3(op, a, b, c, c_dtype, bias, B, H, W, Cin, Cout, (cudaStream_t)stream);
}

extern "C" size_t asis_attention_forward_workspace_bytes(int compute, int B, int T, int H, int hd) {
  (void)hd;
  return compute == ASIS_F32 ? attention_f32_fwd_ws(B, T, H) : 0;
}

extern "C" size_t asis_attention_backward_workspace_bytes(int compute, int B, int T, int H, int hd) {
  return compute == ASIS_F32 ? attention_f32_bwd_ws(B, T, H) : attention_tc_bwd_ws(B, T, H, hd);
}

static int check_attn(int compute, int B, int T, int H, int hd) {
  ASIS_REQUIRE(dtype_ok(compute), "attention: bad compute mode");
  ASIS_REQUIRE(B > 0 && T > 0 && H > 0 && hd > 0, "attention: non-positive dimension");
  ASIS_REQUIRE((int64_t)B * H <= 65535, "attention: B*H too large");
  if (compute == ASIS_BF16 && hd != 64) ASIS_FAIL(ASIS_ERR_UNSUPPORTED, "attention: the tcgen05 path needs head_dim 64 (got %d)", hd);
  return ASIS_OK;
}

extern "C" int asis_attention_forward(int compute, const void *qkv, void *out, float *lse, int B, int T, int H, int hd,
                                      void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = check_attn(compute, B, T, H, hd)) return rc;
  ASIS_REQUIRE(qkv && out && lse, "attention_forward: null pointer");
  const size_t need = asis_attention_forward_workspace_bytes(compute, B, T, H, hd);
  if (need && (!workspace || workspace_bytes < need)) ASIS_FAIL(ASIS_ERR_WORKSPACE, "attention_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (compute == ASIS_F32) return attention_f32_forward((const float *)qkv, (float *)out, lse, B, T, H, hd, (float *)workspace, st);
  return attention_tc_forward(qkv, out, lse, B, T, H, hd, st);
}

extern "C" int asis_attention_backward(int compute, const void *qkv, const void *out, const float *lse,
                                       const void *dout, void *dqkv, int B, int T, int H, int hd, void *workspace,
                                       size_t workspace_bytes, void *stream) {
  if (int rc = check_attn(compute, B, T, H, hd)) return rc;
  ASIS_REQUIRE(qkv && out && lse && dout && dqkv, "attention_backward: null pointer");
  const size_t need = asis_attention_backward_workspace_bytes(compute, B, T, H, hd);
  if (need && (!workspace || workspace_bytes < need)) ASIS_FAIL(ASIS_ERR_WORKSPACE, "attention_backward: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (compute == ASIS_F32) return attention_f32_backward((const float *)qkv, (const float *)out, lse, (const float *)dout, (float *)dqkv, B, T, H, hd, (float *)workspace, st);
  return attention_tc_backward(qkv, out, lse, dout, dqkv, B, T, H, hd, workspace, st);
}
