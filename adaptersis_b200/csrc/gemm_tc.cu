// temporary: tcgen05 paths not built yet
#include "epilogue.cuh"
namespace asis {
int gemm_tc_launch(const void *, int, int64_t, const void *, int, int64_t, int, int, int, const EpiArgs &, cudaStream_t) {
  ASIS_FAIL(ASIS_ERR_UNSUPPORTED, "gemm: tcgen05 path not built");
}
int attention_tc_forward(const void *, void *, float *, int, int, int, int, cudaStream_t) {
  ASIS_FAIL(ASIS_ERR_UNSUPPORTED, "attention: tcgen05 path not built");
}
size_t attention_tc_bwd_ws(int, int, int, int) { return 0; }
int attention_tc_backward(const void *, const void *, const float *, const void *, void *, int, int, int, int, void *, cudaStream_t) {
  ASIS_FAIL(ASIS_ERR_UNSUPPORTED, "attention: tcgen05 path not built");
}
}  // namespace asis
