// gemm_tc.cu -- tcgen05 (5th-gen tensor core) 3xTF32 error-compensated GEMM for the dense towers.
// Placeholder entry points: until the tcgen05 path lands they report HRB_UNSUPPORTED and
// hrb_dense_* (dense.cu) runs the fp32 FFMA kernel instead.
#include "common.cuh"

int hrb_tc_dense_fwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, int32_t, int32_t, int32_t,
                     float*, int64_t, cudaStream_t) {
  return hrb::fail(HRB_UNSUPPORTED, "tcgen05 GEMM path not built for this shape");
}
int hrb_tc_dense_bwd_x(const float*, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, const float*, int64_t,
                       int32_t, float*, int64_t, cudaStream_t) {
  return hrb::fail(HRB_UNSUPPORTED, "tcgen05 GEMM path not built for this shape");
}
int hrb_tc_dense_bwd_w(const float*, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, float*, int64_t, void*,
                       size_t, cudaStream_t) {
  return hrb::fail(HRB_UNSUPPORTED, "tcgen05 GEMM path not built for this shape");
}
