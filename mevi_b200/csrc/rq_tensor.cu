// rq_tensor.cu — tcgen05 split-fp16 prefilter path (placeholder until the kernel lands).
#include "common.cuh"

bool mevi_rq_tensor_supported(mevi_ctx* ctx, int d, int M, int K, int metric) { return false; }

int mevi_rq_tensor_assign(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                          int32_t* codes, int64_t codes_stride, float* residual, int64_t* stats, double* inertia,
                          cudaStream_t st) {
  return mevi_set_error(ctx, MEVI_ERR_UNSUPPORTED, "tensor path not built");
}
