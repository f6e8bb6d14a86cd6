// flat_tensor.cu — tcgen05 tile kernel for the flat search (placeholder until the kernel lands).
#include "common.cuh"

bool mevi_flat_tensor_supported(mevi_ctx* ctx, int d, int k) { return false; }

int mevi_flat_tensor_tiles(mevi_ctx* ctx, const float* Q, int nq, const float* D, int64_t n_begin, int64_t n_end, int d,
                           float* tau, int* count, float* cand_score, int32_t* cand_id, int* overflow, int capg,
                           cudaStream_t st) {
  return mevi_set_error(ctx, MEVI_ERR_UNSUPPORTED, "tensor flat path not built");
}
