// dense.cu — the twin-tower dense scorer.
// Replaces MEVI/document_encoder.py:128-132 compute_similarity(bmm=False):
// torch.matmul(q_reps, p_reps.T) in fp32.  One warp per passage row (row held
// in registers, 128-bit coalesced loads), looped over the queries; exact FMA
// accumulation + shuffle tree.  HBM-bound for the reference's shape (one query
// x <=1024 passages, main_models.py:3948-3968).
#include "common.cuh"

namespace {
template <int NCH>
__global__ void __launch_bounds__(256) dense_scores_kernel(const float* __restrict__ Q, int nq,
                                                           const float* __restrict__ P, int64_t n, int d,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp_global; row < n; row += n_warps) {
    float4 v[NCH];
#pragma unroll
    for (int t = 0; t < NCH; ++t) {
      int c4 = (lane + 32 * t) * 4;
      v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < d) v[t] = ld_stream_f4(P + row * d + c4);
    }
    for (int q = 0; q < nq; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        int c4 = (lane + 32 * t) * 4;
        if (c4 < d) {
          float4 qq = ldg_f4(Q + (int64_t)q * d + c4);
          acc = fmaf(qq.x, v[t].x, acc);
          acc = fmaf(qq.y, v[t].y, acc);
          acc = fmaf(qq.z, v[t].z, acc);
          acc = fmaf(qq.w, v[t].w, acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) out[(int64_t)q * n + row] = acc;
    }
  }
}
}  // namespace

extern "C" int mevi_dense_scores(mevi_ctx* ctx, const float* Q, int nq, const float* P, int64_t n, int d, float* out,
                                 void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, Q && P && out, "NULL argument");
  MEVI_REQUIRE(ctx, d > 0 && d % 4 == 0 && d <= 1024, "dense scorer needs d %% 4 == 0 and d <= 1024 (got %d)", d);
  if (nq <= 0 || n <= 0) return MEVI_OK;
  int64_t want = (n + 7) / 8;
  int grid = (int)(want > (int64_t)ctx->sm_count * 8 ? (int64_t)ctx->sm_count * 8 : want);
  if (d <= 128) dense_scores_kernel<1><<<grid, 256, 0, st>>>(Q, nq, P, n, d, out);
  else if (d <= 256) dense_scores_kernel<2><<<grid, 256, 0, st>>>(Q, nq, P, n, d, out);
  else if (d <= 512) dense_scores_kernel<4><<<grid, 256, 0, st>>>(Q, nq, P, n, d, out);
  else if (d <= 768) dense_scores_kernel<6><<<grid, 256, 0, st>>>(Q, nq, P, n, d, out);
  else dense_scores_kernel<8><<<grid, 256, 0, st>>>(Q, nq, P, n, d, out);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}
