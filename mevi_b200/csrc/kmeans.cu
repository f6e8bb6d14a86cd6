// kmeans.cu — full-batch Lloyd pieces: per-centroid sum/count accumulation,
// centroid update, residual update.
//
// Reference behaviour being replaced: MEVI/pq.py:551-598 trains each RQ level
// with sklearn MiniBatchKMeans on rank 0 and subtracts centers[pred] in numpy
// (591-593).  BASELINE.json asks for the data-parallel form instead: every rank
// assigns its row block, accumulates [K,d] sums and [K] counts, and one
// all-reduce(SUM) of the fused [K*d+K] buffer (precedent: pq.py:384-397)
// gives every rank the same new centroids.
//
// Accumulation is DETERMINISTIC: a CTA walks a contiguous range of rows in
// tiles; inside a tile rows are bucketed by centroid with a stable counting
// sort, and every (centroid, column) accumulator is a register owned by one
// thread that adds the bucket's rows in ascending row order.  Per-CTA partials
// are then reduced in a fixed order.  No float atomics, so sums are
// bit-reproducible for a given (n, grid) and ranks stay in lockstep.
// Roofline: HBM (one more read of the shard: 4*d bytes per row + 4 B assignment).
#include "common.cuh"

int mevi_rq_exact_launch(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                         int32_t* codes, int64_t codes_stride, float* residual, const int32_t* work_rows,
                         const int32_t* work_levels, const int64_t* n_work_dev, int64_t n_items, double* inertia,
                         cudaStream_t st);
int mevi_rq_tensor_assign(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                          int32_t* codes, int64_t codes_stride, float* residual, int64_t* stats, double* inertia,
                          cudaStream_t st);
bool mevi_rq_tensor_supported(mevi_ctx* ctx, int d, int M, int K, int metric);

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_TILE = 256;

// Column-owner accumulation (skew-proof and deterministic).  Thread t owns columns 2t, 2t+1 of
// every centroid: KMAX x 2 fp32 accumulators in registers for the CTA's whole row range.  A tile of
// 256 rows is bucketed by centroid with a stable counting sort; then for each centroid (unrolled, so
// the accumulator is a fixed register) every thread walks the bucket's rows in ascending order and
// adds its two columns (8-byte loads, 256..3072 contiguous bytes per row across the CTA).  All threads
// work on every row, so one dominant cluster (the reference-trained codebooks put >80 % of N(0,1)
// rows into two centroids) costs nothing extra.
template <int KMAX, int NT>
__global__ void __launch_bounds__(NT, (KMAX <= 32 && NT <= 384 ? 2 : 1)) kmeans_accumulate_kernel(const float* __restrict__ R, int64_t n, int d,
                                                                const int32_t* __restrict__ assign,
                                                                int64_t assign_stride, int K,
                                                                float* __restrict__ partial_sums,     // [grid][K][d]
                                                                int32_t* __restrict__ partial_counts)  // [grid][K]
{
  __shared__ int s_assign[KM_TILE];
  __shared__ int s_hist[KMAX + 1];
  __shared__ int s_start[KMAX + 1];
  __shared__ int s_order[KM_TILE];
  __shared__ int s_count_total[KMAX];

  const int tid = threadIdx.x;
  const bool owner = 2 * tid < d;  // this thread owns columns 2*tid, 2*tid+1
  const int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = row_begin + rows_per_cta < n ? row_begin + rows_per_cta : n;

  float2 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int i = tid; i < KMAX; i += blockDim.x) s_count_total[i] = 0;
  __syncthreads();

  for (int64_t tile = row_begin; tile < row_end; tile += KM_TILE) {
    const int rows = (int)((row_end - tile) < KM_TILE ? (row_end - tile) : KM_TILE);
    for (int i = tid; i <= KMAX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    if (tid < rows) {
      int a = assign[(tile + tid) * assign_stride];
      a = a < 0 ? 0 : (a >= K ? K - 1 : a);
      s_assign[tid] = a;
      atomicAdd(&s_hist[a], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int k = 0; k < KMAX; ++k) {
        s_start[k] = run;
        run += s_hist[k];
      }
      s_start[KMAX] = run;
    }
    __syncthreads();
    // stable placement: slot = bucket start + number of earlier rows with the same assignment
    if (tid < rows) {
      const int a = s_assign[tid];
      int rank = 0;
      for (int j = 0; j < tid; ++j) rank += (s_assign[j] == a);
      s_order[s_start[a] + rank] = tid;
    }
    if (tid < KMAX) s_count_total[tid] += s_hist[tid];
    __syncthreads();
    if (owner) {
      const float* base = R + tile * d + 2 * tid;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int b = s_start[k], e = s_start[k + 1];
        int j = b;
        for (; j + 7 < e; j += 8) {  // eight rows in flight per thread
          float2 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float2*>(base + (int64_t)s_order[j + u] * d));
#pragma unroll
          for (int u = 0; u < 8; ++u) {  // ascending row order: the sum is reproducible
            acc[k].x += v[u].x;
            acc[k].y += v[u].y;
          }
        }
        for (; j < e; ++j) {
          const float2 v0 = __ldcs(reinterpret_cast<const float2*>(base + (int64_t)s_order[j] * d));
          acc[k].x += v0.x; acc[k].y += v0.y;
        }
      }
    }
    __syncthreads();
  }
  if (owner) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) *reinterpret_cast<float2*>(partial_sums + ((int64_t)blockIdx.x * K + k) * d + 2 * tid) = acc[k];
  }
  if (tid < K) partial_counts[(int64_t)blockIdx.x * K + tid] = s_count_total[tid];
}

// generic fallback (any K, d % 4 == 0): warp per row, float atomics into sums
__global__ void kmeans_accumulate_atomic_kernel(const float* __restrict__ R, int64_t n, int d,
                                                const int32_t* __restrict__ assign, int64_t assign_stride, int K,
                                                float* __restrict__ sums, float* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp_global; row < n; row += n_warps) {
    int a = assign[row * assign_stride];
    a = a < 0 ? 0 : (a >= K ? K - 1 : a);
    for (int c = lane; c < d; c += 32) atomicAdd(&sums[(int64_t)a * d + c], R[row * d + c]);
    if (lane == 0) atomicAdd(&counts[a], 1.0f);
  }
}

__global__ void kmeans_reduce_partials_kernel(const float* __restrict__ partial_sums,
                                              const int32_t* __restrict__ partial_counts, int G, int K, int d,
                                              float* __restrict__ sums_counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t kd = (int64_t)K * d;
  if (i < kd) {
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += partial_sums[(int64_t)g * kd + i];  // fixed order
    sums_counts[i] = s;
  } else if (i < kd + K) {
    const int k = (int)(i - kd);
    long long c = 0;
    for (int g = 0; g < G; ++g) c += partial_counts[(int64_t)g * K + k];
    sums_counts[i] = (float)c;
  }
}

__global__ void kmeans_update_kernel(const float* __restrict__ sums_counts, int K, int d, float* __restrict__ centroids,
                                     int32_t* __restrict__ n_empty) {
  const int64_t kd = (int64_t)K * d;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kd) {
    const int k = (int)(i / d);
    const float cnt = sums_counts[kd + k];
    if (cnt > 0.f) centroids[i] = sums_counts[i] / cnt;
  }
  if (n_empty && i < K) {
    if (!(sums_counts[kd + i] > 0.f)) atomicAdd(n_empty, 1);
  }
}

__global__ void residual_update_kernel(float* __restrict__ R, int64_t n, int d4, const float* __restrict__ centroids,
                                       int K, const int32_t* __restrict__ assign, int64_t assign_stride) {
  const int64_t total = n * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const int c = (int)(i - row * d4);
    int a = assign[row * assign_stride];
    a = a < 0 ? 0 : (a >= K ? K - 1 : a);
    float4 v = reinterpret_cast<float4*>(R)[i];
    const float4 cc = __ldg(reinterpret_cast<const float4*>(centroids) + (int64_t)a * d4 + c);
    v.x -= cc.x; v.y -= cc.y; v.z -= cc.z; v.w -= cc.w;
    reinterpret_cast<float4*>(R)[i] = v;
  }
}

template <int KMAX>
cudaError_t launch_accumulate(const float* R, int64_t n, int d, const int32_t* assign, int64_t stride, int K, int G,
                              float* ps, int32_t* pc, cudaStream_t st) {
  // one thread per column pair; at least 256 threads because they also run the counting sort
  if (d <= 512) kmeans_accumulate_kernel<KMAX, 256><<<G, 256, 0, st>>>(R, n, d, assign, stride, K, ps, pc);
  else if (d <= 768) kmeans_accumulate_kernel<KMAX, 384><<<G, 384, 0, st>>>(R, n, d, assign, stride, K, ps, pc);
  else kmeans_accumulate_kernel<KMAX, 512><<<G, 512, 0, st>>>(R, n, d, assign, stride, K, ps, pc);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int mevi_kmeans_step(mevi_ctx* ctx, const float* R, int64_t n, int d, const float* centroids, int K, int mode,
                     int32_t* assign_out_or_null, int64_t assign_stride, float* sums_counts, double* inertia_or_null,
                     void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, R && centroids && sums_counts, "NULL argument");
  MEVI_REQUIRE(ctx, n >= 0 && d > 0 && d % 4 == 0 && d <= 1024 && K >= 1, "unsupported shape n=%lld d=%d K=%d",
               (long long)n, d, K);
  MEVI_REQUIRE(ctx, n < (int64_t)2147483647, "shard too large for int32 row ids");
  int32_t* assign = assign_out_or_null;
  int64_t stride = assign_stride;
  if (!assign) {
    assign = (int32_t*)mevi_ws(ctx, WS_KM_ASSIGN, (size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (!assign) return MEVI_ERR_NOMEM;
    stride = 1;
  }
  MEVI_REQUIRE(ctx, stride >= 1, "assign_stride must be >= 1");
  if (inertia_or_null) MEVI_CUDA(ctx, cudaMemsetAsync(inertia_or_null, 0, sizeof(double), st));
  const int64_t kd = (int64_t)K * d;
  if (n == 0) {
    MEVI_CUDA(ctx, cudaMemsetAsync(sums_counts, 0, (size_t)(kd + K) * sizeof(float), st));
    return MEVI_OK;
  }
  // 1. assignment
  bool use_tensor = false;
  if (mode == MEVI_MODE_TENSOR) {
    if (!mevi_rq_tensor_supported(ctx, d, 1, K, MEVI_METRIC_L2))
      return mevi_set_error(ctx, MEVI_ERR_UNSUPPORTED, "tensor k-means assignment unsupported for d=%d K=%d", d, K);
    use_tensor = true;
  } else if (mode == MEVI_MODE_AUTO) {
    use_tensor = mevi_rq_tensor_supported(ctx, d, 1, K, MEVI_METRIC_L2) && n >= 4096;
  }
  int rc;
  if (use_tensor)
    rc = mevi_rq_tensor_assign(ctx, R, n, d, centroids, 1, K, MEVI_METRIC_L2, assign, stride, nullptr, nullptr,
                               inertia_or_null, st);
  else
    rc = mevi_rq_exact_launch(ctx, R, n, d, centroids, 1, K, MEVI_METRIC_L2, assign, stride, nullptr, nullptr, nullptr,
                              nullptr, n, inertia_or_null, st);
  if (rc != MEVI_OK) return rc;

  // 2. accumulation
  const bool fast = K <= 64 && d <= 1024 && d % 2 == 0;
  if (fast) {
    int G = ctx->sm_count * 2;
    int64_t max_g = (n + KM_TILE - 1) / KM_TILE;
    if (G > max_g) G = (int)max_g;
    if (G < 1) G = 1;
    size_t ps_bytes = (size_t)G * kd * sizeof(float);
    size_t pc_bytes = (size_t)G * K * sizeof(int32_t);
    char* ws = (char*)mevi_ws(ctx, WS_KM_PARTIAL, ps_bytes + pc_bytes);
    if (!ws) return MEVI_ERR_NOMEM;
    float* ps = (float*)ws;
    int32_t* pc = (int32_t*)(ws + ps_bytes);
    cudaError_t e;
    if (K <= 16) e = launch_accumulate<16>(R, n, d, assign, stride, K, G, ps, pc, st);
    else if (K <= 32) e = launch_accumulate<32>(R, n, d, assign, stride, K, G, ps, pc, st);
    else e = launch_accumulate<64>(R, n, d, assign, stride, K, G, ps, pc, st);
    if (e != cudaSuccess) return mevi_set_error(ctx, MEVI_ERR_CUDA, "kmeans_accumulate launch: %s", cudaGetErrorString(e));
    int threads = 256;
    int blocks = (int)((kd + K + threads - 1) / threads);
    kmeans_reduce_partials_kernel<<<blocks, threads, 0, st>>>(ps, pc, G, K, d, sums_counts);
    MEVI_COUNT_LAUNCH(ctx, 2);
    MEVI_CUDA(ctx, cudaGetLastError());
  } else {
    MEVI_CUDA(ctx, cudaMemsetAsync(sums_counts, 0, (size_t)(kd + K) * sizeof(float), st));
    int grid = ctx->sm_count * 8;
    kmeans_accumulate_atomic_kernel<<<grid, 256, 0, st>>>(R, n, d, assign, stride, K, sums_counts, sums_counts + kd);
    MEVI_COUNT_LAUNCH(ctx, 1);
    MEVI_CUDA(ctx, cudaGetLastError());
  }
  return MEVI_OK;
}

int mevi_kmeans_update(mevi_ctx* ctx, const float* sums_counts, int K, int d, float* centroids, int32_t* n_empty_or_null,
                       void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, sums_counts && centroids && K >= 1 && d >= 1, "bad argument");
  if (n_empty_or_null) MEVI_CUDA(ctx, cudaMemsetAsync(n_empty_or_null, 0, sizeof(int32_t), st));
  const int64_t kd = (int64_t)K * d;
  int threads = 256;
  int blocks = (int)((kd + threads - 1) / threads);
  kmeans_update_kernel<<<blocks, threads, 0, st>>>(sums_counts, K, d, centroids, n_empty_or_null);
  MEVI_COUNT_LAUNCH(ctx, 1);
  MEVI_CUDA(ctx, cudaGetLastError());
  return MEVI_OK;
}

int mevi_residual_update(mevi_ctx* ctx, float* R, int64_t n, int d, const float* centroids, int K, const int32_t* assign,
                         int64_t assign_stride, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, R && centroids && assign && d > 0 && d % 4 == 0 && K >= 1 && assign_stride >= 1, "bad argument");
  if (n <= 0) return MEVI_OK;
  int grid = ctx->sm_count * 16;
  residual_update_kernel<<<grid, 256, 0, st>>>(R, n, d / 4, centroids, K, assign, assign_stride);
  MEVI_COUNT_LAUNCH(ctx, 1);
  MEVI_CUDA(ctx, cudaGetLastError());
  return MEVI_OK;
}

}  // extern "C"
