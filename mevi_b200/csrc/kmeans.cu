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
#include "ptx.cuh"

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
// every centroid: KMAX x 2 fp32 accumulators in registers for the CTA's whole (contiguous) row range.
// Rows arrive in sub-tiles of up to 32 rows — one contiguous byte range, fetched with a single bulk
// async copy (TMA engine) into a double-buffered shared-memory stage, so ~100-190 KB per SM are in
// flight without holding registers.  Warp 0 buckets the sub-tile's rows by centroid (match_any +
// popc = stable counting sort); then for each centroid (unrolled, so the accumulator is a fixed
// register) every thread walks the bucket's rows in ascending order and adds its two columns from
// shared memory.  All threads work on every row, so one dominant cluster (the reference-trained
// codebooks put >80 % of N(0,1) rows into two centroids) costs nothing extra.
template <int KMAX, int NT>
__global__ void __launch_bounds__(NT, 1) kmeans_accumulate_kernel(const float* __restrict__ R, int64_t n, int d,
                                                                  const int32_t* __restrict__ assign,
                                                                  int64_t assign_stride, int K, int sub_rows,
                                                                  float* __restrict__ partial_sums,     // [grid][K][d]
                                                                  int32_t* __restrict__ partial_counts)  // [grid][K]
{
  extern __shared__ __align__(128) unsigned char km_smem[];  // [2][sub_rows][d] fp32
  __shared__ __align__(8) uint64_t full_bar[2];
  __shared__ int s_start[KMAX + 1];
  __shared__ int s_order[32];
  __shared__ int s_count_total[KMAX];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool owner = 2 * tid < d;  // this thread owns columns 2*tid, 2*tid+1
  const int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = row_begin + rows_per_cta < n ? row_begin + rows_per_cta : n;
  const int64_t my_rows = row_end > row_begin ? row_end - row_begin : 0;
  const int64_t nsub = (my_rows + sub_rows - 1) / sub_rows;
  const size_t stage_bytes = (size_t)sub_rows * d * sizeof(float);

  float2 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int i = tid; i < KMAX; i += NT) s_count_total[i] = 0;
  if (tid == 0) {
    ptx::mbar_init(&full_bar[0], 1);
    ptx::mbar_init(&full_bar[1], 1);
    ptx::mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int64_t i) {  // thread 0: bulk copy of sub-tile i into stage i & 1
    const int64_t r0 = row_begin + i * sub_rows;
    const int rows = (int)((row_end - r0) < sub_rows ? (row_end - r0) : sub_rows);
    const uint32_t bytes = (uint32_t)rows * (uint32_t)d * 4u;
    ptx::mbar_arrive_expect_tx(&full_bar[i & 1], bytes);
    ptx::bulk_g2s(km_smem + (size_t)(i & 1) * stage_bytes, R + r0 * d, bytes, &full_bar[i & 1]);
  };
  if (tid == 0 && nsub > 0) issue(0);
  // warp 0 prefetches the assignments of the next sub-tile one iteration ahead
  int next_a = KMAX;
  if (warp == 0 && nsub > 0) {
    const int64_t r = row_begin + lane;
    if (lane < sub_rows && r < row_end) {
      int a = assign[r * assign_stride];
      next_a = a < 0 ? 0 : (a >= K ? K - 1 : a);
    }
  }

  for (int64_t i = 0; i < nsub; ++i) {
    const int64_t r0 = row_begin + i * sub_rows;
    const int rows = (int)((row_end - r0) < sub_rows ? (row_end - r0) : sub_rows);
    if (tid == 0 && i + 1 < nsub) issue(i + 1);  // stage (i+1)&1 was released by the barrier that ended iteration i-1
    if (warp == 0) {
      // stable counting sort of <= 32 rows by centroid
      const int a = next_a;  // KMAX for lanes past the sub-tile
      next_a = KMAX;
      if (i + 1 < nsub) {
        const int64_t r = r0 + sub_rows + lane;
        if (lane < sub_rows && r < row_end) {
          int an = assign[r * assign_stride];
          next_a = an < 0 ? 0 : (an >= K ? K - 1 : an);
        }
      }
      const unsigned same = __match_any_sync(MEVI_FULL_MASK, a);
      const int rank = __popc(same & ((1u << lane) - 1u));
      // bucket starts: lane k (and k+32) counts the rows assigned to centroid k, then an exclusive scan
      int start_lo = 0, start_hi = 0;
      {
        int c_lo = 0, c_hi = 0;
#pragma unroll
        for (int src = 0; src < 32; ++src) {
          const int av = __shfl_sync(MEVI_FULL_MASK, a, src);
          c_lo += (av == lane);
          c_hi += (av == lane + 32);
        }
        if (lane < KMAX) s_count_total[lane] += c_lo;
        if (KMAX > 32 && lane + 32 < KMAX) s_count_total[lane + 32] += c_hi;
        int inc = c_lo;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(MEVI_FULL_MASK, inc, o);
          if (lane >= o) inc += v;
        }
        start_lo = inc - c_lo;
        const int total_lo = __shfl_sync(MEVI_FULL_MASK, inc, 31);
        int inc2 = c_hi;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(MEVI_FULL_MASK, inc2, o);
          if (lane >= o) inc2 += v;
        }
        start_hi = total_lo + inc2 - c_hi;
        if (lane < KMAX) s_start[lane] = start_lo;
        if (KMAX > 32 && lane + 32 < KMAX) s_start[lane + 32] = start_hi;
        if (lane == 0) s_start[KMAX] = rows;
      }
      {  // every lane takes part in the broadcast (the source lane = centroid id need not share `a`)
        const int src = a & 31;
        const int from_lo = __shfl_sync(MEVI_FULL_MASK, start_lo, src);
        const int from_hi = __shfl_sync(MEVI_FULL_MASK, start_hi, src);
        if (a < KMAX) s_order[(a < 32 ? from_lo : from_hi) + rank] = lane;
      }
    }
    ptx::mbar_wait(&full_bar[i & 1], (uint32_t)((i >> 1) & 1));
    __syncthreads();
    if (owner) {
      const float* base = reinterpret_cast<const float*>(km_smem + (size_t)(i & 1) * stage_bytes) + 2 * tid;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int b = s_start[k], e = s_start[k + 1];
        for (int j = b; j < e; ++j) {  // ascending row order inside the bucket: reproducible sums
          const float2 v = *reinterpret_cast<const float2*>(base + (size_t)s_order[j] * d);
          acc[k].x += v.x;
          acc[k].y += v.y;
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

template <int KMAX, int NT>
cudaError_t launch_accumulate_nt(const float* R, int64_t n, int d, const int32_t* assign, int64_t stride, int K, int G,
                                 float* ps, int32_t* pc, cudaStream_t st) {
  int sub_rows = (int)(96 * 1024 / ((size_t)d * 4));
  if (sub_rows > 32) sub_rows = 32;
  if (sub_rows < 1) sub_rows = 1;
  const size_t smem = 2 * (size_t)sub_rows * d * 4 + 128;
  cudaError_t e = cudaFuncSetAttribute(kmeans_accumulate_kernel<KMAX, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kmeans_accumulate_kernel<KMAX, NT><<<G, NT, smem, st>>>(R, n, d, assign, stride, K, sub_rows, ps, pc);
  return cudaGetLastError();
}

template <int KMAX>
cudaError_t launch_accumulate(const float* R, int64_t n, int d, const int32_t* assign, int64_t stride, int K, int G,
                              float* ps, int32_t* pc, cudaStream_t st) {
  // one thread per column pair
  if (d <= 256) return launch_accumulate_nt<KMAX, 128>(R, n, d, assign, stride, K, G, ps, pc, st);
  if (d <= 512) return launch_accumulate_nt<KMAX, 256>(R, n, d, assign, stride, K, G, ps, pc, st);
  if (d <= 768) return launch_accumulate_nt<KMAX, 384>(R, n, d, assign, stride, K, G, ps, pc, st);
  return launch_accumulate_nt<KMAX, 512>(R, n, d, assign, stride, K, G, ps, pc, st);
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
  const bool fast = K <= 64 && d <= 1024 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0;
  if (fast) {
    int G = ctx->sm_count;  // one CTA per SM (shared-memory stages), persistent over its row range
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
