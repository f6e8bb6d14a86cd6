// rq_exact.cu — fp32 direct-form residual-quantisation kernels on CUDA cores.
//
// These kernels restate MEVI/pq.py:281-305 literally: per level the distance of
// the row's current fp32 residual to every centroid is -(sum (x-c)^2) in fp32
// (pq.py:124-131), the nearest centroid (lowest index on exact ties, as
// torch.max at pq.py:302) is recorded and subtracted elementwise in fp32
// (pq.py:304-305).  They are (a) the whole encode in MEVI_MODE_EXACT, (b) the
// arbiter for rows the tensor-core prefilter flags as too close to call, and
// (c) the assignment step of the k-means trainer in exact mode.
//
// Layout: one CTA owns a group of 32 rows whose fp32 residuals stay in shared
// memory across all levels; lane l of a warp covers float4 chunks l, l+32, ...
// (NCH chunks -> d <= 128*NCH).  Each warp streams centroids through registers,
// two at a time, and every centroid is reused by all 32 rows, so the codebook is
// read once per group.  Eight rows' partial sums are combined with a 9-shuffle
// reduce-scatter instead of 8 full reductions.  The summation order per distance
// (4 sequential FMAs per chunk, chunks in order, then the lane tree) is fixed, so
// results are deterministic and independent of the grouping.
// Roofline: FP32 issue-bound (2*d*K*M lane-ops per row = 196,608 at 768/32/4).
#include "common.cuh"

namespace {

struct RqExactParams {
  const float* X;
  int64_t n;          // rows in X
  int d;
  const float* cb;    // [M][K][d]
  int M, K, metric;
  int32_t* codes;     // [*, codes_stride]
  int64_t codes_stride;
  float* residual;    // [n, d] or null
  const int32_t* work_rows;    // optional worklist of row ids
  const int32_t* work_levels;  // optional: level from which the row is re-decided (codes below it are kept)
  const int64_t* n_work_dev;   // optional: device count of worklist entries (read at kernel start)
  int64_t n_items;    // rows to process (upper bound when n_work_dev is set)
  double* inertia;    // optional: += distance to the chosen centroid of the last level
};

constexpr int GR = 32;        // rows per CTA group
constexpr int KB = 32;        // centroids per distance block
constexpr int EX_THREADS = 256;

// One CTA owns a group of GR rows: the fp32 residuals live in shared memory for all M levels, the
// level's centroids stream through registers (two at a time per warp, each reused by all GR rows), so
// the codebook is read once per group instead of once per row.
template <int NCH>
__global__ void __launch_bounds__(EX_THREADS, 2) rq_exact_group_kernel(RqExactParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int DP = NCH * 128;              // padded row width in floats
  float* sX = sm;                            // [GR][DP]
  float* sDist = sX + GR * DP;               // [GR][KB+1]
  float* sBest = sDist + GR * (KB + 1);      // [GR]
  int* sBestI = reinterpret_cast<int*>(sBest + GR);  // [GR]
  int* sLvl0 = sBestI + GR;                  // [GR]
  int64_t* sRow = reinterpret_cast<int64_t*>(sLvl0 + GR);  // [GR]  (-1 = padding)
  __shared__ int s_min_lvl0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = p.d;
  int64_t n_items = p.n_items;
  if (p.n_work_dev) {
    const int64_t nw = *p.n_work_dev;
    if (nw < n_items) n_items = nw;
  }
  double inertia_acc = 0.0;

  for (int64_t g0 = (int64_t)blockIdx.x * GR; g0 < n_items; g0 += (int64_t)gridDim.x * GR) {
    __syncthreads();
    if (tid == 0) s_min_lvl0 = 0x7fffffff;
    __syncthreads();
    if (tid < GR) {
      const int64_t it = g0 + tid;
      int64_t row = -1;
      int l0 = 0x7fffffff;
      if (it < n_items) {
        row = p.work_rows ? (int64_t)p.work_rows[it] : it;
        l0 = p.work_levels ? p.work_levels[it] : 0;
        atomicMin(&s_min_lvl0, l0);
      }
      sRow[tid] = row;
      sLvl0[tid] = l0;
    }
    __syncthreads();
    // rows -> shared memory (warp per row, 512 contiguous bytes per instruction)
    for (int r = warp; r < GR; r += EX_THREADS / 32) {
      const int64_t row = sRow[r];
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        const int c4 = (lane + 32 * t) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= 0 && c4 < d) v = ld_stream_f4(p.X + row * d + c4);
        *reinterpret_cast<float4*>(sX + r * DP + c4) = v;
      }
    }
    __syncthreads();
    const int min_lvl0 = s_min_lvl0;

    for (int m = 0; m < p.M; ++m) {
      const float* cbm = p.cb + (int64_t)m * p.K * d;
      if (tid < GR) {
        sBest[tid] = INFINITY;
        sBestI[tid] = 0;
      }
      if (m >= min_lvl0) {
        for (int k0 = 0; k0 < p.K; k0 += KB) {
          // distances of all GR rows to centroids k0 .. k0+31: warp w takes k0+w, +8, +16, +24, two at a time
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            const int ka = k0 + warp + 16 * pass, kb = ka + 8;
            float4 ca[NCH], cb2[NCH];
#pragma unroll
            for (int t = 0; t < NCH; ++t) {
              const int c4 = (lane + 32 * t) * 4;
              ca[t] = make_float4(0.f, 0.f, 0.f, 0.f);
              cb2[t] = ca[t];
              if (c4 < d) {
                if (ka < p.K) ca[t] = ldg_f4(cbm + (int64_t)ka * d + c4);
                if (kb < p.K) cb2[t] = ldg_f4(cbm + (int64_t)kb * d + c4);
              }
            }
            for (int r0 = 0; r0 < GR; r0 += 8) {
              float pa[8], pb[8];
#pragma unroll
              for (int rr = 0; rr < 8; ++rr) {
                pa[rr] = 0.f;
                pb[rr] = 0.f;
#pragma unroll
                for (int t = 0; t < NCH; ++t) {
                  const float4 x = *reinterpret_cast<const float4*>(sX + (r0 + rr) * DP + (lane + 32 * t) * 4);
                  if (p.metric == MEVI_METRIC_L2) {
                    float e0 = x.x - ca[t].x, e1 = x.y - ca[t].y, e2 = x.z - ca[t].z, e3 = x.w - ca[t].w;
                    pa[rr] = fmaf(e0, e0, pa[rr]); pa[rr] = fmaf(e1, e1, pa[rr]);
                    pa[rr] = fmaf(e2, e2, pa[rr]); pa[rr] = fmaf(e3, e3, pa[rr]);
                    e0 = x.x - cb2[t].x; e1 = x.y - cb2[t].y; e2 = x.z - cb2[t].z; e3 = x.w - cb2[t].w;
                    pb[rr] = fmaf(e0, e0, pb[rr]); pb[rr] = fmaf(e1, e1, pb[rr]);
                    pb[rr] = fmaf(e2, e2, pb[rr]); pb[rr] = fmaf(e3, e3, pb[rr]);
                  } else {
                    pa[rr] = fmaf(x.x, ca[t].x, pa[rr]); pa[rr] = fmaf(x.y, ca[t].y, pa[rr]);
                    pa[rr] = fmaf(x.z, ca[t].z, pa[rr]); pa[rr] = fmaf(x.w, ca[t].w, pa[rr]);
                    pb[rr] = fmaf(x.x, cb2[t].x, pb[rr]); pb[rr] = fmaf(x.y, cb2[t].y, pb[rr]);
                    pb[rr] = fmaf(x.z, cb2[t].z, pb[rr]); pb[rr] = fmaf(x.w, cb2[t].w, pb[rr]);
                  }
                }
              }
              // reduce-scatter over the warp: lane l ends with the total of row r0 + (l & 7)
#pragma unroll
              for (int off = 4; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                  float send = up ? pa[i] : pa[i + off];
                  float keep = up ? pa[i + off] : pa[i];
                  pa[i] = keep + __shfl_xor_sync(MEVI_FULL_MASK, send, off);
                  send = up ? pb[i] : pb[i + off];
                  keep = up ? pb[i + off] : pb[i];
                  pb[i] = keep + __shfl_xor_sync(MEVI_FULL_MASK, send, off);
                }
              }
              float va = pa[0], vb = pb[0];
              va += __shfl_xor_sync(MEVI_FULL_MASK, va, 8);
              va += __shfl_xor_sync(MEVI_FULL_MASK, va, 16);
              vb += __shfl_xor_sync(MEVI_FULL_MASK, vb, 8);
              vb += __shfl_xor_sync(MEVI_FULL_MASK, vb, 16);
              if (lane < 8) {
                sDist[(r0 + lane) * (KB + 1) + (ka - k0)] = va;
                sDist[(r0 + lane) * (KB + 1) + (kb - k0)] = vb;
              }
            }
          }
          __syncthreads();
          if (tid < GR) {
            float best = sBest[tid];
            int besti = sBestI[tid];
            for (int kk = 0; kk < KB && k0 + kk < p.K; ++kk) {
              float v = sDist[tid * (KB + 1) + kk];
              if (p.metric != MEVI_METRIC_L2) v = -v;  // argmax of the inner product
              if (v < best) {                          // strict: the lowest index wins exact ties
                best = v;
                besti = k0 + kk;
              }
            }
            sBest[tid] = best;
            sBestI[tid] = besti;
          }
          __syncthreads();
        }
      }
      __syncthreads();
      if (tid < GR) {
        const int64_t row = sRow[tid];
        if (row >= 0) {
          int idx = sBestI[tid];
          if (m < sLvl0[tid]) {
            idx = p.codes[row * p.codes_stride + m];  // keep the earlier decision
          } else {
            p.codes[row * p.codes_stride + m] = idx;
            if (p.inertia && m == p.M - 1) inertia_acc += (double)(p.metric == MEVI_METRIC_L2 ? sBest[tid] : -sBest[tid]);
          }
          sBestI[tid] = idx;
        } else {
          sBestI[tid] = 0;
        }
      }
      __syncthreads();
      // residual -= centroid[idx]  (pq.py:304-305, after every level)
      for (int r = warp; r < GR; r += EX_THREADS / 32) {
        const float* ck = cbm + (int64_t)sBestI[r] * d;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          const int c4 = (lane + 32 * t) * 4;
          if (c4 < d) {
            float4 x = *reinterpret_cast<float4*>(sX + r * DP + c4);
            const float4 c = ldg_f4(ck + c4);
            x.x -= c.x; x.y -= c.y; x.z -= c.z; x.w -= c.w;
            *reinterpret_cast<float4*>(sX + r * DP + c4) = x;
          }
        }
      }
      __syncthreads();
    }
    if (p.residual) {
      for (int r = warp; r < GR; r += EX_THREADS / 32) {
        const int64_t row = sRow[r];
        if (row < 0) continue;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          const int c4 = (lane + 32 * t) * 4;
          if (c4 < d) *reinterpret_cast<float4*>(p.residual + row * d + c4) = *reinterpret_cast<const float4*>(sX + r * DP + c4);
        }
      }
    }
  }
  if (p.inertia && tid < GR && inertia_acc != 0.0) atomicAdd(p.inertia, inertia_acc);
}

template <int NCH>
cudaError_t launch_group(const RqExactParams& p, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)GR * NCH * 128 * 4 + (size_t)GR * (KB + 1) * 4 + GR * 4 * 2 + GR * 4 + GR * 8 + 64;
  cudaError_t e = cudaFuncSetAttribute(rq_exact_group_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t want = (p.n_items + GR - 1) / GR;
  int64_t cap = (int64_t)sm_count * 2;
  int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  rq_exact_group_kernel<NCH><<<grid, EX_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

// internal entry used by the encode API, the tensor fix-up and the k-means step
int mevi_rq_exact_launch(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                         int32_t* codes, int64_t codes_stride, float* residual, const int32_t* work_rows,
                         const int32_t* work_levels, const int64_t* n_work_dev, int64_t n_items, double* inertia,
                         cudaStream_t st) {
  MEVI_REQUIRE(ctx, d > 0 && d % 4 == 0 && d <= 1024, "exact RQ kernel needs d %% 4 == 0 and d <= 1024 (got %d)", d);
  MEVI_REQUIRE(ctx, M >= 1 && K >= 1, "bad codebook shape M=%d K=%d", M, K);
  MEVI_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0,
               "X and codebook must be 16-byte aligned");
  if (n_items <= 0) return MEVI_OK;
  RqExactParams p;
  p.X = X; p.n = n; p.d = d; p.cb = cb; p.M = M; p.K = K; p.metric = metric;
  p.codes = codes; p.codes_stride = codes_stride; p.residual = residual;
  p.work_rows = work_rows; p.work_levels = work_levels; p.n_work_dev = n_work_dev; p.n_items = n_items;
  p.inertia = inertia;
  cudaError_t e;
  if (d <= 128) e = launch_group<1>(p, ctx->sm_count, st);
  else if (d <= 256) e = launch_group<2>(p, ctx->sm_count, st);
  else if (d <= 512) e = launch_group<4>(p, ctx->sm_count, st);
  else if (d <= 768) e = launch_group<6>(p, ctx->sm_count, st);
  else e = launch_group<8>(p, ctx->sm_count, st);
  if (e != cudaSuccess) return mevi_set_error(ctx, MEVI_ERR_CUDA, "rq_encode_exact launch: %s", cudaGetErrorString(e));
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}
