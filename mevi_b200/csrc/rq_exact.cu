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
// Layout: one warp owns RPW rows; lane l holds float4 chunks l, l+32, ... of
// each row in registers (NCH chunks → d <= 128*NCH).  Centroids are read through
// L1 (one level = K*d*4 B = 96 KB at the shipped shape, L1-resident) and shared
// by the RPW rows.  The 8 partial sums a lane accumulates per centroid group
// are combined with a 9-shuffle reduce-scatter instead of 8 full reductions.
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

template <int NCH, int RPW>
__global__ void __launch_bounds__(256) rq_encode_exact_kernel(RqExactParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t n_items = p.n_items;
  if (p.n_work_dev) {
    int64_t nw = *p.n_work_dev;
    if (nw < n_items) n_items = nw;
  }
  const int d = p.d;
  double inertia_acc = 0.0;

  for (int64_t base = warp_global * RPW; base < n_items; base += n_warps * RPW) {
    int64_t row[RPW];
    int lvl0[RPW];
    bool valid[RPW];
    float4 x[RPW][NCH];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      int64_t it = base + r;
      valid[r] = it < n_items;
      row[r] = 0;
      lvl0[r] = 0;
      if (valid[r]) {
        row[r] = p.work_rows ? (int64_t)p.work_rows[it] : it;
        lvl0[r] = p.work_levels ? p.work_levels[it] : 0;
      }
#pragma unroll
      for (int t = 0; t < NCH; ++t) {
        int c4 = (lane + 32 * t) * 4;
        x[r][t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid[r] && c4 < d) x[r][t] = ld_stream_f4(p.X + row[r] * d + c4);
      }
    }

    for (int m = 0; m < p.M; ++m) {
      const float* cbm = p.cb + (int64_t)m * p.K * d;
      float best[RPW];
      int besti[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        best[r] = INFINITY;
        besti[r] = 0x7fffffff;
      }
      for (int k0 = 0; k0 < p.K; k0 += 8) {
        float part[RPW][8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int k = k0 + kk;
#pragma unroll
          for (int r = 0; r < RPW; ++r) part[r][kk] = 0.f;
          if (k < p.K) {
            const float* ck = cbm + (int64_t)k * d;
#pragma unroll
            for (int t = 0; t < NCH; ++t) {
              int c4 = (lane + 32 * t) * 4;
              if (c4 < d) {
                float4 c = ldg_f4(ck + c4);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                  if (p.metric == MEVI_METRIC_L2) {
                    float dx = x[r][t].x - c.x, dy = x[r][t].y - c.y, dz = x[r][t].z - c.z, dw = x[r][t].w - c.w;
                    part[r][kk] = fmaf(dx, dx, part[r][kk]);
                    part[r][kk] = fmaf(dy, dy, part[r][kk]);
                    part[r][kk] = fmaf(dz, dz, part[r][kk]);
                    part[r][kk] = fmaf(dw, dw, part[r][kk]);
                  } else {
                    part[r][kk] = fmaf(x[r][t].x, c.x, part[r][kk]);
                    part[r][kk] = fmaf(x[r][t].y, c.y, part[r][kk]);
                    part[r][kk] = fmaf(x[r][t].z, c.z, part[r][kk]);
                    part[r][kk] = fmaf(x[r][t].w, c.w, part[r][kk]);
                  }
                }
              }
            }
          }
        }
        // reduce-scatter 8 values over the warp: afterwards every lane holds the full
        // sum for centroid k0 + (lane & 7)
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
#pragma unroll
          for (int off = 4; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              float send = up ? part[r][i] : part[r][i + off];
              float keep = up ? part[r][i + off] : part[r][i];
              part[r][i] = keep + __shfl_xor_sync(MEVI_FULL_MASK, send, off);
            }
          }
          float v = part[r][0];
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 8);
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 16);
          const int k = k0 + (lane & 7);
          if (p.metric != MEVI_METRIC_L2) v = -v;  // argmax of the inner product
          if (k < p.K && v < best[r]) {             // strict: the lowest index wins exact ties
            best[r] = v;
            besti[r] = k;
          }
        }
      }
      // argmin over the 8 distinct lanes (value, then index)
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
          float ov = __shfl_xor_sync(MEVI_FULL_MASK, best[r], off);
          int oi = __shfl_xor_sync(MEVI_FULL_MASK, besti[r], off);
          if (ov < best[r] || (ov == best[r] && oi < besti[r])) {
            best[r] = ov;
            besti[r] = oi;
          }
        }
        int idx = besti[r];
        if (valid[r]) {
          if (m < lvl0[r]) {
            idx = p.codes[row[r] * p.codes_stride + m];  // keep the earlier decision
          } else if (lane == 0) {
            p.codes[row[r] * p.codes_stride + m] = idx;
          }
        } else {
          idx = 0;
        }
        if (p.inertia && m == p.M - 1 && valid[r] && lane == 0)
          inertia_acc += (double)(p.metric == MEVI_METRIC_L2 ? best[r] : -best[r]);
        // residual -= centroid[idx]  (pq.py:304-305, after every level)
        const float* ck = cbm + (int64_t)idx * d;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          int c4 = (lane + 32 * t) * 4;
          if (c4 < d) {
            float4 c = ldg_f4(ck + c4);
            x[r][t].x -= c.x;
            x[r][t].y -= c.y;
            x[r][t].z -= c.z;
            x[r][t].w -= c.w;
          }
        }
      }
    }
    if (p.residual) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        if (!valid[r]) continue;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          int c4 = (lane + 32 * t) * 4;
          if (c4 < d) *reinterpret_cast<float4*>(p.residual + row[r] * d + c4) = x[r][t];
        }
      }
    }
  }
  if (p.inertia && lane == 0 && inertia_acc != 0.0) atomicAdd(p.inertia, inertia_acc);
}

template <int NCH, int RPW>
cudaError_t launch_exact(const RqExactParams& p, int sm_count, cudaStream_t st) {
  const int threads = 256;
  const int warps_per_block = threads / 32;
  int64_t want = (p.n_items + (int64_t)RPW * warps_per_block - 1) / ((int64_t)RPW * warps_per_block);
  int64_t cap = (int64_t)sm_count * 8;
  int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  rq_encode_exact_kernel<NCH, RPW><<<grid, threads, 0, st>>>(p);
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
  if (work_rows) {
    // sparse re-decision of flagged rows: one row per warp, light on registers -> many warps per SM
    if (d <= 128) e = launch_exact<1, 1>(p, ctx->sm_count * 2, st);
    else if (d <= 256) e = launch_exact<2, 1>(p, ctx->sm_count * 2, st);
    else if (d <= 512) e = launch_exact<4, 1>(p, ctx->sm_count * 2, st);
    else if (d <= 768) e = launch_exact<6, 1>(p, ctx->sm_count * 2, st);
    else e = launch_exact<8, 1>(p, ctx->sm_count * 2, st);
  } else if (d <= 128) e = launch_exact<1, 4>(p, ctx->sm_count, st);
  else if (d <= 256) e = launch_exact<2, 4>(p, ctx->sm_count, st);
  else if (d <= 512) e = launch_exact<4, 4>(p, ctx->sm_count, st);
  else if (d <= 768) e = launch_exact<6, 4>(p, ctx->sm_count, st);
  else e = launch_exact<8, 2>(p, ctx->sm_count, st);
  if (e != cudaSuccess) return mevi_set_error(ctx, MEVI_ERR_CUDA, "rq_encode_exact launch: %s", cudaGetErrorString(e));
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}
