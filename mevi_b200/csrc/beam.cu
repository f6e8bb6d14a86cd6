// beam.cu — RQ beam search on the device: the leaf producer of the re-rank path.
// Replaces MEVI/pq.py:613-713 (rq branch): per level softmax(compute_scores(residual, codebook[i])) (661-662),
// times the running beam score when rq_topk_score == 'prod' (663-665), top-`num_beams` over beam x K
// (690; every candidate kept while there are at most num_beams of them, 701-707), gather of the winning
// beams' code prefixes (696-697) and residuals (698-700).
//
// The reference materialises a [bs, beams, K, d] temporary per level (31 MB per query row at beams=100).
// Here nothing of width d is ever formed after the first step: with r = x - sum_{m<i} c^m_{k_m},
//     r.c  = x.c - sum_{m<i} G[(m,k_m), c]            G = Gram matrix of all M*K centroids
//     |r - c|^2 = |r|^2 - 2 r.c + |c|^2               and |r_{i+1}|^2 is the distance chosen at level i
// so one pass computes the M*K inner products x.c and every level is table arithmetic over beam x K
// candidates.  Tables and the recurrences are kept in fp64, so each logit is the correctly rounded fp32
// value (the reference's own fp32 sum over d carries ~1e-4 absolute error at |x|^2 ~ 1e3).
// One CTA per query row; top-k in (score desc, candidate index asc) order: a radix select finds the num_beams-th best
// probability, the candidates at or above it (num_beams plus ties) are compacted and only those are sorted; the full
// shared-memory bitonic sort of all beam x K candidates remains for the degenerate case of more than BEAM_SEL ties.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int BEAM_THREADS = 256;
constexpr int BEAM_SEL_MAX = 1024;  // selection buffer: power of two >= 2 * num_beams, at most this

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MEVI_FULL_MASK, v, o);
  return v;
}

// G[a][b] = c_a . c_b in fp64, one warp per pair
__global__ void __launch_bounds__(256) beam_gram_kernel(const float* __restrict__ cb, int NT, int d, double* __restrict__ G) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (int64_t)NT * NT) return;
  const int a = (int)(warp / NT), b = (int)(warp % NT);
  if (b < a) return;  // symmetric: the (a, b >= a) warp writes both entries
  const float* ca = cb + (int64_t)a * d;
  const float* cbb = cb + (int64_t)b * d;
  double acc = 0.0;
  for (int c = lane; c < d; c += 32) acc = fma((double)ca[c], (double)cbb[c], acc);
  acc = warp_sum_d(acc);
  if (lane == 0) {
    G[(int64_t)a * NT + b] = acc;
    G[(int64_t)b * NT + a] = acc;
  }
}

struct BeamParams {
  const float* X;
  const float* cb;
  const double* G;
  int64_t bs;
  int d, M, K, NT, num_beams, cap, sel, metric, prod;
  int32_t* labels;
  float* scores;
};

__global__ void __launch_bounds__(BEAM_THREADS) beam_search_kernel(BeamParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = p.d, M = p.M, K = p.K, NT = p.NT, B = p.num_beams;
  double* sxc = reinterpret_cast<double*>(smem);                 // [NT]
  double* sr2 = sxc + NT;                                        // [2][B]
  float* sx = reinterpret_cast<float*>(sr2 + 2 * B);             // [d]
  float* sscore = sx + d;                                        // [2][B]
  float* cscore = sscore + 2 * B;                                // [cap]
  int32_t* cidx = reinterpret_cast<int32_t*>(cscore + p.cap);    // [cap]
  int32_t* scodes = cidx + p.cap;                                // [2][B*M]
  float* sel_score = reinterpret_cast<float*>(scodes + 2 * B * M);  // [p.sel]
  int32_t* sel_idx = reinterpret_cast<int32_t*>(sel_score + p.sel); // [p.sel]
  __shared__ double s_xn2;
  __shared__ unsigned s_hist[256], s_scratch[4];
  __shared__ int s_nsel;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BEAM_THREADS / 32;
  const bool l2 = p.metric == MEVI_METRIC_L2;

  for (int64_t q = blockIdx.x; q < p.bs; q += gridDim.x) {
    const float* x = p.X + q * d;
    for (int c = tid; c < d; c += BEAM_THREADS) sx[c] = x[c];
    __syncthreads();
    // x . c for every centroid of every level, |x|^2
    for (int t = warp; t <= NT; t += nwarps) {
      double acc = 0.0;
      if (t < NT) {
        const float* c = p.cb + (int64_t)t * d;
        for (int e = lane; e < d; e += 32) acc = fma((double)sx[e], (double)__ldg(c + e), acc);
      } else {
        for (int e = lane; e < d; e += 32) acc = fma((double)sx[e], (double)sx[e], acc);
      }
      acc = warp_sum_d(acc);
      if (lane == 0) {
        if (t < NT) sxc[t] = acc;
        else s_xn2 = acc;
      }
    }
    __syncthreads();
    int cur = 0, prev = 1;
    if (tid == 0) {
      sscore[0] = 1.f;
      sr2[0] = s_xn2;
    }
    __syncthreads();
    for (int i = 0; i < M; ++i) {
      const int ncand = prev * K;
      const int32_t* codes_cur = scodes + cur * B * M;
      // logits of every (beam, centroid) candidate
      for (int t = tid; t < ncand; t += BEAM_THREADS) {
        const int b = t / K, k = t - b * K;
        const int col = i * K + k;
        double rc = sxc[col];
        for (int m = 0; m < i; ++m) rc -= p.G[(int64_t)(m * K + codes_cur[b * M + m]) * NT + col];
        const double val = l2 ? -(sr2[cur * B + b] - 2.0 * rc + p.G[(int64_t)col * NT + col]) : rc;
        cscore[t] = (float)val;
      }
      __syncthreads();
      // softmax over the K centroids of each beam (pq.py:662), times the beam score ('prod', 664-665)
      for (int b = warp; b < prev; b += nwarps) {
        float* row = cscore + b * K;
        float mx = -CUDART_INF_F;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(MEVI_FULL_MASK, mx, o));
        float sum = 0.f;
        for (int k = lane; k < K; k += 32) {
          const float e = expf(row[k] - mx);
          row[k] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        const float bsc = sscore[cur * B + b];
        for (int k = lane; k < K; k += 32) {
          float pr = row[k] / sum;
          if (p.prod) pr = bsc * pr;
          row[k] = pr;
          cidx[b * K + k] = b * K + k;
        }
      }
      __syncthreads();
      int nb;
      if (B < ncand) {
        // the B-th best probability, then everything at or above it (B candidates plus ties at the boundary)
        const float kth = block256_select_kth(cscore, ncand, B, s_hist, s_scratch);
        if (tid == 0) s_nsel = 0;
        __syncthreads();
        for (int t = tid; t < ncand; t += BEAM_THREADS) {
          const float v = cscore[t];
          if (!(v < kth)) {
            const int pos = atomicAdd(&s_nsel, 1);
            if (pos < p.sel) { sel_score[pos] = v; sel_idx[pos] = cidx[t]; }
          }
        }
        __syncthreads();
        const int nsel = s_nsel;
        if (nsel <= p.sel) {
          int n2 = 32;
          while (n2 < nsel) n2 <<= 1;
          for (int t = nsel + tid; t < n2; t += BEAM_THREADS) {
            sel_score[t] = -CUDART_INF_F;
            sel_idx[t] = 0x7fffffff;
          }
          __syncthreads();
          block_bitonic_sort<int32_t>(sel_score, sel_idx, n2);
          for (int t = tid; t < B; t += BEAM_THREADS) {
            cscore[t] = sel_score[t];
            cidx[t] = sel_idx[t];
          }
          __syncthreads();
        } else {  // more ties at the boundary than the buffer holds: sort everything
          int n2 = 1;
          while (n2 < ncand) n2 <<= 1;
          for (int t = ncand + tid; t < n2; t += BEAM_THREADS) {
            cscore[t] = -CUDART_INF_F;
            cidx[t] = 0x7fffffff;
          }
          __syncthreads();
          block_bitonic_sort<int32_t>(cscore, cidx, n2);
        }
        nb = B;
      } else {
        nb = ncand;  // every candidate survives, in (beam, centroid) order (pq.py:701-707)
      }
      const int nxt = cur ^ 1;
      int32_t* codes_nxt = scodes + nxt * B * M;
      for (int t = tid; t < nb; t += BEAM_THREADS) {
        const int idx = cidx[t];
        const int pb = idx / K, k = idx - pb * K;
        sscore[nxt * B + t] = cscore[t];
        for (int m = 0; m < i; ++m) codes_nxt[t * M + m] = codes_cur[pb * M + m];
        codes_nxt[t * M + i] = k;
        if (l2 && i != M - 1) {
          const int col = i * K + k;
          double rc = sxc[col];
          for (int m = 0; m < i; ++m) rc -= p.G[(int64_t)(m * K + codes_cur[pb * M + m]) * NT + col];
          sr2[nxt * B + t] = sr2[cur * B + pb] - 2.0 * rc + p.G[(int64_t)col * NT + col];
        }
      }
      __syncthreads();
      cur = nxt;
      prev = nb;
    }
    // prev == num_beams here (checked on the host: K^M >= num_beams)
    for (int t = tid; t < prev * M; t += BEAM_THREADS) p.labels[q * B * M + t] = scodes[cur * B * M + t];
    for (int t = tid; t < prev; t += BEAM_THREADS) p.scores[q * B + t] = sscore[cur * B + t];
    __syncthreads();
  }
}

}  // namespace

extern "C" int mevi_rq_beam_search(mevi_ctx* ctx, const float* X, int64_t bs, int d, const float* codebook, int M, int K,
                                   int metric, int num_beams, int prod, int32_t* labels, float* scores, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, X && codebook && labels && scores, "NULL argument");
  MEVI_REQUIRE(ctx, d > 0 && M > 0 && K > 0 && num_beams > 0, "bad shape");
  MEVI_REQUIRE(ctx, metric == MEVI_METRIC_L2 || metric == MEVI_METRIC_IP, "bad metric %d", metric);
  const int NT = M * K;
  MEVI_REQUIRE(ctx, NT <= 2048, "beam search supports M*K <= 2048 (got %d)", NT);
  {
    // pq.py:708 asserts beam_scores.size(1) == num_beams: there must be at least num_beams leaves
    double leaves = 1.0;
    for (int j = 0; j < M; ++j) leaves *= (double)K;
    MEVI_REQUIRE(ctx, leaves >= (double)num_beams, "num_beams=%d exceeds the K^M=%.0f leaves of the tree", num_beams, leaves);
  }
  if (bs <= 0) return MEVI_OK;
  // live beams never exceed num_beams, so a level never sees more than num_beams*K candidates
  int cap = 1;
  while (cap < num_beams * K) cap <<= 1;
  MEVI_REQUIRE(ctx, cap <= 16384, "beam search supports num_beams*K <= 16384 (got %d)", num_beams * K);
  double* G = (double*)mevi_ws(ctx, WS_MISC, sizeof(double) * (size_t)NT * NT);
  if (!G) return MEVI_ERR_NOMEM;
  {
    const int64_t warps = (int64_t)NT * NT;
    beam_gram_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(codebook, NT, d, G);
    MEVI_CUDA(ctx, cudaGetLastError());
  }
  BeamParams p;
  p.X = X; p.cb = codebook; p.G = G; p.bs = bs; p.d = d; p.M = M; p.K = K; p.NT = NT; p.num_beams = num_beams;
  p.cap = cap; p.metric = metric; p.prod = prod ? 1 : 0; p.labels = labels; p.scores = scores;
  int sel = 64;
  while (sel < 2 * num_beams && sel < BEAM_SEL_MAX) sel <<= 1;
  p.sel = sel;
  const size_t smem = sizeof(double) * (size_t)(NT + 2 * num_beams) + sizeof(float) * (size_t)(d + 2 * num_beams + cap) +
                      sizeof(int32_t) * (size_t)(cap + 2 * num_beams * M) + 8 * (size_t)sel;
  MEVI_REQUIRE(ctx, smem <= 200 * 1024, "beam search state (%zu bytes) does not fit shared memory", smem);
  MEVI_CUDA(ctx, cudaFuncSetAttribute(beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t max_grid = (int64_t)ctx->sm_count * 8;
  const int grid = (int)(bs < max_grid ? bs : max_grid);
  beam_search_kernel<<<grid, BEAM_THREADS, smem, st>>>(p);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 2);
  return MEVI_OK;
}
