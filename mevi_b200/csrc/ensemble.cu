// ensemble.cu — result fusion as a device epilogue of the re-rank / flat search.
// Replaces the per-query python dictionaries of MEVI/ensemble_marco.py:181-191 (rank of a document's RQ leaf in the
// query's beam-search leaves), 221-240 (score' = score + alpha / (beta * crank + 1), punished by (1 - gamma * alpha)
// when the leaf is not among them; later occurrences of a document overwrite earlier ones; ranking by descending
// fused score, ties in first-insertion order) and ensemble_nqdpr.py:232-251, plus the list look-ups of the two
// evaluators (ensemble_marco.py:20-31 `preds.index(g)`, ensemble_nqdpr.py:23-33 first hit).
// Arithmetic is float64 with explicitly rounded operations (no FMA contraction): the fused scores are bit-identical
// to the reference's python floats.  One CTA per query; lists of up to 4,096 candidates are sorted in shared memory.
#include <math_constants.h>

#include "common.cuh"

namespace {
constexpr int ENS_THREADS = 256;
constexpr int ENS_MAX_P = 4096;

// ---- cluster ranks ----------------------------------------------------------------------------------------
// python: cr = {}; for i, clus in enumerate(leaves[q]): cr[tuple(clus)] = i   (a repeated leaf keeps its LAST index,
// len(cr) counts distinct leaves); rank(p) = cr.get(mapping[p] if p != -1 else -1, len(cr)).
__global__ void __launch_bounds__(ENS_THREADS)
ensemble_cluster_ranks_kernel(const int64_t* __restrict__ cand_ids, const int32_t* __restrict__ cand_count, int P,
                              const int32_t* __restrict__ codes, int64_t n_docs, int M,
                              const int32_t* __restrict__ query_leaves, int L, int32_t* __restrict__ cranks,
                              int32_t* __restrict__ num_leaves) {
  extern __shared__ int32_t s_leaf[];  // [L][M], then [L] "is the last occurrence" flags
  int32_t* s_last = s_leaf + (size_t)L * M;
  __shared__ int s_distinct;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) s_distinct = 0;
  for (int i = threadIdx.x; i < L * M; i += blockDim.x) s_leaf[i] = query_leaves[(int64_t)q * L * M + i];
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    bool last = true;
    for (int j = i + 1; j < L && last; ++j) {
      bool same = true;
      for (int m = 0; m < M; ++m) same &= (s_leaf[j * M + m] == s_leaf[i * M + m]);
      if (same) last = false;
    }
    s_last[i] = last ? 1 : 0;
    if (last) atomicAdd(&s_distinct, 1);
  }
  __syncthreads();
  const int distinct = s_distinct;
  if (threadIdx.x == 0) num_leaves[q] = distinct;
  const int cnt = cand_count ? min(cand_count[q], P) : P;
  for (int c = threadIdx.x; c < cnt; c += blockDim.x) {
    const int64_t p = cand_ids[(int64_t)q * P + c];
    int rank = distinct;
    if (p == -1) {
      rank = distinct;  // cr.get(-1, len(cr)): an int is never a leaf tuple
    } else if (p < 0 || p >= n_docs) {
      rank = -2;  // mapping[p] would raise KeyError: reported to the host
    } else {
      const int32_t* code = codes + p * M;
      if (code[0] == INT32_MIN) rank = -2;  // a hole of the dense mapping: not a key of the dictionary
      for (int i = 0; i < L && rank != -2; ++i) {
        if (!s_last[i]) continue;
        bool same = true;
        for (int m = 0; m < M; ++m) same &= (s_leaf[i * M + m] == code[m]);
        if (same) { rank = i; break; }
      }
    }
    cranks[(int64_t)q * P + c] = rank;
  }
}

// ---- fusion + ranking -------------------------------------------------------------------------------------
template <typename KeyLess>
__device__ __forceinline__ void bitonic_pass(int n, int k, int j, KeyLess&& swap_if_needed) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int ixj = i ^ j;
    if (ixj > i) swap_if_needed(i, ixj, (i & k) == 0);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ENS_THREADS)
ensemble_fuse_kernel(const int64_t* __restrict__ cand_ids, const double* __restrict__ cand_scores,
                     const int32_t* __restrict__ cranks, const int32_t* __restrict__ cand_count, int P, int P2,
                     double alpha, double beta, double gamma, int num_leaves, int64_t* __restrict__ out_ids,
                     double* __restrict__ out_scores, int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char ens_smem[];
  int64_t* s_id = reinterpret_cast<int64_t*>(ens_smem);        // [P2] sort 1 key (document)
  double* s_v = reinterpret_cast<double*>(s_id + P2);          // [P2] fused value by ORIGINAL position
  double* s_kv = s_v + P2;                                      // [P2] sort 2 key (value)
  int32_t* s_pos = reinterpret_cast<int32_t*>(s_kv + P2);      // [P2] sort 1 payload (position)
  int32_t* s_kpos = s_pos + P2;                                 // [P2] sort 2 key (first position)
  __shared__ int s_unique;
  const int q = blockIdx.x;
  const int cnt = cand_count ? min(cand_count[q], P) : P;
  if (threadIdx.x == 0) s_unique = 0;
  // v = s + alpha / (beta * crank + 1); v *= (1 - gamma * alpha) when crank == num_leaves — every operation rounded
  // on its own, as the python expression is
  const double punish = __dsub_rn(1.0, __dmul_rn(gamma, alpha));
  for (int i = threadIdx.x; i < P2; i += blockDim.x) {
    if (i < cnt) {
      const int crank = cranks[(int64_t)q * P + i];
      double v = __dadd_rn(cand_scores[(int64_t)q * P + i],
                           __ddiv_rn(alpha, __dadd_rn(__dmul_rn(beta, (double)crank), 1.0)));
      if (crank == num_leaves) v = __dmul_rn(v, punish);
      s_v[i] = v;
      s_id[i] = cand_ids[(int64_t)q * P + i];
      s_pos[i] = i;
    } else {
      s_v[i] = 0.0;
      s_id[i] = INT64_MAX;  // padding sorts last; real ids are document numbers or -1
      s_pos[i] = INT32_MAX;
    }
  }
  __syncthreads();
  // sort 1: (document, position) ascending
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1)
      bitonic_pass(P2, k, j, [&](int a, int b, bool up) {
        int64_t ia = s_id[a], ib = s_id[b];
        int32_t pa = s_pos[a], pb = s_pos[b];
        bool a_first = (ia < ib) || (ia == ib && pa < pb);
        if (a_first != up) { s_id[a] = ib; s_id[b] = ia; s_pos[a] = pb; s_pos[b] = pa; }
      });
  // a run of equal documents = one dictionary entry: inserted at its first position, holding its last value
  int mine = 0;
  for (int i = threadIdx.x; i < P2; i += blockDim.x) {
    double kv = -CUDART_INF;
    int32_t kp = INT32_MAX;
    if (i < cnt) {
      const int64_t id = s_id[i];
      if (i == 0 || s_id[i - 1] != id) {
        int lo = i, hi = cnt;  // last index of the run: upper bound of id in [i, cnt)
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (s_id[mid] == id) lo = mid; else hi = mid;
        }
        kv = s_v[s_pos[lo]];
        kp = s_pos[i];
        ++mine;
      }
    }
    s_kv[i] = kv;
    s_kpos[i] = kp;
  }
  if (mine) atomicAdd(&s_unique, mine);
  __syncthreads();
  // sort 2: value descending, first position ascending (python's stable sort on -value)
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1)
      bitonic_pass(P2, k, j, [&](int a, int b, bool up) {
        double va = s_kv[a], vb = s_kv[b];
        int32_t pa = s_kpos[a], pb = s_kpos[b];
        bool a_first = (va > vb) || (va == vb && pa < pb);
        if (a_first != up) { s_kv[a] = vb; s_kv[b] = va; s_kpos[a] = pb; s_kpos[b] = pa; }
      });
  const int uniq = s_unique;
  if (threadIdx.x == 0) out_count[q] = uniq;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < uniq) {
      out_ids[(int64_t)q * P + i] = cand_ids[(int64_t)q * P + s_kpos[i]];
      out_scores[(int64_t)q * P + i] = s_kv[i];
    } else {
      out_ids[(int64_t)q * P + i] = -1;
      out_scores[(int64_t)q * P + i] = -CUDART_INF;
    }
  }
}

// ---- evaluator look-ups -----------------------------------------------------------------------------------
// positions[q, g] = ranked[q].index(targets[q, g]) or -1 (ensemble_marco.py:20-22)
__global__ void __launch_bounds__(ENS_THREADS)
ensemble_positions_kernel(const int64_t* __restrict__ ranked, const int32_t* __restrict__ ranked_count, int P,
                          const int64_t* __restrict__ targets, const int32_t* __restrict__ target_count, int G,
                          int32_t* __restrict__ positions) {
  extern __shared__ int32_t s_best[];  // [G]
  const int q = blockIdx.x;
  const int cnt = ranked_count ? min(ranked_count[q], P) : P;
  const int gcnt = target_count ? min(target_count[q], G) : G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) s_best[g] = INT32_MAX;
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t id = ranked[(int64_t)q * P + i];
    for (int g = 0; g < gcnt; ++g)
      if (targets[(int64_t)q * G + g] == id) atomicMin(&s_best[g], i);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x)
    positions[(int64_t)q * G + g] = (g < gcnt && s_best[g] != INT32_MAX) ? s_best[g] : -1;
}

// first_hit[q] = first j with query_index[q] in array[offsets[ranked[q, j]] : offsets[ranked[q, j] + 1]] or -1
// (ensemble_nqdpr.py:23-33)
__global__ void __launch_bounds__(ENS_THREADS)
ensemble_first_hit_kernel(const int64_t* __restrict__ ranked, const int32_t* __restrict__ ranked_count, int P,
                          const int64_t* __restrict__ query_index, const int32_t* __restrict__ offsets,
                          int64_t n_offsets, const int32_t* __restrict__ array, int32_t* __restrict__ first_hit) {
  __shared__ int s_first;
  const int q = blockIdx.x;
  const int cnt = ranked_count ? min(ranked_count[q], P) : P;
  const int64_t want = query_index[q];
  if (threadIdx.x == 0) s_first = INT32_MAX;
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    if (i > s_first) break;  // benign race: only skips work behind a hit already found
    const int64_t doc = ranked[(int64_t)q * P + i];
    if (doc < 0 || doc + 1 >= n_offsets) continue;
    for (int32_t t = offsets[doc]; t < offsets[doc + 1]; ++t)
      if ((int64_t)array[t] == want) { atomicMin(&s_first, i); break; }
  }
  __syncthreads();
  if (threadIdx.x == 0) first_hit[q] = s_first == INT32_MAX ? -1 : s_first;
}
}  // namespace

extern "C" int mevi_ensemble_cluster_ranks(mevi_ctx* ctx, const int64_t* cand_ids, const int32_t* cand_count, int nq,
                                           int P, const int32_t* codes, int64_t n_docs, int M,
                                           const int32_t* query_leaves, int L, int32_t* cranks, int32_t* num_leaves,
                                           void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  MEVI_REQUIRE(ctx, cand_ids && codes && query_leaves && cranks && num_leaves, "NULL argument");
  MEVI_REQUIRE(ctx, P > 0 && M > 0 && L > 0 && n_docs >= 0, "bad extents (P %d, M %d, L %d)", P, M, L);
  size_t smem = ((size_t)L * M + L) * sizeof(int32_t);
  MEVI_REQUIRE(ctx, smem <= 200 * 1024, "leaf list of %d x %d codes does not fit shared memory", L, M);
  if (nq <= 0) return MEVI_OK;
  if (smem > 48 * 1024)
    MEVI_CUDA(ctx, cudaFuncSetAttribute(ensemble_cluster_ranks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  ensemble_cluster_ranks_kernel<<<nq, ENS_THREADS, smem, (cudaStream_t)stream>>>(
      cand_ids, cand_count, P, codes, n_docs, M, query_leaves, L, cranks, num_leaves);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}

extern "C" int mevi_ensemble_fuse(mevi_ctx* ctx, const int64_t* cand_ids, const double* cand_scores,
                                  const int32_t* cranks, const int32_t* cand_count, int nq, int P, double alpha,
                                  double beta, double gamma, int num_leaves, int64_t* out_ids, double* out_scores,
                                  int32_t* out_count, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  MEVI_REQUIRE(ctx, cand_ids && cand_scores && cranks && out_ids && out_scores && out_count, "NULL argument");
  MEVI_REQUIRE(ctx, P > 0 && P <= ENS_MAX_P, "fusion takes lists of 1..%d candidates (got %d)", ENS_MAX_P, P);
  if (nq <= 0) return MEVI_OK;
  int P2 = 32;
  while (P2 < P) P2 <<= 1;
  size_t smem = (size_t)P2 * (8 + 8 + 8 + 4 + 4);
  if (smem > 48 * 1024)  // (per device: set on every call that needs it)
    MEVI_CUDA(ctx, cudaFuncSetAttribute(ensemble_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ensemble_fuse_kernel<<<nq, ENS_THREADS, smem, (cudaStream_t)stream>>>(
      cand_ids, cand_scores, cranks, cand_count, P, P2, alpha, beta, gamma, num_leaves, out_ids, out_scores, out_count);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}

extern "C" int mevi_ensemble_positions(mevi_ctx* ctx, const int64_t* ranked, const int32_t* ranked_count, int nq, int P,
                                       const int64_t* targets, const int32_t* target_count, int G, int32_t* positions,
                                       void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  MEVI_REQUIRE(ctx, ranked && targets && positions, "NULL argument");
  MEVI_REQUIRE(ctx, P > 0 && G > 0 && G <= 8192, "bad extents (P %d, G %d)", P, G);
  if (nq <= 0) return MEVI_OK;
  ensemble_positions_kernel<<<nq, ENS_THREADS, (size_t)G * sizeof(int32_t), (cudaStream_t)stream>>>(
      ranked, ranked_count, P, targets, target_count, G, positions);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}

extern "C" int mevi_ensemble_first_hit(mevi_ctx* ctx, const int64_t* ranked, const int32_t* ranked_count, int nq, int P,
                                       const int64_t* query_index, const int32_t* offsets, int64_t n_offsets,
                                       const int32_t* array, int32_t* first_hit, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  MEVI_REQUIRE(ctx, ranked && query_index && offsets && array && first_hit, "NULL argument");
  MEVI_REQUIRE(ctx, P > 0 && n_offsets >= 1, "bad extents (P %d)", P);
  if (nq <= 0) return MEVI_OK;
  ensemble_first_hit_kernel<<<nq, ENS_THREADS, 0, (cudaStream_t)stream>>>(ranked, ranked_count, P, query_index,
                                                                            offsets, n_offsets, array, first_hit);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  return MEVI_OK;
}
