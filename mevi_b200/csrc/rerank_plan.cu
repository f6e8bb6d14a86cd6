// rerank_plan.cu — the round plan of the leaf-grouped re-rank (K3g), made on the device.
// The reference visits (query, leaf) pairs one by one (main_models.py:3915-3947); K3g turns a call's pairs into work
// items (document tile of a leaf) x (groups of <= 64 queries that chose the leaf).  This file is the "tiles" plan of
// mevi_b200/rerank.py (plan_grouped_tile_rounds keeps the torch restatement for tests) without the ~120 small torch
// launches and host round trips that cost 2.7 ms of a 15.9 ms call:
//   classes: a pair's class is the number of boundaries boot[0..n_boot) that its leaf rank has reached;
//   set c < n_boot  (round c, a threshold sample): pairs of class c  x  the FIRST tile of their leaf;
//   set n_boot      (last round, part A):          pairs of the last class x the first tile of their leaf;
//   set n_boot + 1  (last round, part B):          ALL pairs x every further tile of their leaf.
// Per set and leaf: groups = ceil(pairs / 64), items = tiles * ceil(groups / maxg) (an item takes up to maxg consecutive
// groups, see grouped_gemm_kernel); one exclusive scan over [set][groups | items][leaf] gives every leaf its first group
// and item.  A pair's column is its arrival number among the leaf's pairs (atomic counter): which queries share a
// group does not influence any score.  One host synchronisation (the sizes the caller allocates).
#include <stdlib.h>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {
constexpr int PL_MAX_BOOT = 3;
constexpr int PL_MAX_SETS = PL_MAX_BOOT + 2;
constexpr int PL_GROUP = 64, PL_TILE = 128;

struct PlanParams {
  const int32_t* ql; int nq; int L;
  const int64_t* leaf_offsets; const int64_t* leaf_tile0; int64_t n_leaves;
  int n_boot; int boot[PL_MAX_BOOT]; int maxg[PL_MAX_SETS];
  int32_t* cnt;      // [n_sets][n_leaves] pairs per (set, leaf)
  int32_t* scan_in;  // [n_sets][2][n_leaves] groups | items, + 1 trailing zero
  int32_t* scan_out; // exclusive sums of scan_in
  int32_t* pos;      // [nq*L][2] arrival number of the pair in its class set / in the all-pairs set
  int32_t* seg;      // [2*n_sets + 1] scan_out at the segment starts (+ the grand total), then n_weak
};

__device__ __forceinline__ int pl_class(const PlanParams& p, int rank) {
  int c = 0;
  for (int b = 0; b < p.n_boot; ++b) c += rank >= p.boot[b];
  return c;
}

// one CTA per query: counts, arrival numbers, candidate totals, the "weak bootstrap" flag
__global__ void __launch_bounds__(128) plan_count_kernel(PlanParams p, int boot_min_rows, int k, int pass_budget, int32_t* __restrict__ ncand,
                                                         int32_t* __restrict__ weak) {
  const int q = blockIdx.x;
  const int n_sets = p.n_boot + 2;
  long long total = 0, boot_rows = 0;
  const int boot_last = p.n_boot ? p.boot[p.n_boot - 1] : 0;
  for (int r = threadIdx.x; r < p.L; r += blockDim.x) {
    const int64_t pair = (int64_t)q * p.L + r;
    const int leaf = p.ql[pair];
    if (leaf < 0 || leaf >= p.n_leaves) continue;
    const int c = pl_class(p, r);
    p.pos[2 * pair + 0] = atomicAdd(&p.cnt[(int64_t)c * p.n_leaves + leaf], 1);
    p.pos[2 * pair + 1] = atomicAdd(&p.cnt[(int64_t)(n_sets - 1) * p.n_leaves + leaf], 1);
    const long long size = p.leaf_offsets[leaf + 1] - p.leaf_offsets[leaf];
    total += size;
    if (r < boot_last) boot_rows += size < PL_TILE ? size : PL_TILE;
  }
  __shared__ long long s_tot[4], s_boot[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(MEVI_FULL_MASK, total, o);
    boot_rows += __shfl_xor_sync(MEVI_FULL_MASK, boot_rows, o);
  }
  if ((threadIdx.x & 31) == 0) { s_tot[threadIdx.x >> 5] = total; s_boot[threadIdx.x >> 5] = boot_rows; }
  __syncthreads();
  if (threadIdx.x == 0) {
    total = s_tot[0] + s_tot[1] + s_tot[2] + s_tot[3];
    boot_rows = s_boot[0] + s_boot[1] + s_boot[2] + s_boot[3];
    ncand[q] = (int32_t)(total > 0x7FFFFFFFll ? 0x7FFFFFFFll : total);
    // a sample of s rows lets about total * k / s candidates through the last round; they must fit the candidate buffer
    // (pass_budget of its 8,192 slots).  A query with few candidates needs a small sample or none at all.
    const long long need = total * k / pass_budget;
    const long long floor_rows = 2ll * k > need ? 2ll * k : need;
    const long long limit = boot_min_rows < floor_rows ? boot_min_rows : floor_rows;
    const int w = (total > pass_budget && boot_rows < limit && total > boot_rows) ? 1 : 0;
    weak[q] = w;
    if (w) atomicAdd(&p.seg[2 * n_sets + 1], 1);
  }
}

__device__ __forceinline__ int pl_tiles_of_set(const PlanParams& p, int set, int64_t leaf) {
  const int tpl = (int)(p.leaf_tile0[leaf + 1] - p.leaf_tile0[leaf]);
  return set == p.n_boot + 1 ? max(tpl - 1, 0) : min(tpl, 1);
}

__global__ void plan_leaf_kernel(PlanParams p) {
  const int64_t leaf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_sets = p.n_boot + 2;
  if (leaf == 0) p.scan_in[(int64_t)2 * n_sets * p.n_leaves] = 0;
  if (leaf >= p.n_leaves) return;
  for (int s = 0; s < n_sets; ++s) {
    const int tiles = pl_tiles_of_set(p, s, leaf);
    const int groups = tiles > 0 ? (p.cnt[(int64_t)s * p.n_leaves + leaf] + PL_GROUP - 1) / PL_GROUP : 0;
    const int wide = (groups + p.maxg[s] - 1) / p.maxg[s];
    p.scan_in[((int64_t)2 * s + 0) * p.n_leaves + leaf] = groups;
    p.scan_in[((int64_t)2 * s + 1) * p.n_leaves + leaf] = tiles * wide;
  }
}

__global__ void plan_segments_kernel(PlanParams p) {
  const int n_sets = p.n_boot + 2;
  const int j = threadIdx.x;
  if (j <= 2 * n_sets) p.seg[j] = p.scan_out[(int64_t)j * p.n_leaves];
}

struct FillParams {
  int n_parts;            // 1 (a sample round) or 2 (the last round: part A then part B)
  int set[2];
  int32_t group_off[2];   // first group of the part inside the round
  int32_t item_off[2];    // first item of the part inside the round
  int32_t n_items[2];
  int32_t* item_tile; int32_t* item_group; int32_t* group_qid;
};

// one thread per (query, rank) pair: its column in the round's groups
__global__ void plan_fill_groups_kernel(PlanParams p, FillParams f) {
  const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= (int64_t)p.nq * p.L) return;
  const int leaf = p.ql[pair];
  if (leaf < 0 || leaf >= p.n_leaves) return;
  const int q = (int)(pair / p.L), r = (int)(pair % p.L);
  const int c = pl_class(p, r);
  for (int part = 0; part < f.n_parts; ++part) {
    const int s = f.set[part];
    const bool all_pairs = s == p.n_boot + 1;
    if (!all_pairs && s != c) continue;
    if (pl_tiles_of_set(p, s, leaf) <= 0) continue;
    const int arrival = p.pos[2 * pair + (all_pairs ? 1 : 0)];
    const int64_t g = f.group_off[part] + (p.scan_out[((int64_t)2 * s) * p.n_leaves + leaf] - p.seg[2 * s]) + arrival / PL_GROUP;
    f.group_qid[g * PL_GROUP + arrival % PL_GROUP] = q;
  }
}

// one thread per work item of the round: leaf by binary search in the items scan, tile-major inside the leaf
__global__ void plan_fill_items_kernel(PlanParams p, FillParams f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)f.n_items[0] + (f.n_parts > 1 ? f.n_items[1] : 0);
  if (i >= total) return;
  const int part = (f.n_parts > 1 && i >= f.n_items[0]) ? 1 : 0;
  const int s = f.set[part];
  const int32_t li = (int32_t)(i - (part ? f.n_items[0] : 0));
  const int32_t* item0 = p.scan_out + ((int64_t)2 * s + 1) * p.n_leaves;
  const int32_t base = p.seg[2 * s + 1];
  int64_t lo = 0, hi = p.n_leaves;  // last leaf with item0[leaf] - base <= li
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (item0[mid] - base <= li) lo = mid; else hi = mid;
  }
  const int64_t leaf = lo;
  const int groups = p.scan_in[((int64_t)2 * s) * p.n_leaves + leaf];
  const int maxg = p.maxg[s];
  const int wide = (groups + maxg - 1) / maxg;
  const int loc = li - (item0[leaf] - base);
  const int t = loc / wide, j = loc % wide;
  const int32_t grp0 = f.group_off[part] + (p.scan_out[((int64_t)2 * s) * p.n_leaves + leaf] - p.seg[2 * s]);
  const int ng = min(maxg, groups - j * maxg);
  f.item_tile[f.item_off[part] + li] = (int32_t)(p.leaf_tile0[leaf] + (s == p.n_boot + 1 ? 1 : 0) + t);
  f.item_group[f.item_off[part] + li] = (grp0 + j * maxg) | (ng << 24);
}

struct PlanHost {  // plain data: lives in mevi_ctx::gr_plan (one grouped call at a time per context)
  PlanParams p;
  int32_t seg[2 * PL_MAX_SETS + 2];
  int valid;
};
PlanHost* plan_of(mevi_ctx* ctx) {
  if (!ctx->gr_plan) ctx->gr_plan = calloc(1, sizeof(PlanHost));
  return (PlanHost*)ctx->gr_plan;
}

size_t pl_layout(PlanParams* p, char* ws) {
  const int n_sets = p->n_boot + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t o_cnt = take((size_t)n_sets * p->n_leaves * 4), o_in = take(((size_t)2 * n_sets * p->n_leaves + 1) * 4),
               o_out = take(((size_t)2 * n_sets * p->n_leaves + 1) * 4), o_pos = take((size_t)p->nq * p->L * 8),
               o_seg = take((2 * PL_MAX_SETS + 2) * 4);
  if (ws) {
    p->cnt = (int32_t*)(ws + o_cnt); p->scan_in = (int32_t*)(ws + o_in); p->scan_out = (int32_t*)(ws + o_out);
    p->pos = (int32_t*)(ws + o_pos); p->seg = (int32_t*)(ws + o_seg);
  }
  return off;
}
}  // namespace

extern "C" int mevi_rerank_grouped_plan(mevi_ctx* ctx, const int32_t* ql, int nq, int L, const int64_t* leaf_offsets,
                                        const int64_t* leaf_tile0, int64_t n_leaves, const int32_t* boot_leaves,
                                        int n_boot, int boot_min_rows, int k, int pass_budget, int max_groups_sample,
                                        int max_groups_last, int32_t* ncand, int32_t* weak, int64_t* sizes_host, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, ql && leaf_offsets && leaf_tile0 && ncand && weak && sizes_host, "NULL argument");
  MEVI_REQUIRE(ctx, nq > 0 && L > 0 && k > 0 && pass_budget > 0 && n_leaves > 0 && n_leaves < (int64_t)1 << 27, "bad extents");
  MEVI_REQUIRE(ctx, n_boot >= 0 && n_boot <= PL_MAX_BOOT && (n_boot == 0 || boot_leaves), "0..%d bootstrap boundaries", PL_MAX_BOOT);
  MEVI_REQUIRE(ctx, (max_groups_sample == 1 || max_groups_sample == 2 || max_groups_sample == 4) &&
                        (max_groups_last == 1 || max_groups_last == 2 || max_groups_last == 4),
               "an item takes 1, 2 or 4 groups");
  PlanHost* hp = plan_of(ctx);
  if (!hp) return MEVI_ERR_NOMEM;
  PlanHost& h = *hp;
  h.valid = 0;
  PlanParams& p = h.p;
  p.ql = ql; p.nq = nq; p.L = L; p.leaf_offsets = leaf_offsets; p.leaf_tile0 = leaf_tile0; p.n_leaves = n_leaves;
  p.n_boot = n_boot;
  for (int b = 0; b < n_boot; ++b) {
    p.boot[b] = boot_leaves[b];
    MEVI_REQUIRE(ctx, p.boot[b] > (b ? p.boot[b - 1] : 0), "bootstrap boundaries must ascend");
  }
  const int n_sets = n_boot + 2;
  for (int s = 0; s < n_sets; ++s) p.maxg[s] = s < n_boot ? max_groups_sample : max_groups_last;
  const size_t bytes = pl_layout(&p, nullptr);
  char* ws = (char*)mevi_ws(ctx, WS_GR_PLAN, bytes);
  if (!ws) return MEVI_ERR_NOMEM;
  pl_layout(&p, ws);
  const int64_t scan_len = (int64_t)2 * n_sets * n_leaves + 1;
  size_t tmp_bytes = 0;
  MEVI_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, p.scan_in, p.scan_out, (int)scan_len, st));
  void* tmp = mevi_ws(ctx, WS_SORT_TMP, tmp_bytes);
  if (!tmp) return MEVI_ERR_NOMEM;
  MEVI_CUDA(ctx, cudaMemsetAsync(p.cnt, 0, (size_t)n_sets * n_leaves * 4, st));
  MEVI_CUDA(ctx, cudaMemsetAsync(p.seg, 0, (2 * PL_MAX_SETS + 2) * 4, st));
  plan_count_kernel<<<nq, 128, 0, st>>>(p, boot_min_rows, k, pass_budget, ncand, weak);
  plan_leaf_kernel<<<(unsigned)((n_leaves + 255) / 256), 256, 0, st>>>(p);
  MEVI_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, p.scan_in, p.scan_out, (int)scan_len, st));
  plan_segments_kernel<<<1, 32, 0, st>>>(p);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 4);
  MEVI_CUDA(ctx, cudaMemcpyAsync(h.seg, p.seg, sizeof(h.seg), cudaMemcpyDeviceToHost, st));
  MEVI_CUDA(ctx, cudaStreamSynchronize(st));
  if (int drc = mevi_deferred_error(ctx)) return drc;
  // sizes_host: [round][items, groups] for the n_boot + 1 rounds, then the number of weak queries
  for (int r = 0; r <= n_boot; ++r) {
    int64_t items = 0, groups = 0;
    for (int s = r; s < (r == n_boot ? n_sets : r + 1); ++s) {
      groups += h.seg[2 * s + 1] - h.seg[2 * s];
      items += h.seg[2 * s + 2] - h.seg[2 * s + 1];
    }
    sizes_host[2 * r] = items;
    sizes_host[2 * r + 1] = groups;
  }
  sizes_host[2 * (n_boot + 1)] = h.seg[2 * n_sets + 1];
  h.valid = 1;
  return MEVI_OK;
}

extern "C" int mevi_rerank_grouped_plan_fill(mevi_ctx* ctx, int round, int32_t* item_tile, int32_t* item_group,
                                             int32_t* group_qid, void* stream) {
  MEVI_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  MEVI_REQUIRE(ctx, ctx->gr_plan && ((PlanHost*)ctx->gr_plan)->valid, "no plan: call mevi_rerank_grouped_plan first");
  PlanHost& h = *(PlanHost*)ctx->gr_plan;
  const PlanParams& p = h.p;
  MEVI_REQUIRE(ctx, round >= 0 && round <= p.n_boot, "round %d of %d", round, p.n_boot + 1);
  FillParams f;
  f.n_parts = round == p.n_boot ? 2 : 1;
  int64_t items = 0, groups = 0;
  for (int part = 0; part < f.n_parts; ++part) {
    const int s = round + part;
    f.set[part] = s;
    f.group_off[part] = (int32_t)groups;
    f.item_off[part] = (int32_t)items;
    f.n_items[part] = h.seg[2 * s + 2] - h.seg[2 * s + 1];
    groups += h.seg[2 * s + 1] - h.seg[2 * s];
    items += f.n_items[part];
  }
  if (items <= 0 || groups <= 0) return MEVI_OK;
  MEVI_REQUIRE(ctx, item_tile && item_group && group_qid, "NULL argument");
  f.item_tile = item_tile; f.item_group = item_group; f.group_qid = group_qid;
  MEVI_CUDA(ctx, cudaMemsetAsync(group_qid, 0xFF, (size_t)groups * PL_GROUP * 4, st));
  const int64_t pairs = (int64_t)p.nq * p.L;
  plan_fill_groups_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(p, f);
  plan_fill_items_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(p, f);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 2);
  return MEVI_OK;
}
