// rq_tensor.cu — K1: tcgen05/TMEM residual-quantisation encode + k-means assignment.
//
// Replaces the arithmetic of MEVI/pq.py:281-305 (and the assignment half of 551-598) for the
// shipped shape family (d % 64 == 0, K % 32 == 0, M*K <= 128).
//
// Algebraic form.  The reference subtracts the chosen centroid after every level, so level j sees
// r_j = x - sum_{m<j} c^m_{k_m}.  Distances to level-j centroids only need
//     r_j . c^j_k = x . c^j_k - sum_{m<j} (c^m_{k_m} . c^j_k)
// so ONE contraction X[rows,d] . C_all[M*K,d]^T serves all levels; the second term comes from a
// precomputed cross-level Gram table.  X is read from HBM exactly once (4*d bytes per row).
//
// Precision.  fp32 inputs are split x*2^s = hi + lo with hi, lo in fp16 (22 significant bits) and the
// contraction is evaluated as hi.hi + hi.lo + lo.hi on the tensor cores with fp32 accumulation in TMEM
// (kind::f16, 1.5x the work of a TF32 pass; measured error ~2^-22 |x||c|, see profiles/probe_r01.txt).
// The tensor result is only a PREFILTER: per (row, level) the best and second-best distances are
// compared against a rigorous error bound; rows whose gap is inside the bound are appended to a work
// list and re-decided from that level on by the fp32 direct-form kernel of rq_exact.cu, the literal
// restatement of the reference arithmetic.  Rows outside the bound provably have the same argmin in
// exact arithmetic, so codes agree with the reference except at fp32-epsilon ties.
//
// Pipeline of one persistent CTA (512 threads, 1 CTA/SM, tile = 128 rows):
//   warps 4-11  converters: coalesced 16 B loads of X straight from global (L1 no-allocate, next chunk
//               prefetched in registers) -> scale, split, write the hi|lo fp16 operand tiles into a
//               3-stage shared-memory ring in the UMMA K-major 128B-swizzle layout; row norms on the fly
//   warp 0      B producer: one bulk async copy (TMA engine) per 64-wide K chunk of the pre-swizzled
//               [C_hi | C_lo] image (L2 resident) into its own 3-stage ring, mbarrier complete_tx
//   warp 1      one thread issues tcgen05.mma: A_hi x [C_hi|C_lo] (N = 2*M*K) and A_lo x C_hi (N = M*K),
//               fp32 accumulators double-buffered in TMEM (2 x 256 columns); tcgen05.commit frees stages
//   warps 12-15 epilogue: tcgen05.ld the tile's accumulators, greedy per-level argmin with Gram
//               corrections, error-bound test, code store, work-list append, TMEM buffer release
// Roofline: HBM (4*d B/row); tensor work is 72 cycles/row/SM, shared-memory traffic ~13.5 KB/row.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

int mevi_rq_exact_launch(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                         int32_t* codes, int64_t codes_stride, float* residual, const int32_t* work_rows,
                         const int32_t* work_levels, const int64_t* n_work_dev, int64_t n_items, double* inertia,
                         cudaStream_t st);

namespace {

constexpr int TM = 128;                      // rows per tile (UMMA M)
constexpr int KC = 64;                       // K elements per chunk: 64 fp16 = one 128-byte swizzle row
constexpr int NSA = 3, NSB = 3;              // ring depths
constexpr int THREADS = 512;
constexpr int CONV_WARP0 = 4, CONV_WARPS = 8, EPI_WARP0 = 12;
constexpr int A_TILE_BYTES = TM * 128;       // one fp16 operand tile (hi or lo)
constexpr int A_STAGE_BYTES = 2 * A_TILE_BYTES;
constexpr int TMEM_COLS = 512, TMEM_BUF_COLS = 256;
constexpr float U_REL = 1.0f / 524288.0f;    // 2^-19: relative bound on the split-fp16 contraction error (|x||c| units)

enum { C_SC = 0, C_SX, C_INV, C_INV_SX2, C_FX, C_FC, C_NUM = 8 };

struct Params {
  const float* X; int64_t n; int d; int nchunks;
  int M, K, NT, N1, metric;
  const __half* Bimg; const float* cn2; const float* cnorm; const float* gram; const float* consts;
  int gram_floats;
  int32_t* codes; int64_t codes_stride;
  int32_t* work_rows; int32_t* work_levels; unsigned long long* work_count;
  double* inertia; int* err_flag;
  int64_t n_tiles;
};

struct SmemLayout {
  int a_off, b_off, gram_off, cn2_off, cnorm_off, stats_off, bar_off, holder_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int M, int K, int NT, int N1) {
  SmemLayout L;
  L.a_off = 0;
  L.b_off = L.a_off + NSA * A_STAGE_BYTES;
  L.gram_off = L.b_off + NSB * N1 * 128;
  int gram_pad = 0;
  for (int j = 1; j < M; ++j) gram_pad += j * K * (K + 1);
  L.cn2_off = L.gram_off + gram_pad * 4;
  L.cnorm_off = L.cn2_off + NT * 4;
  L.stats_off = L.cnorm_off + NT * 4;
  L.bar_off = (L.stats_off + 2 * TM * 4 + 7) & ~7;
  L.holder_off = L.bar_off + (2 * NSA + 2 * NSB + 6) * 8;
  L.total = L.holder_off + 16;
  return L;
}

__global__ void __launch_bounds__(THREADS, 1) rq_tensor_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout L = smem_layout(p.M, p.K, p.NT, p.N1);
  uint8_t* sA = smem + L.a_off;
  uint8_t* sB = smem + L.b_off;
  float* sGram = reinterpret_cast<float*>(smem + L.gram_off);
  float* sCn2 = reinterpret_cast<float*>(smem + L.cn2_off);
  float* sCnorm = reinterpret_cast<float*>(smem + L.cnorm_off);
  float* sStats = reinterpret_cast<float*>(smem + L.stats_off);  // [2][TM] squared row norms
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NSA;
  uint64_t* b_full = a_empty + NSA;
  uint64_t* b_empty = b_full + NSB;
  uint64_t* acc_full = b_empty + NSB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* st_full = acc_empty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L.holder_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, NT = p.NT, N1 = p.N1;
  const uint32_t b_stage_bytes = (uint32_t)N1 * 128u;

  // ---- one-time setup ---------------------------------------------------------------------
  {  // Gram table: global rows of K floats -> shared rows padded to K+1 (bank-conflict-free per-thread rows)
    const int rows = p.gram_floats / K;
    for (int i = tid; i < p.gram_floats; i += THREADS) {
      const int r = i / K, c = i - r * K;
      sGram[r * (K + 1) + c] = p.gram[i];
    }
    (void)rows;
    for (int i = tid; i < NT; i += THREADS) {
      sCn2[i] = p.cn2[i];
      sCnorm[i] = p.cnorm[i];
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSA; ++s) { ptx::mbar_init(&a_full[s], CONV_WARPS); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NSB; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 4); ptx::mbar_init(&st_full[b], CONV_WARPS); }
    ptx::mbar_fence_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_holder, TMEM_COLS);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  const int64_t first_tile = blockIdx.x;
  const int64_t tile_stride = gridDim.x;
  const int nchunks = p.nchunks;

  if (warp == 0) {
    // ===== B producer =========================================================================
    if (lane == 0) {
      uint32_t g = 0;
      bool ok = true;
      for (int64_t tile = first_tile; tile < p.n_tiles && ok; tile += tile_stride) {
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t s = g % NSB, ph = (g / NSB) & 1;
          if (!ptx::mbar_wait(&b_empty[s], ph ^ 1)) { atomicExch(p.err_flag, 1); ok = false; break; }
          ptx::mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
          ptx::bulk_g2s(sB + (size_t)s * b_stage_bytes, p.Bimg + (size_t)c * N1 * KC, b_stage_bytes, &b_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===========================================================================
    if (lane == 0) {
      const uint32_t idesc_n1 = ptx::umma_idesc_f16_m128((uint32_t)N1);
      const uint32_t idesc_nt = ptx::umma_idesc_f16_m128((uint32_t)NT);
      uint32_t g = 0, it = 0;
      bool ok = true;
      for (int64_t tile = first_tile; tile < p.n_tiles && ok; tile += tile_stride, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        if (!ptx::mbar_wait(&acc_empty[buf], ph ^ 1)) { atomicExch(p.err_flag, 2); ok = false; break; }
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * TMEM_BUF_COLS;
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t sa = g % NSA, pa = (g / NSA) & 1, sb = g % NSB, pb = (g / NSB) & 1;
          if (!ptx::mbar_wait(&a_full[sa], pa) || !ptx::mbar_wait(&b_full[sb], pb)) { atomicExch(p.err_flag, 3); ok = false; break; }
          ptx::tc_fence_after_sync();
          const uint32_t a_hi = ptx::smem_u32(sA + (size_t)sa * A_STAGE_BYTES);
          const uint32_t a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_ad = ptx::smem_u32(sB + (size_t)sb * b_stage_bytes);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)
            ptx::umma_f16(d_tmem, ptx::umma_desc_sw128(a_hi + ks * 32), ptx::umma_desc_sw128(b_ad + ks * 32), idesc_n1,
                          (c | ks) != 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)
            ptx::umma_f16(d_tmem, ptx::umma_desc_sw128(a_lo + ks * 32), ptx::umma_desc_sw128(b_ad + ks * 32), idesc_nt, 1u);
          ptx::umma_commit(&a_empty[sa]);
          ptx::umma_commit(&b_empty[sb]);
        }
        if (ok) ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= CONV_WARP0 && warp < CONV_WARP0 + CONV_WARPS) {
    // ===== converters ===========================================================================
    const int cw = warp - CONV_WARP0;
    const int half = lane >> 4, l16 = lane & 15;
    const float sx = p.consts[C_SX], inv_sx2 = p.consts[C_INV_SX2];
    int64_t my_tiles = 0;
    if (first_tile < p.n_tiles) my_tiles = (p.n_tiles - first_tile + tile_stride - 1) / tile_stride;
    const int64_t total_chunks = my_tiles * nchunks;
    float4 cur[8], nxt[8];
    float norm[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) norm[q] = 0.f;

    auto load_chunk = [&](int64_t gg, float4 (&v)[8]) {
      const int64_t it = gg / nchunks;
      const int c = (int)(gg - it * nchunks);
      const int64_t row0 = (first_tile + it * tile_stride) * TM;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int64_t row = row0 + 2 * (cw + 8 * q) + half;
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.n) v[q] = ld_stream_f4(p.X + row * p.d + c * KC + l16 * 4);
      }
    };
    bool ok = true;
    if (total_chunks > 0) load_chunk(0, cur);
    for (int64_t gg = 0; gg < total_chunks && ok; ++gg) {
      if (gg + 1 < total_chunks) load_chunk(gg + 1, nxt);
      const uint32_t sa = (uint32_t)(gg % NSA), pa = (uint32_t)((gg / NSA) & 1);
      if (!ptx::mbar_wait(&a_empty[sa], pa ^ 1)) { atomicExch(p.err_flag, 4); ok = false; break; }
      uint8_t* stage = sA + (size_t)sa * A_STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int rl = 2 * (cw + 8 * q) + half;
        const float t0 = cur[q].x * sx, t1 = cur[q].y * sx, t2 = cur[q].z * sx, t3 = cur[q].w * sx;
        norm[q] = fmaf(t0, t0, norm[q]);
        norm[q] = fmaf(t1, t1, norm[q]);
        norm[q] = fmaf(t2, t2, norm[q]);
        norm[q] = fmaf(t3, t3, norm[q]);
        const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
        const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(t0 - b01.x, t1 - b01.y), l23 = __floats2half2_rn(t2 - b23.x, t3 - b23.y);
        const uint32_t off = (uint32_t)rl * 128u + ((uint32_t)((l16 >> 1) ^ (rl & 7)) << 4) + ((uint32_t)(l16 & 1) << 3);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01);
        hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01);
        lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(stage + off) = hv;
        *reinterpret_cast<uint2*>(stage + A_TILE_BYTES + off) = lv;
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a_full[sa]);
      const int64_t it = gg / nchunks;
      if (gg - it * nchunks == nchunks - 1) {
        // tile finished: publish squared row norms for the epilogue
        const uint32_t buf = (uint32_t)(it & 1), ph = (uint32_t)((it >> 1) & 1);
        if (!ptx::mbar_wait(&acc_empty[buf], ph ^ 1)) { atomicExch(p.err_flag, 5); ok = false; break; }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v = norm[q];
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 8);
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 4);
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 2);
          v += __shfl_xor_sync(MEVI_FULL_MASK, v, 1);
          if (l16 == 0) sStats[buf * TM + 2 * (cw + 8 * q) + half] = v * inv_sx2;
          norm[q] = 0.f;
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&st_full[buf]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue ===============================================================================
    const int ew = warp - EPI_WARP0;  // == warp % 4: TMEM lanes 32*ew .. 32*ew+31
    const int rl = ew * 32 + lane;
    const float inv = p.consts[C_INV], fx = p.consts[C_FX], fc = p.consts[C_FC];
    const bool l2 = p.metric == MEVI_METRIC_L2;
    double inertia_acc = 0.0;
    uint32_t it = 0;
    bool ok = true;
    for (int64_t tile = first_tile; tile < p.n_tiles && ok; tile += tile_stride, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      if (!ptx::mbar_wait(&acc_full[buf], ph) || !ptx::mbar_wait(&st_full[buf], ph)) { atomicExch(p.err_flag, 6); ok = false; break; }
      ptx::tc_fence_after_sync();
      const float xn2 = sStats[buf * TM + rl];
      const float xn = sqrtf(xn2);
      const uint32_t taddr = tmem_base + buf * TMEM_BUF_COLS + ((uint32_t)(ew * 32) << 16);
      const int64_t row = tile * TM + rl;
      int code[8];
      int flag_level = -1;
      float last_best = 0.f;
      int goff = 0;  // offset of level j's Gram block inside sGram
      for (int j = 0; j < p.M; ++j) {
        float best = CUDART_INF_F, second = CUDART_INF_F, best_err = 0.f, second_err = 0.f;
        int besti = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
          uint32_t rm[32], rc[32];
          ptx::tmem_ld32(taddr + j * K + k0, rm);
          ptx::tmem_ld32(taddr + NT + j * K + k0, rc);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const int k = k0 + kk;
            const float a = (__uint_as_float(rm[kk]) + __uint_as_float(rc[kk])) * inv;
            float g = 0.f;
            for (int m = 0; m < j; ++m) g += sGram[goff + (m * K + code[m]) * (K + 1) + k];
            const float cnk = sCnorm[j * K + k];
            const float dot = a - g;
            const float dist = l2 ? fmaf(-2.f, dot, sCn2[j * K + k]) : -dot;
            // bound on |dot - exact|: contraction + representation floors + fp32 epilogue arithmetic
            const float err = U_REL * xn * cnk + fx * cnk + fc * xn + 2.4e-7f * (fabsf(a) + fabsf(g) + (l2 ? sCn2[j * K + k] : 0.f));
            if (dist < best) {
              second = best; second_err = best_err;
              best = dist; best_err = err; besti = k;
            } else if (dist < second) {
              second = dist; second_err = err;
            }
          }
        }
        code[j] = besti;
        const float margin = (l2 ? 2.f : 1.f) * (best_err + second_err);
        const bool clear = (second - best) > margin;  // NaN/inf -> not clear -> exact path decides
        if (!clear && flag_level < 0) flag_level = j;
        last_best = best;
        goff += j > 0 ? j * K * (K + 1) : 0;
        if (j == 0) goff = 0;
      }
      if (row < p.n) {
        int32_t* dst = p.codes + row * p.codes_stride;
        if (p.M == 4 && p.codes_stride == 4) {
          *reinterpret_cast<int4*>(dst) = make_int4(code[0], code[1], code[2], code[3]);
        } else {
          for (int j = 0; j < p.M; ++j) dst[j] = code[j];
        }
        if (flag_level >= 0) {
          const unsigned long long slot = atomicAdd(p.work_count, 1ull);
          p.work_rows[slot] = (int32_t)row;
          p.work_levels[slot] = flag_level;
        }
        if (p.inertia) inertia_acc += (double)(l2 ? fmaxf(last_best + xn2, 0.f) : -last_best);
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
    if (p.inertia) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) inertia_acc += __shfl_xor_sync(MEVI_FULL_MASK, inertia_acc, o);
      if (lane == 0 && inertia_acc != 0.0) atomicAdd(p.inertia, inertia_acc);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- preparation kernels ---------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ p, int64_t rows, int d, int64_t row_step, unsigned* out) {
  // max |v| over rows 0, row_step, 2*row_step, ...   (non-negative floats order like their bit patterns)
  unsigned m = 0;
  const int64_t nsel = (rows + row_step - 1) / row_step;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsel * d; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = (i / d) * row_step;
    const float v = fabsf(p[r * d + (i % d)]);
    if (v == v && v < CUDART_INF_F) m = max(m, __float_as_uint(v));
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(MEVI_FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__device__ __forceinline__ float pow2_scale(float amax) {
  // 2^s with amax * 2^s in [2^13, 2^14): two bits of headroom below the fp16 maximum
  if (!(amax > 0.f)) return 1.f;
  return ldexpf(1.f, 13 - ilogbf(amax));
}

__global__ void consts_kernel(const unsigned* absmax2, int d, float* consts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float sc = pow2_scale(__uint_as_float(absmax2[0]));
    const float sx = pow2_scale(__uint_as_float(absmax2[1]));
    consts[C_SC] = sc;
    consts[C_SX] = sx;
    consts[C_INV] = 1.f / (sc * sx);
    consts[C_INV_SX2] = (1.f / sx) * (1.f / sx);
    const float floor_abs = sqrtf((float)d) * 5.9604645e-8f;  // sqrt(d) * 2^-24: fp16 subnormal spacing of hi+lo
    consts[C_FX] = floor_abs / sx;
    consts[C_FC] = floor_abs / sc;
  }
}

// Bimg[chunk][row][64 halfs], rows 0..NT-1 = hi(c*sc), NT..2NT-1 = lo; 16-byte units XOR-swizzled by (row & 7)
__global__ void bimg_kernel(const float* __restrict__ cb, int rows_valid, int d, int NT, const float* __restrict__ consts,
                            __half* __restrict__ Bimg) {
  const int units_per_row = d / 8;
  const int total = NT * units_per_row;
  const float sc = consts[C_SC];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / units_per_row, ug = i - r * units_per_row;
    const int chunk = ug / 8, u = ug & 7;
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = r < rows_valid ? cb[(size_t)r * d + ug * 8 + e] * sc : 0.f;
      hi[e] = __float2half_rn(t);
      lo[e] = __float2half_rn(t - __half2float(hi[e]));
    }
    const size_t base = (size_t)chunk * (2 * NT) * KC;
    const int up = u ^ (r & 7);
    *reinterpret_cast<uint4*>(Bimg + base + (size_t)r * KC + up * 8) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(Bimg + base + (size_t)(NT + r) * KC + up * 8) = *reinterpret_cast<const uint4*>(lo);
  }
}

// one warp per centroid row: squared norm (double accumulation)
__global__ void cnorm_kernel(const float* __restrict__ cb, int rows_valid, int d, int NT, float* __restrict__ cn2,
                             float* __restrict__ cnorm) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= NT) return;
  double s = 0.0;
  if (r < rows_valid)
    for (int c = threadIdx.x & 31; c < d; c += 32) {
      const double v = cb[(size_t)r * d + c];
      s += v * v;
    }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(MEVI_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) {
    cn2[r] = (float)s;
    cnorm[r] = (float)sqrt(s);
  }
}

// gram[(level j block) + prow*K + k] = c_prow . c^j_k   for prow in [0, j*K); one warp per entry
__global__ void gram_kernel(const float* __restrict__ cb, int M, int K, int d, float* __restrict__ gram, int total) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= total) return;
  int j = 1, off = 0;
  while (e >= off + j * K * K) {
    off += j * K * K;
    ++j;
  }
  const int loc = e - off;
  const int prow = loc / K, k = loc - prow * K;
  const float* a = cb + (size_t)prow * d;
  const float* b = cb + ((size_t)j * K + k) * d;
  double s = 0.0;
  for (int c = threadIdx.x & 31; c < d; c += 32) s += (double)a[c] * (double)b[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(MEVI_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) gram[e] = (float)s;
}

__global__ void finish_stats_kernel(const unsigned long long* work_count, int64_t rows, int64_t* stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && stats) {
    atomicAdd((unsigned long long*)&stats[0], *work_count);
    atomicAdd((unsigned long long*)&stats[1], (unsigned long long)rows);
  }
}

// residual[row] = ((x - c0) - c1) - ...   in the reference's order (pq.py:304-305)
__global__ void residual_from_codes_kernel(const float* __restrict__ X, int64_t n, int d4, const float* __restrict__ cb,
                                           int M, int K, const int32_t* __restrict__ codes, int64_t codes_stride,
                                           float* __restrict__ residual) {
  const int64_t total = n * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const int c = (int)(i - row * d4);
    float4 v = reinterpret_cast<const float4*>(X)[i];
    for (int m = 0; m < M; ++m) {
      const int code = codes[row * codes_stride + m];
      const float4 cc = __ldg(reinterpret_cast<const float4*>(cb) + ((int64_t)m * K + code) * d4 + c);
      v.x -= cc.x; v.y -= cc.y; v.z -= cc.z; v.w -= cc.w;
    }
    reinterpret_cast<float4*>(residual)[i] = v;
  }
}

}  // namespace

bool mevi_rq_tensor_supported(mevi_ctx* ctx, int d, int M, int K, int metric) {
  if (!ctx || ctx->cc_major != 10) return false;
  if (d < KC || d % KC != 0 || d > 8192) return false;
  if (M < 1 || M > 8 || K < 32 || K % 32 != 0) return false;
  const int NT = M * K;
  if (NT > 128 || NT % 16 != 0) return false;
  if (metric != MEVI_METRIC_L2 && metric != MEVI_METRIC_IP) return false;
  const SmemLayout L = smem_layout(M, K, NT, 2 * NT);
  return L.total + 1024 <= 227 * 1024;
}

int mevi_rq_tensor_assign(mevi_ctx* ctx, const float* X, int64_t n, int d, const float* cb, int M, int K, int metric,
                          int32_t* codes, int64_t codes_stride, float* residual, int64_t* stats, double* inertia,
                          cudaStream_t st) {
  if (!mevi_rq_tensor_supported(ctx, d, M, K, metric))
    return mevi_set_error(ctx, MEVI_ERR_UNSUPPORTED, "tensor RQ path unsupported for d=%d M=%d K=%d", d, M, K);
  MEVI_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0,
               "X and codebook must be 16-byte aligned");
  if (n <= 0) return MEVI_OK;
  // a protocol time-out in an earlier launch is reported here (deferred so calls stay asynchronous)
  if (ctx->pinned[3] && *reinterpret_cast<volatile int*>(ctx->pinned[3]) != 0) {
    const int code = *reinterpret_cast<volatile int*>(ctx->pinned[3]);
    *reinterpret_cast<volatile int*>(ctx->pinned[3]) = 0;
    return mevi_set_error(ctx, MEVI_ERR_CUDA, "tensor RQ kernel reported a pipeline time-out (code %d) in a previous call", code);
  }
  const int NT = M * K, N1 = 2 * NT, nchunks = d / KC;
  int gram_floats = 0;
  for (int j = 1; j < M; ++j) gram_floats += j * K * K;
  // ---- scratch: [consts | absmax2 | err | work_count | cn2 | cnorm | gram | Bimg]
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t o_consts = take(C_NUM * 4), o_abs = take(8), o_err = take(4), o_cnt = take(8), o_cn2 = take(NT * 4),
               o_cnorm = take(NT * 4), o_gram = take((size_t)(gram_floats ? gram_floats : 1) * 4),
               o_bimg = take((size_t)nchunks * N1 * KC * 2);
  char* ws = (char*)mevi_ws(ctx, WS_RQ_PREP, off);
  if (!ws) return MEVI_ERR_NOMEM;
  float* consts = (float*)(ws + o_consts);
  unsigned* absmax2 = (unsigned*)(ws + o_abs);
  int* err_flag = (int*)(ws + o_err);
  unsigned long long* work_count = (unsigned long long*)(ws + o_cnt);
  float* cn2 = (float*)(ws + o_cn2);
  float* cnorm = (float*)(ws + o_cnorm);
  float* gram = (float*)(ws + o_gram);
  __half* Bimg = (__half*)(ws + o_bimg);
  int32_t* work = (int32_t*)mevi_ws(ctx, WS_RQ_WORK, (size_t)n * 8);
  if (!work) return MEVI_ERR_NOMEM;
  int* host_err = (int*)mevi_pinned(ctx, 3, 64);
  if (!host_err) return MEVI_ERR_NOMEM;

  MEVI_CUDA(ctx, cudaMemsetAsync(ws + o_abs, 0, o_cn2 - o_abs, st));  // absmax2, err flag, work count
  absmax_kernel<<<32, 256, 0, st>>>(cb, (int64_t)M * K, d, 1, absmax2);
  const int64_t sample_rows = 8192;
  const int64_t row_step = n > sample_rows ? n / sample_rows : 1;
  absmax_kernel<<<ctx->sm_count, 256, 0, st>>>(X, n, d, row_step, absmax2 + 1);
  consts_kernel<<<1, 32, 0, st>>>(absmax2, d, consts);
  bimg_kernel<<<(NT * (d / 8) + 255) / 256, 256, 0, st>>>(cb, M * K, d, NT, consts, Bimg);
  cnorm_kernel<<<(NT + 7) / 8, 256, 0, st>>>(cb, M * K, d, NT, cn2, cnorm);
  if (gram_floats) gram_kernel<<<(gram_floats + 7) / 8, 256, 0, st>>>(cb, M, K, d, gram, gram_floats);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, gram_floats ? 6 : 5);

  Params p;
  p.X = X; p.n = n; p.d = d; p.nchunks = nchunks; p.M = M; p.K = K; p.NT = NT; p.N1 = N1; p.metric = metric;
  p.Bimg = Bimg; p.cn2 = cn2; p.cnorm = cnorm; p.gram = gram; p.consts = consts; p.gram_floats = gram_floats;
  p.codes = codes; p.codes_stride = codes_stride;
  p.work_rows = work; p.work_levels = work + n; p.work_count = work_count;
  p.inertia = inertia; p.err_flag = err_flag;
  p.n_tiles = (n + TM - 1) / TM;
  const SmemLayout L = smem_layout(M, K, NT, N1);
  const size_t smem_bytes = (size_t)L.total + 1024;  // slack for the 1024-byte alignment of the dynamic base
  MEVI_CUDA(ctx, cudaFuncSetAttribute(rq_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int grid = (int)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
  rq_tensor_kernel<<<grid, THREADS, smem_bytes, st>>>(p);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, 1);
  MEVI_CUDA(ctx, cudaMemcpyAsync(host_err, err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));

  // exact re-decision of the flagged rows (count stays on the device)
  int rc = mevi_rq_exact_launch(ctx, X, n, d, cb, M, K, metric, codes, codes_stride, nullptr, work, work + n,
                                reinterpret_cast<const int64_t*>(work_count), n, nullptr, st);
  if (rc != MEVI_OK) return rc;
  if (stats) {
    finish_stats_kernel<<<1, 32, 0, st>>>(work_count, n, stats);
    MEVI_COUNT_LAUNCH(ctx, 1);
  }
  if (residual) {
    residual_from_codes_kernel<<<ctx->sm_count * 16, 256, 0, st>>>(X, n, d / 4, cb, M, K, codes, codes_stride, residual);
    MEVI_COUNT_LAUNCH(ctx, 1);
  }
  MEVI_CUDA(ctx, cudaGetLastError());
  return MEVI_OK;
}
