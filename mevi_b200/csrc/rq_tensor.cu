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
#include <stdlib.h>
#include <cstdio>
#include <vector>
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
#ifndef MEVI_NSA
#define MEVI_NSA 3
#endif
#ifndef MEVI_NSB
#define MEVI_NSB 3
#endif
constexpr int NSA = MEVI_NSA, NSB = MEVI_NSB;  // ring depths
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
  const __half* Bimg; const float* cn2; const float* e1; const float* lvl; const float* gram; const float* consts;
  int gram_floats;
  int32_t* codes; int64_t codes_stride;
  int32_t* work_rows; int32_t* work_levels; unsigned long long* work_count;
  double* inertia; int* err_flag;
  int64_t n_tiles;
  int debug;  // MEVI_RQ_DEBUG bit mask for pipeline ablations (timing experiments only; results are wrong)
  unsigned long long* trace;  // MEVI_RQ_TRACE=<file>: per-warp event timestamps of CTA 0 (pipeline debugging), else null
};

// Pipeline trace (debug aid): lane 0 of every warp of CTA 0 appends (clock << 16 | event << 12 | tile_it << 8 | chunk)
// for tile iterations [TRACE_IT0, TRACE_IT1).  tools/rq_trace.py turns the dump into a per-role timeline.
constexpr int TRACE_SLOTS = 512, TRACE_WARPS = 20, TRACE_IT0 = 6, TRACE_IT1 = 9;
// (SM clock, global nanosecond timer) at the start (which = 0) and end (which = 1) of CTA 0: the SM clock the kernel really ran at
__device__ __forceinline__ void trace_clock(const Params& p, int which) {
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.trace[TRACE_SLOTS * TRACE_WARPS + 2 * which] = (unsigned long long)clock64();
    p.trace[TRACE_SLOTS * TRACE_WARPS + 2 * which + 1] = ns;
  }
}
__device__ __forceinline__ void trace_ev(const Params& p, int warp, int lane, uint32_t& idx, uint32_t it, uint32_t c, uint32_t ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && it >= TRACE_IT0 && it < TRACE_IT1 && idx < TRACE_SLOTS) {
    p.trace[warp * TRACE_SLOTS + idx] = ((unsigned long long)clock64() << 16) | (ev << 12) | ((it & 15u) << 8) | (c & 255u);
    ++idx;
  }
}

struct SmemLayout {
  int a_off, b_off, gram_off, cn2_off, cnorm_off, lvl_off, stats_off, bar_off, holder_off, total;
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
  L.lvl_off = L.cnorm_off + NT * 4;
  L.stats_off = L.lvl_off + 4 * 4 * 4;
  L.bar_off = (L.stats_off + 2 * TM * 4 + 7) & ~7;
  L.holder_off = L.bar_off + (2 * NSA + 2 * NSB + 6) * 8;
  L.total = L.holder_off + 16;
  return L;
}

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// register budgets per warpgroup after setmaxnreg (launch allocates 128 x 512 = 65536):
//   control warps 0-3: 32, converter warps 4-11: 184, epilogue warps 12-15: 112  -> 65,536 registers
constexpr int REGS_CTRL = 32, REGS_CONV = 184, REGS_EPI = 112;


// Converter warp: streams this CTA's tiles chunk by chunk.  Lane (half, l16) of warp cw owns rows
// 2*(cw+8q)+half (q = 0..7) and the 16 bytes at float column 4*l16 of every 64-wide chunk: one
// LDG.128 per q covers two full 256-byte row pieces per warp.  Four register buffers rotate so three
// chunks (96 KB per SM) are always in flight while the fourth is being converted.
template <bool SCALE>
__device__ __forceinline__ void converter_loop(const Params& p, uint8_t* sA, float* sStats, uint64_t* a_full,
                                               uint64_t* a_empty, uint64_t* acc_empty, uint64_t* st_full, int cw, int lane) {
  const int half = lane >> 4, l16 = lane & 15;
  const int rl0 = 2 * cw + half;  // row of q = 0; q adds 16 rows (same row & 7 -> same swizzle phase)
  const float sx = p.consts[C_SX], inv_sx2 = p.consts[C_INV_SX2];
  const int nchunks = p.nchunks;
  const int64_t tile_stride = gridDim.x;
  const int64_t qstride = (int64_t)16 * p.d;
  // shared-memory byte offset of this lane's 8 bytes inside an operand tile (q = 0)
  const uint32_t soff = (uint32_t)rl0 * 128u + ((uint32_t)((l16 >> 1) ^ (rl0 & 7)) << 4) + ((uint32_t)(l16 & 1) << 3);
  const uint32_t sA_u32 = ptx::smem_u32(sA);

  // load cursor: the chunk that the NEXT load instruction belongs to
  int64_t l_tile = blockIdx.x;
  int l_c = 0;
  const float* l_ptr = p.X + (l_tile * TM + rl0) * p.d + l16 * 4;
  int l_valid = 0;  // number of q with a row inside the matrix for the cursor's tile (-1: past the end)
  auto set_valid = [&]() {
    const int64_t left = p.n - (l_tile * TM + rl0);  // rows from this lane's first row to the end
    l_valid = l_tile < p.n_tiles ? (left <= 0 ? 0 : (left >= 128 ? 8 : (int)((left + 15) >> 4))) : -1;
  };
  set_valid();
  auto load_one = [&](float4& v, int q) {  // row q of the cursor's chunk
    if (l_valid == 8) {
      v = ld_stream_f4(l_ptr + q * qstride);
    } else {  // ragged last tile / past the end: rows outside the matrix read as zero
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < l_valid) v = ld_stream_f4(l_ptr + q * qstride);
    }
  };
  auto advance_load = [&]() {
    if (l_valid < 0) return;
    if (++l_c == nchunks) {
      l_c = 0;
      l_tile += tile_stride;
      l_ptr = p.X + (l_tile * TM + rl0) * p.d + l16 * 4;
      set_valid();
    } else {
      l_ptr += KC;
    }
  };

  // process cursor
  int64_t p_tile = blockIdx.x;
  int p_c = 0;
  uint32_t p_stage = 0, p_phase = 0, p_it = 0;
  float norm[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) norm[q] = 0.f;
  bool ok = true;
  // Loads are issued as one burst of 8 per buffer: interleaving single reloads with the conversion
  // makes freshly issued loads share a scoreboard slot with the data about to be consumed and
  // serialises every chunk on a full memory latency (measured: 1.7x slower).
  auto load_chunk = [&](float4 (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) load_one(v[q], q);
    advance_load();
  };
  auto process_chunk = [&](const float4 (&v)[8]) {
    if (p_tile >= p.n_tiles || !ok) return;
    if (!ptx::mbar_wait(&a_empty[p_stage], p_phase ^ 1)) { atomicExch(p.err_flag, 4); ok = false; return; }
    const uint32_t st_hi = sA_u32 + p_stage * A_STAGE_BYTES + soff;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float t0 = v[q].x, t1 = v[q].y, t2 = v[q].z, t3 = v[q].w;
      if (p.debug & 8) { norm[q] += t0 + t1 + t2 + t3; continue; }
      if (SCALE) { t0 *= sx; t1 *= sx; t2 *= sx; t3 *= sx; }
      norm[q] = fmaf(t0, t0, norm[q]);
      norm[q] = fmaf(t1, t1, norm[q]);
      norm[q] = fmaf(t2, t2, norm[q]);
      norm[q] = fmaf(t3, t3, norm[q]);
      const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
      const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(t0 - b01.x, t1 - b01.y), l23 = __floats2half2_rn(t2 - b23.x, t3 - b23.y);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_hi + q * 2048), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                   "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_hi + A_TILE_BYTES + q * 2048),
                   "r"(*reinterpret_cast<const uint32_t*>(&l01)), "r"(*reinterpret_cast<const uint32_t*>(&l23)) : "memory");
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&a_full[p_stage]);
    if (++p_stage == NSA) { p_stage = 0; p_phase ^= 1; }
    if (++p_c == nchunks) {
      // tile finished: publish squared row norms for the epilogue
      const uint32_t buf = p_it & 1, ph = (p_it >> 1) & 1;
      if (!ptx::mbar_wait(&acc_empty[buf], ph ^ 1)) { atomicExch(p.err_flag, 5); ok = false; return; }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float s = norm[q];
        s += __shfl_xor_sync(MEVI_FULL_MASK, s, 8);
        s += __shfl_xor_sync(MEVI_FULL_MASK, s, 4);
        s += __shfl_xor_sync(MEVI_FULL_MASK, s, 2);
        s += __shfl_xor_sync(MEVI_FULL_MASK, s, 1);
        if (l16 == 0) sStats[buf * TM + rl0 + 16 * q] = SCALE ? s * inv_sx2 : s;
        norm[q] = 0.f;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&st_full[buf]);
      p_c = 0;
      p_tile += tile_stride;
      ++p_it;
    }
  };
  // four register buffers in rotation: three chunks (96 KB per SM) stay in flight while one is converted
  float4 b0[8], b1[8], b2[8], b3[8];
  load_chunk(b0);
  load_chunk(b1);
  load_chunk(b2);
  while (p_tile < p.n_tiles && ok) {
    load_chunk(b3);
    process_chunk(b0);
    load_chunk(b0);
    process_chunk(b1);
    load_chunk(b1);
    process_chunk(b2);
    load_chunk(b2);
    process_chunk(b3);
  }
}

template <int M>
__global__ void __launch_bounds__(THREADS, 1) rq_tensor_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout L = smem_layout(M, p.K, p.NT, p.N1);
  uint8_t* sA = smem + L.a_off;
  uint8_t* sB = smem + L.b_off;
  float* sGram = reinterpret_cast<float*>(smem + L.gram_off);
  float* sCn2 = reinterpret_cast<float*>(smem + L.cn2_off);
  float* sE1 = reinterpret_cast<float*>(smem + L.cnorm_off);      // [NT] per-candidate error coefficient (x |x|)
  float* sLvl = reinterpret_cast<float*>(smem + L.lvl_off);       // [M][4] per-level margin constants
  float* sStats = reinterpret_cast<float*>(smem + L.stats_off);   // [2][TM] squared row norms
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
  // Gram table: global rows of K floats -> shared rows padded to K+1 (conflict-free per-thread rows)
  for (int i = tid; i < p.gram_floats; i += THREADS) {
    const int r = i / K, c = i - r * K;
    sGram[r * (K + 1) + c] = p.gram[i];
  }
  for (int i = tid; i < NT; i += THREADS) {
    sCn2[i] = p.cn2[i];
    sE1[i] = p.e1[i];
  }
  for (int i = tid; i < M * 4; i += THREADS) sLvl[i] = p.lvl[i];
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

  if (warp < 4) {
    reg_dec<REGS_CTRL>();
    if (warp == 0 && lane == 0) {
      // ===== B producer: one bulk copy per K chunk of the pre-swizzled [C_hi|C_lo] image ==========
      uint32_t g = 0;
      bool ok = true;
      for (int64_t tile = first_tile; tile < p.n_tiles && ok; tile += tile_stride) {
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t s = g % NSB, ph = (g / NSB) & 1;
          if (!ptx::mbar_wait_backoff(&b_empty[s], ph ^ 1, 64)) { atomicExch(p.err_flag, 1); ok = false; break; }
          if ((p.debug & 1) && g >= NSB) { ptx::mbar_arrive(&b_full[s]); continue; }
          ptx::mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
          ptx::bulk_g2s(sB + (size_t)s * b_stage_bytes, p.Bimg + (size_t)c * N1 * KC, b_stage_bytes, &b_full[s]);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===== MMA issuer ===========================================================================
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
          if (!(p.debug & 2)) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              ptx::umma_f16(d_tmem, ptx::umma_desc_sw128(a_hi + ks * 32), ptx::umma_desc_sw128(b_ad + ks * 32), idesc_n1,
                            (c | ks) != 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              ptx::umma_f16(d_tmem, ptx::umma_desc_sw128(a_lo + ks * 32), ptx::umma_desc_sw128(b_ad + ks * 32), idesc_nt, 1u);
          }
          ptx::umma_commit(&a_empty[sa]);
          ptx::umma_commit(&b_empty[sb]);
        }
        if (ok) ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < EPI_WARP0) {
    // ===== converters ===========================================================================
    reg_inc<REGS_CONV>();
    if (p.consts[C_SX] == 1.f)
      converter_loop<false>(p, sA, sStats, a_full, a_empty, acc_empty, st_full, warp - CONV_WARP0, lane);
    else
      converter_loop<true>(p, sA, sStats, a_full, a_empty, acc_empty, st_full, warp - CONV_WARP0, lane);
  } else {
    // ===== epilogue ===============================================================================
    reg_dec<REGS_EPI>();
    const int ew = warp - EPI_WARP0;  // == warp % 4: TMEM lanes 32*ew .. 32*ew+31
    const int rl = ew * 32 + lane;
    const float m2inv = (p.metric == MEVI_METRIC_L2 ? -2.f : -1.f) * p.consts[C_INV];
    const bool l2 = p.metric == MEVI_METRIC_L2;
    double inertia_acc = 0.0;
    uint32_t it = 0;
    bool ok = true;
    for (int64_t tile = first_tile; tile < p.n_tiles && ok; tile += tile_stride, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      if (!ptx::mbar_wait_backoff(&acc_full[buf], ph, 96) || !ptx::mbar_wait_backoff(&st_full[buf], ph, 32)) { atomicExch(p.err_flag, 6); ok = false; break; }
      ptx::tc_fence_after_sync();
      const float xn2 = sStats[buf * TM + rl];
      const float xn = sqrtf(xn2), nxn = -xn;
      const uint32_t taddr = tmem_base + buf * TMEM_BUF_COLS + ((uint32_t)(ew * 32) << 16);
      const int64_t row = tile * TM + rl;
      int code[M];
      int flag_level = -1;
      float last_best = 0.f;
#pragma unroll
      for (int j = 0; j < M; ++j) code[j] = 0;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        if (p.debug & 4) break;
        // Gram block of level j starts after the blocks of levels 1..j-1: sum_{t<j} t*K rows
        const float* gj = sGram + (j * (j - 1) / 2) * K * (K + 1);
        const float* grow[M > 1 ? M - 1 : 1];
#pragma unroll
        for (int m = 0; m < j; ++m) grow[m] = gj + (m * K + code[m]) * (K + 1);
        // best by distance; (u1,u2) = two smallest LOWER bounds u_k = d_k - |x| E1_k over all candidates
        float m1 = CUDART_INF_F, ub = CUDART_INF_F, eb = 0.f, u1 = CUDART_INF_F, u2 = CUDART_INF_F;
        int besti = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
          uint32_t rm[32], rc[32];
          ptx::tmem_ld32(taddr + j * K + k0, rm);
          ptx::tmem_ld32(taddr + NT + j * K + k0, rc);
          ptx::tmem_ld_wait();
          float dk[32];
          float c1 = CUDART_INF_F;
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            float base = l2 ? sCn2[j * K + k0 + kk] : 0.f;
            float g = 0.f;
#pragma unroll
            for (int m = 0; m < j; ++m) g += grow[m][k0 + kk];
            base = l2 ? fmaf(2.f, g, base) : g;
            // dist = |c|^2 - 2 (x.c - g)  (L2)   or   -(x.c - g)  (IP)
            dk[kk] = fmaf(__uint_as_float(rm[kk]) + __uint_as_float(rc[kk]), m2inv, base);
            c1 = fminf(c1, dk[kk]);
            const float u = fmaf(nxn, sE1[j * K + k0 + kk], dk[kk]);
            u2 = fminf(u2, fmaxf(u1, u));
            u1 = fminf(u1, u);
          }
          int ci = 0;
#pragma unroll
          for (int kk = 31; kk >= 0; --kk)
            if (dk[kk] == c1) ci = kk;  // lowest index among equals
          if (c1 < m1) {
            m1 = c1;
            besti = k0 + ci;
            eb = xn * sE1[j * K + besti];
            ub = fmaf(nxn, sE1[j * K + besti], c1);
          }
        }
        code[j] = besti;
        // the best candidate's own lower bound is u1 unless another candidate undercuts it
        const float other_lo = (ub == u1) ? u2 : u1;
        const bool clear = other_lo > m1 + eb + sLvl[j * 4 + 1];  // inf/NaN -> not clear -> exact kernel decides
        if (!clear && flag_level < 0) flag_level = j;
        last_best = m1;
      }
      if (row < p.n) {
        int32_t* dst = p.codes + row * p.codes_stride;
        if (M == 4 && p.codes_stride == 4) {
          *reinterpret_cast<int4*>(dst) = make_int4(code[0], code[M > 1 ? 1 : 0], code[M > 2 ? 2 : 0], code[M > 3 ? 3 : 0]);
        } else {
#pragma unroll
          for (int j = 0; j < M; ++j) dst[j] = code[j];
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


// Error model of the prefilter (all in distance units; f = 2 for L2 where dist = |c|^2 - 2 dot, 1 for IP):
//   |dist_k - exact| <= |x| * E1_k + B_j/2,
//   E1_k = f (U_REL + EPS_A) |c_k| + f Fc      contraction error + fp16 floor of the codebook + fp32 epilogue
//   B_j  = 2 [ f Fx cmax_j + EPS_A (c2max_j + f j gmax_j) ]   fp16 floor of x + roundings of |c|^2 and Gram terms
// A (row, level) is decided by the prefilter only if  d_other - |x| E1_other > d_best + |x| E1_best + B_j
// for every other candidate; otherwise the exact kernel re-decides it.  One block per level.
__global__ void level_consts_kernel(const float* __restrict__ cnorm, const float* __restrict__ cn2,
                                    const float* __restrict__ gram, int M, int K, int metric,
                                    const float* __restrict__ consts, float* __restrict__ e1, float* __restrict__ lvl) {
  const int j = blockIdx.x, lane = threadIdx.x;
  const float f = metric == MEVI_METRIC_L2 ? 2.f : 1.f;
  const float EPS_A = 4.8e-7f;  // 2^-21
  float cmax = 0.f, c2max = 0.f, gmax = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float cn = cnorm[j * K + k];
    cmax = fmaxf(cmax, cn);
    c2max = fmaxf(c2max, cn2[j * K + k]);
    e1[j * K + k] = f * (U_REL + EPS_A) * cn + f * consts[C_FC];
  }
  if (j > 0) {
    const float* gj = gram + (size_t)(j * (j - 1) / 2) * K * K;
    for (int i = lane; i < j * K * K; i += 32) gmax = fmaxf(gmax, fabsf(gj[i]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    cmax = fmaxf(cmax, __shfl_xor_sync(MEVI_FULL_MASK, cmax, o));
    c2max = fmaxf(c2max, __shfl_xor_sync(MEVI_FULL_MASK, c2max, o));
    gmax = fmaxf(gmax, __shfl_xor_sync(MEVI_FULL_MASK, gmax, o));
  }
  if (lane == 0) {
    lvl[j * 4 + 0] = 0.f;
    lvl[j * 4 + 1] = 2.f * (f * consts[C_FX] * cmax + EPS_A * (c2max + f * (float)j * gmax));
    lvl[j * 4 + 2] = cmax;
    lvl[j * 4 + 3] = gmax;
  }
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

#include "rq_tensor3.cuh"
#include "rq_tensor4.cuh"
constexpr int V4_PRE_DEFAULT = 0;  // levels decided after the early TMEM release (0 = off)
#include "rq_tensor5.cuh"

}  // namespace

bool mevi_rq_tensor_supported(mevi_ctx* ctx, int d, int M, int K, int metric) {
  if (!ctx || ctx->cc_major != 10) return false;
  if (d < KC || d % KC != 0 || d > 8192) return false;
  if (M < 1 || M > 4 || K < 32 || K % 32 != 0) return false;
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
               o_cnorm = take(NT * 4), o_e1 = take(NT * 4), o_lvl = take(8 * 4 * 4), o_gram = take((size_t)(gram_floats ? gram_floats : 1) * 4),
               o_bimg = take((size_t)nchunks * N1 * KC * 2 + 1024);
  char* ws = (char*)mevi_ws(ctx, WS_RQ_PREP, off);
  if (!ws) return MEVI_ERR_NOMEM;
  float* consts = (float*)(ws + o_consts);
  unsigned* absmax2 = (unsigned*)(ws + o_abs);
  int* err_flag = (int*)(ws + o_err);
  unsigned long long* work_count = (unsigned long long*)(ws + o_cnt);
  float* cn2 = (float*)(ws + o_cn2);
  float* cnorm = (float*)(ws + o_cnorm);
  float* lvl = (float*)(ws + o_lvl);
  float* e1 = (float*)(ws + o_e1);
  float* gram = (float*)(ws + o_gram);
  __half* Bimg = (__half*)(ws + o_bimg);
  int32_t* work = (int32_t*)mevi_ws(ctx, WS_RQ_WORK, (size_t)n * 8);
  if (!work) return MEVI_ERR_NOMEM;
  int* host_err = (int*)mevi_pinned(ctx, 3, 64);
  if (!host_err) return MEVI_ERR_NOMEM;

  MEVI_CUDA(ctx, cudaMemsetAsync(ws + o_abs, 0, o_cn2 - o_abs, st));  // absmax2, err flag, work count
  absmax_kernel<<<32, 256, 0, st>>>(cb, (int64_t)M * K, d, 1, absmax2);
  const int64_t sample_rows = 2048;
  const int64_t row_step = n > sample_rows ? n / sample_rows : 1;
  absmax_kernel<<<ctx->sm_count, 256, 0, st>>>(X, n, d, row_step, absmax2 + 1);
  consts_kernel<<<1, 32, 0, st>>>(absmax2, d, consts);
  bimg_kernel<<<(NT * (d / 8) + 255) / 256, 256, 0, st>>>(cb, M * K, d, NT, consts, Bimg);
  cnorm_kernel<<<(NT + 7) / 8, 256, 0, st>>>(cb, M * K, d, NT, cn2, cnorm);
  if (gram_floats) gram_kernel<<<(gram_floats + 7) / 8, 256, 0, st>>>(cb, M, K, d, gram, gram_floats);
  level_consts_kernel<<<M, 32, 0, st>>>(cnorm, cn2, gram, M, K, metric, consts, e1, lvl);
  MEVI_CUDA(ctx, cudaGetLastError());
  MEVI_COUNT_LAUNCH(ctx, gram_floats ? 7 : 6);

  Params p;
  p.X = X; p.n = n; p.d = d; p.nchunks = nchunks; p.M = M; p.K = K; p.NT = NT; p.N1 = N1; p.metric = metric;
  p.Bimg = Bimg; p.cn2 = cn2; p.e1 = e1; p.lvl = lvl; p.gram = gram; p.consts = consts; p.gram_floats = gram_floats;
  p.codes = codes; p.codes_stride = codes_stride;
  p.work_rows = work; p.work_levels = work + n; p.work_count = work_count;
  p.inertia = inertia; p.err_flag = err_flag;
  p.n_tiles = (n + TM - 1) / TM;
  {
    const char* dbg = getenv("MEVI_RQ_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.trace = nullptr;
  const char* trace_path = getenv("MEVI_RQ_TRACE");
  if (trace_path && *trace_path) {
    MEVI_CUDA(ctx, cudaMalloc(&p.trace, sizeof(unsigned long long) * (TRACE_SLOTS * TRACE_WARPS + 4)));
    MEVI_CUDA(ctx, cudaMemsetAsync(p.trace, 0, sizeof(unsigned long long) * (TRACE_SLOTS * TRACE_WARPS + 4), st));
  }
  const char* ver = getenv("MEVI_RQ_KERNEL");
  // default: fourth generation (document operand through tensor memory, rq_tensor4.cuh).  MEVI_RQ_KERNEL selects the
  // others for comparison: 5 = 128-row tiles with double-buffered accumulators, 3 = shared-memory operands with a
  // TMA-fed ring, 2 = register-staged loads.  All produce identical codes; DESIGN.md has the measurements.
  const int kver = ver ? atoi(ver) : 4;
  const bool use_v3 = kver == 3;
  if (kver >= 5) {
    // fifth-generation kernel: 128-row tiles, operand through tensor memory, double-buffered accumulators (rq_tensor5.cuh)
    CUtensorMap tmap;
    int trc = v5::make_x_tensormap5(ctx, X, n, d, &tmap);
    if (trc != MEVI_OK) return trc;
    v3::bimg32_kernel<<<(NT * (d / 8) + 255) / 256, 256, 0, st>>>(cb, M * K, d, NT, consts, Bimg);
    MEVI_COUNT_LAUNCH(ctx, 1);
    p.n_tiles = (n + v5::TM5 - 1) / v5::TM5;
    const v5::Smem5 L5 = v5::smem5_layout(M, K, NT);
    const size_t smem5 = (size_t)L5.total + 1024;
    const int grid5 = (int)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
#define MEVI_LAUNCH_RQ_TENSOR5(MM)                                                                                          \
  do {                                                                                                                      \
    MEVI_CUDA(ctx, cudaFuncSetAttribute(v5::rq_tensor5_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5)); \
    v5::rq_tensor5_kernel<MM><<<grid5, v5::THREADS5, smem5, st>>>(p, tmap);                                                 \
  } while (0)
    switch (M) {
      case 1: MEVI_LAUNCH_RQ_TENSOR5(1); break;
      case 2: MEVI_LAUNCH_RQ_TENSOR5(2); break;
      case 3: MEVI_LAUNCH_RQ_TENSOR5(3); break;
      default: MEVI_LAUNCH_RQ_TENSOR5(4); break;
    }
#undef MEVI_LAUNCH_RQ_TENSOR5
  } else if (kver >= 4) {
    // fourth-generation kernel: document operand through tensor memory (rq_tensor4.cuh)
    CUtensorMap tmap;
    int trc = v4::make_x_tensormap4(ctx, X, n, d, &tmap);
    if (trc != MEVI_OK) return trc;
    v3::bimg32_kernel<<<(NT * (d / 8) + 255) / 256, 256, 0, st>>>(cb, M * K, d, NT, consts, Bimg);
    MEVI_COUNT_LAUNCH(ctx, 1);
    p.n_tiles = (n + v4::TM4 - 1) / v4::TM4;
    const v4::Smem4 L4 = v4::smem4_layout(M, K, NT);
    const size_t smem4 = (size_t)L4.total + 1024;
    const int grid4 = (int)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
    // early accumulator release (epilogue decides the last PRE levels from registers): K == 32 only;
    // MEVI_RQ_EARLY=0|2|3|4 overrides the default
    int pre4 = (K == 32) ? V4_PRE_DEFAULT : 0;
    if (const char* e = getenv("MEVI_RQ_EARLY")) pre4 = (K == 32) ? atoi(e) : 0;
#define MEVI_LAUNCH_RQ_TENSOR4_P(MM, PP)                                                                                     \
  do {                                                                                                                      \
    MEVI_CUDA(ctx, cudaFuncSetAttribute(v4::rq_tensor4_kernel<MM, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4)); \
    v4::rq_tensor4_kernel<MM, PP><<<grid4, v4::THREADS4, smem4, st>>>(p, tmap);                                             \
  } while (0)
#define MEVI_LAUNCH_RQ_TENSOR4(MM)                                                                                          \
  do {                                                                                                                      \
    if (pre4 >= 4) MEVI_LAUNCH_RQ_TENSOR4_P(MM, 4);                                                                         \
    else if (pre4 == 3) MEVI_LAUNCH_RQ_TENSOR4_P(MM, 3);                                                                    \
    else if (pre4 >= 1) MEVI_LAUNCH_RQ_TENSOR4_P(MM, 2);                                                                    \
    else MEVI_LAUNCH_RQ_TENSOR4_P(MM, 0);                                                                                   \
  } while (0)
    switch (M) {
      case 1: MEVI_LAUNCH_RQ_TENSOR4(1); break;
      case 2: MEVI_LAUNCH_RQ_TENSOR4(2); break;
      case 3: MEVI_LAUNCH_RQ_TENSOR4(3); break;
      default: MEVI_LAUNCH_RQ_TENSOR4(4); break;
    }
#undef MEVI_LAUNCH_RQ_TENSOR4
#undef MEVI_LAUNCH_RQ_TENSOR4_P
  } else if (use_v3) {
    // third-generation kernel: TMA-fed fp32 ring, 256-row tiles, 64B-swizzle operands (rq_tensor3.cuh)
    CUtensorMap tmap;
    int trc = v3::make_x_tensormap(ctx, X, n, d, &tmap);
    if (trc != MEVI_OK) return trc;
    v3::bimg32_kernel<<<(NT * (d / 8) + 255) / 256, 256, 0, st>>>(cb, M * K, d, NT, consts, Bimg);
    MEVI_COUNT_LAUNCH(ctx, 1);
    p.n_tiles = (n + v3::TM3 - 1) / v3::TM3;
    const v3::Smem3 L3 = v3::smem3_layout(M, K, NT);
    const size_t smem3 = (size_t)L3.total + 1024;
    const int grid3 = (int)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
#define MEVI_LAUNCH_RQ_TENSOR3(MM)                                                                                          \
  do {                                                                                                                      \
    MEVI_CUDA(ctx, cudaFuncSetAttribute(v3::rq_tensor3_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
    v3::rq_tensor3_kernel<MM><<<grid3, v3::THREADS3, smem3, st>>>(p, tmap);                                                 \
  } while (0)
    switch (M) {
      case 1: MEVI_LAUNCH_RQ_TENSOR3(1); break;
      case 2: MEVI_LAUNCH_RQ_TENSOR3(2); break;
      case 3: MEVI_LAUNCH_RQ_TENSOR3(3); break;
      default: MEVI_LAUNCH_RQ_TENSOR3(4); break;
    }
#undef MEVI_LAUNCH_RQ_TENSOR3
  } else {
  const SmemLayout L = smem_layout(M, K, NT, N1);
  const size_t smem_bytes = (size_t)L.total + 1024;  // slack for the 1024-byte alignment of the dynamic base
  const int grid = (int)(p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count);
#define MEVI_LAUNCH_RQ_TENSOR(MM)                                                                                         \
  do {                                                                                                                    \
    MEVI_CUDA(ctx, cudaFuncSetAttribute(rq_tensor_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)); \
    rq_tensor_kernel<MM><<<grid, THREADS, smem_bytes, st>>>(p);                                                           \
  } while (0)
  switch (M) {
    case 1: MEVI_LAUNCH_RQ_TENSOR(1); break;
    case 2: MEVI_LAUNCH_RQ_TENSOR(2); break;
    case 3: MEVI_LAUNCH_RQ_TENSOR(3); break;
    default: MEVI_LAUNCH_RQ_TENSOR(4); break;
  }
#undef MEVI_LAUNCH_RQ_TENSOR
  }
  MEVI_CUDA(ctx, cudaGetLastError());
  if (p.trace) {
    std::vector<unsigned long long> host((size_t)TRACE_SLOTS * TRACE_WARPS + 4);
    MEVI_CUDA(ctx, cudaStreamSynchronize(st));
    MEVI_CUDA(ctx, cudaMemcpy(host.data(), p.trace, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(host.data(), sizeof(unsigned long long), host.size(), f);
      fclose(f);
    }
  }
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
