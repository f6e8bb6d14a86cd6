// rq_tensor5.cuh — K1, fifth generation: 128-row tiles, document operand in TENSOR MEMORY, double-buffered
// accumulators, two alternating MMA warps.
//
// What the pipeline trace of v4 showed (tools/rq_trace.py, profiles/r01_trace_v4.txt): inside a tile v4 runs at the
// speed of the tensor pipe, but its single accumulator buffer stops the pipe for the whole epilogue of every tile
// (~25 % of the tile time), and TMEM has no room for a second 256-column accumulator next to the operand stages.
// Halving the tile to 128 rows halves the accumulator (128 columns), so TWO of them fit beside the operand ring and
// the epilogue of tile t runs under the MMAs of tile t+1.  The price is that a codebook chunk now serves 128 rows
// instead of 256 (twice the L2 -> shared-memory codebook traffic; measured cost 3-4 %).
//
//   warp 0       TMA producer: [128 rows x 32 fp32] boxes, 128B-swizzled, 6-stage ring (96 KB in flight)
//   warp 3       codebook producer: 16 KB bulk copies of the pre-swizzled [C_hi|C_lo] chunk image, 6 stages in lockstep
//                with the operand stages
//   warps 4-7    converter group 0 (even chunks), warps 8-11 converter group 1 (odd chunks): ONE THREAD PER ROW
//                (thread = TMEM lane): 8 conflict-free 16-byte loads of the row's chunk, scale/split,
//                2 x tcgen05.st.x16 into the chunk's TMEM operand stage; per-thread partial row norms
//   warps 1, 2   tcgen05.mma with A in TMEM, alternating chunks (warp 1 even, warp 2 odd; an mbarrier hand-shake
//                keeps the issue order): per K step A_hi.C_hi + A_hi.C_lo + A_lo.C_hi into the tile's accumulator,
//                ONE commit per chunk releases the operand + codebook stage.  Two issuing threads because a
//                tcgen05.commit leaves a bubble in the issuing thread's MMA stream (tools/umma_rate.cu)
//   warps 12-15  epilogue of even tile iterations (accumulator 0), warps 16-19 of odd ones (accumulator 1)
// TMEM map (512 columns): [0,128) accumulator 0, [128,256) accumulator 1, [256,448) 6 operand stages x (16 hi + 16 lo).
#pragma once

namespace v5 {

constexpr int TM5 = 128;
constexpr int KC5 = 32;
constexpr int NS5 = 6;                   // depth of the X, operand and codebook rings (they advance in lockstep)
constexpr int X_STAGE5 = TM5 * KC5 * 4;  // 16 KB
constexpr int THREADS5 = 640;
constexpr uint32_t A_COL0_5 = 256;       // first TMEM column of the operand stages
constexpr int GROUP_WARPS = 4;           // warps per converter group / epilogue group

struct Smem5 {
  int x_off, b_off, gram_off, cn2_off, e1_off, lvl_off, stats_off, bar_off, holder_off, total;
};
__host__ __device__ inline Smem5 smem5_layout(int M, int K, int NT) {
  Smem5 L;
  L.x_off = 0;
  L.b_off = L.x_off + NS5 * X_STAGE5;
  L.gram_off = L.b_off + NS5 * (2 * NT) * 64;
  int gram_pad = 0;
  for (int j = 1; j < M; ++j) gram_pad += j * K * (K + 1);
  L.cn2_off = L.gram_off + gram_pad * 4;
  L.e1_off = L.cn2_off + NT * 4;
  L.lvl_off = L.e1_off + NT * 4;
  L.stats_off = L.lvl_off + 64;
  L.bar_off = (L.stats_off + 2 * 2 * TM5 * 4 + 7) & ~7;  // [accumulator][converter group][row] partial norms
  L.holder_off = L.bar_off + 40 * 8;
  L.total = L.holder_off + 16;
  return L;
}

struct Bars5 {
  uint64_t *x_full, *x_empty, *a_full, *a_empty, *b_full, *issued, *acc_full, *acc_empty, *st_full;
};

// Converter group `grp` takes the chunks with (global chunk index & 1) == grp.
template <bool SCALE>
__device__ __forceinline__ void converter_loop5(const Params& p, uint8_t* sX, float* sStats, uint32_t tmem_base, const Bars5& B,
                                                int grp, int gw, int lane, int warp) {
  const int row = gw * 32 + lane;  // row inside the tile == TMEM lane == row of the TMA box
  const float sx = p.consts[C_SX], inv_sx2 = p.consts[C_INV_SX2];
  const int nchunks = p.d / KC5;
  const uint32_t src_row = ptx::smem_u32(sX) + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);  // 128B swizzle: 16-byte unit j of this row sits at j ^ (row & 7)
  const uint32_t t_lane = tmem_base + ((uint32_t)(gw * 32) << 16) + A_COL0_5;
  uint32_t s = 0, ph = 0, it = 0, g = 0, tix = 0;
  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    float norm = 0.f;
    for (int c = 0; c < nchunks; ++c, ++g) {
      if ((int)(g & 1u) == grp) {
        if (!ptx::mbar_wait(&B.x_full[s], ph)) { atomicExch(p.err_flag, 4); return; }
        trace_ev(p, warp, lane, tix, it, c, 0);  // X stage landed
        if (!ptx::mbar_wait(&B.a_empty[s], ph ^ 1)) { atomicExch(p.err_flag, 4); return; }
        trace_ev(p, warp, lane, tix, it, c, 1);  // operand stage free
        ptx::tc_fence_after_sync();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t0, t1, t2, t3;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t0), "=f"(t1), "=f"(t2), "=f"(t3)
                       : "r"(src_row + s * X_STAGE5 + (((uint32_t)j ^ sw) << 4)));
          if (SCALE) { t0 *= sx; t1 *= sx; t2 *= sx; t3 *= sx; }
          norm = fmaf(t0, t0, norm);
          norm = fmaf(t1, t1, norm);
          norm = fmaf(t2, t2, norm);
          norm = fmaf(t3, t3, norm);
          const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
          const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(t0 - b01.x, t1 - b01.y), l23 = __floats2half2_rn(t2 - b23.x, t3 - b23.y);
          hi[2 * j] = *reinterpret_cast<const uint32_t*>(&h01);      // K elements 4j, 4j+1
          hi[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h23);  // K elements 4j+2, 4j+3
          lo[2 * j] = *reinterpret_cast<const uint32_t*>(&l01);
          lo[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }
        ptx::tmem_st16(t_lane + s * 32, hi);
        ptx::tmem_st16(t_lane + s * 32 + 16, lo);
        // the stores consumed every value loaded from the X stage: hand it back to the TMA producer
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.x_empty[s]);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.a_full[s]);
        trace_ev(p, warp, lane, tix, it, c, 2);  // converted and published
      }
      if (++s == NS5) { s = 0; ph ^= 1; }
    }
    // tile finished: publish this group's partial squared row norms (the epilogue adds the two groups)
    const uint32_t buf = it & 1, bph = (it >> 1) & 1;
    if (!ptx::mbar_wait(&B.acc_empty[buf], bph ^ 1)) { atomicExch(p.err_flag, 5); return; }  // stats slot of tile it-2 consumed
    sStats[(buf * 2 + grp) * TM5 + row] = SCALE ? norm * inv_sx2 : norm;
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&B.st_full[buf]);
  }
}

template <int M>
__global__ void __launch_bounds__(THREADS5, 1) rq_tensor5_kernel(Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int K = p.K, NT = p.NT;
  const Smem5 L = smem5_layout(M, K, NT);
  uint8_t* sX = smem + L.x_off;
  uint8_t* sB = smem + L.b_off;
  float* sGram = reinterpret_cast<float*>(smem + L.gram_off);
  float* sCn2 = reinterpret_cast<float*>(smem + L.cn2_off);
  float* sE1 = reinterpret_cast<float*>(smem + L.e1_off);
  float* sLvl = reinterpret_cast<float*>(smem + L.lvl_off);
  float* sStats = reinterpret_cast<float*>(smem + L.stats_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  Bars5 B;
  B.x_full = bars;
  B.x_empty = B.x_full + NS5;
  B.a_full = B.x_empty + NS5;
  B.a_empty = B.a_full + NS5;
  B.b_full = B.a_empty + NS5;
  B.issued = B.b_full + NS5;      // [2] issue-order hand-shake between the two MMA warps
  B.acc_full = B.issued + 2;      // [2]
  B.acc_empty = B.acc_full + 2;   // [2]
  B.st_full = B.acc_empty + 2;    // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L.holder_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b_stage_bytes = (uint32_t)(2 * NT) * 64u;
  const int nchunks = p.d / KC5;

  for (int i = tid; i < p.gram_floats; i += THREADS5) {
    const int r = i / K, c = i - r * K;
    sGram[r * (K + 1) + c] = p.gram[i];
  }
  for (int i = tid; i < NT; i += THREADS5) {
    sCn2[i] = p.cn2[i];
    sE1[i] = p.e1[i];
  }
  for (int i = tid; i < M * 4; i += THREADS5) sLvl[i] = p.lvl[i];
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS5; ++s) {
      ptx::mbar_init(&B.x_full[s], 1);
      ptx::mbar_init(&B.x_empty[s], GROUP_WARPS);
      ptx::mbar_init(&B.a_full[s], GROUP_WARPS);
      ptx::mbar_init(&B.a_empty[s], 1);
      ptx::mbar_init(&B.b_full[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&B.issued[b], 1);
      ptx::mbar_init(&B.acc_full[b], 2);           // one arrival per MMA warp
      ptx::mbar_init(&B.acc_empty[b], GROUP_WARPS);
      ptx::mbar_init(&B.st_full[b], 2 * GROUP_WARPS);
    }
    ptx::mbar_fence_init();
  }
  if (warp == 0 && lane == 0) ptx::tma_prefetch_desc(&tmap);
  if (warp == 2) ptx::tmem_alloc(tmem_holder, TMEM_COLS);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  trace_clock(p, 0);

  // Control warps walk their loops with all 32 lanes (operands stay warp-uniform) and issue from one elected lane:
  // `if (lane == 0)` would wrap every TMA / tcgen05 instruction in an R2UR waterfall loop.
  if (warp == 0) {
    uint32_t s = 0, ph = 0, tix = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c < nchunks; ++c) {
        if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait_backoff(&B.x_empty[s], ph ^ 1, 32))) {
          if (lane == 0) atomicExch(p.err_flag, 1);
          return;
        }
        trace_ev(p, warp, lane, tix, it, c, 0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&B.x_full[s], X_STAGE5);
          ptx::tma_load_2d(sX + (size_t)s * X_STAGE5, &tmap, c * KC5, (int)(tile * TM5), &B.x_full[s]);
        }
        __syncwarp();
        if (++s == NS5) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    uint32_t s = 0, ph = 0, tix = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c < nchunks; ++c) {
        if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait_backoff(&B.a_empty[s], ph ^ 1, 32))) {
          if (lane == 0) atomicExch(p.err_flag, 7);
          return;
        }
        trace_ev(p, warp, lane, tix, it, c, 0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&B.b_full[s], b_stage_bytes);
          ptx::bulk_g2s(sB + (size_t)s * b_stage_bytes, p.Bimg + (size_t)c * (2 * NT) * KC5, b_stage_bytes, &B.b_full[s]);
        }
        __syncwarp();
        if (++s == NS5) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // MMA warp `mw` takes the chunks with (global chunk index & 1) == mw.  Chunk g may only be issued after chunk
    // g-1 (the first MMA of a tile overwrites the accumulator): warp mw arrives on issued[mw] after every chunk,
    // the other warp waits for that arrival before issuing the next one.
    const int mw = warp - 1;
    const uint32_t idesc = ptx::umma_idesc_f16_m128((uint32_t)NT);
    uint32_t s = 0, ph = 0, it = 0, g = 0, tix = 0, n_other = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      const uint32_t d_tmem = tmem_base + buf * 128;
      bool mine_in_tile = false;
      for (int c = 0; c < nchunks; ++c, ++g) {
        if ((int)(g & 1u) == mw) {
          bool ok = true;
          if (!mine_in_tile) {
            // first chunk of mine in this tile: the epilogue of tile it-2 must have drained this accumulator
            ok = ptx::mbar_wait(&B.acc_empty[buf], bph ^ 1);
            mine_in_tile = true;
          }
          ok = ok && ptx::mbar_wait(&B.a_full[s], ph) && ptx::mbar_wait(&B.b_full[s], ph);
          // chunk g-1 (the other warp's) has been issued: its n_other-th arrival on issued[mw ^ 1]
          if (g > 0) ok = ok && ptx::mbar_wait(&B.issued[mw ^ 1], (n_other - 1) & 1);
          if (!__all_sync(MEVI_FULL_MASK, ok)) {
            if (lane == 0) atomicExch(p.err_flag, 3);
            return;
          }
          trace_ev(p, warp, lane, tix, it, c, 1);  // everything this chunk needs is there
          ptx::tc_fence_after_sync();
          const uint32_t b_hi = ptx::smem_u32(sB + (size_t)s * b_stage_bytes);
          const uint32_t b_lo = b_hi + (uint32_t)NT * 64u;
          if (ptx::elect_one()) {
            if (!(p.debug & 2)) {
              const uint32_t a_hi = tmem_base + A_COL0_5 + s * 32, a_lo = a_hi + 16;
#pragma unroll
              for (int ks = 0; ks < KC5 / 16; ++ks) {
                ptx::umma_f16_ts(d_tmem, a_hi + ks * 8, ptx::umma_desc_sw64(b_hi + ks * 32), idesc, (c | ks) != 0 ? 1u : 0u);
                ptx::umma_f16_ts(d_tmem, a_hi + ks * 8, ptx::umma_desc_sw64(b_lo + ks * 32), idesc, 1u);
                ptx::umma_f16_ts(d_tmem, a_lo + ks * 8, ptx::umma_desc_sw64(b_hi + ks * 32), idesc, 1u);
              }
            }
            ptx::mbar_arrive(&B.issued[mw]);
            // ONE commit per chunk: operand (TMEM) and codebook (smem) stage s are released together
            ptx::umma_commit(&B.a_empty[s]);
          }
          __syncwarp();
          trace_ev(p, warp, lane, tix, it, c, 2);  // issued + committed
        } else {
          ++n_other;  // the other warp issues this chunk
        }
        if (++s == NS5) { s = 0; ph ^= 1; }
      }
      // accumulator complete for this warp's share of the tile.  tcgen05.commit tracks the executing thread's
      // own MMAs, so BOTH warps arrive on acc_full (a warp without a chunk in this tile arrives directly).
      if (ptx::elect_one()) {
        if (mine_in_tile) ptx::umma_commit(&B.acc_full[buf]);
        else ptx::mbar_arrive(&B.acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
    const int cw = warp - CONV_WARP0;
    if (p.consts[C_SX] == 1.f)
      converter_loop5<false>(p, sX, sStats, tmem_base, B, cw >> 2, cw & 3, lane, warp);
    else
      converter_loop5<true>(p, sX, sStats, tmem_base, B, cw >> 2, cw & 3, lane, warp);
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;
    const int eg = ew >> 2, q = ew & 3;  // epilogue group (tile parity), 32-lane quarter of TMEM (== warp % 4)
    const float m2inv = (p.metric == MEVI_METRIC_L2 ? -2.f : -1.f) * p.consts[C_INV];
    const bool l2 = p.metric == MEVI_METRIC_L2;
    double inertia_acc = 0.0;
    uint32_t it = 0, tix = 0;
    bool ok = true;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, ++it) {
      if ((int)(it & 1u) != eg) continue;
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      if (!ptx::mbar_wait_backoff(&B.acc_full[buf], bph, 64) || !ptx::mbar_wait_backoff(&B.st_full[buf], bph, 32)) {
        atomicExch(p.err_flag, 6);
        ok = false;
        break;
      }
      ptx::tc_fence_after_sync();
      trace_ev(p, warp, lane, tix, it, 255, 0);  // accumulator ready
      {
        const int rl = q * 32 + lane;
        const float xn2 = sStats[(buf * 2 + 0) * TM5 + rl] + sStats[(buf * 2 + 1) * TM5 + rl];
        const float xn = sqrtf(xn2), nxn = -xn;
        const uint32_t taddr = tmem_base + buf * 128 + ((uint32_t)(q * 32) << 16);
        const int64_t row = tile * TM5 + rl;
        int code[M];
        int flag_level = -1;
        float last_best = 0.f;
#pragma unroll
        for (int j = 0; j < M; ++j) code[j] = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (p.debug & 4) break;
          const float* gj = sGram + (j * (j - 1) / 2) * K * (K + 1);
          const float* grow[M > 1 ? M - 1 : 1];
#pragma unroll
          for (int m = 0; m < j; ++m) grow[m] = gj + (m * K + code[m]) * (K + 1);
          float m1 = CUDART_INF_F, ub = CUDART_INF_F, eb = 0.f, u1 = CUDART_INF_F, u2 = CUDART_INF_F;
          int besti = 0;
          for (int k0 = 0; k0 < K; k0 += 32) {
            uint32_t ra[32];
            ptx::tmem_ld32(taddr + j * K + k0, ra);
            ptx::tmem_ld_wait();
            float dk[32];
            float c1 = CUDART_INF_F;
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) {
              float base = l2 ? sCn2[j * K + k0 + kk] : 0.f;
              float gsum = 0.f;
#pragma unroll
              for (int m = 0; m < j; ++m) gsum += grow[m][k0 + kk];
              base = l2 ? fmaf(2.f, gsum, base) : gsum;
              dk[kk] = fmaf(__uint_as_float(ra[kk]), m2inv, base);
              c1 = fminf(c1, dk[kk]);
              const float u = fmaf(nxn, sE1[j * K + k0 + kk], dk[kk]);
              u2 = fminf(u2, fmaxf(u1, u));
              u1 = fminf(u1, u);
            }
            int ci = 0;
#pragma unroll
            for (int kk = 31; kk >= 0; --kk)
              if (dk[kk] == c1) ci = kk;
            if (c1 < m1) {
              m1 = c1;
              besti = k0 + ci;
              eb = xn * sE1[j * K + besti];
              ub = fmaf(nxn, sE1[j * K + besti], c1);
            }
          }
          code[j] = besti;
          const float other_lo = (ub == u1) ? u2 : u1;
          const bool clear = other_lo > m1 + eb + sLvl[j * 4 + 1];
          if (!clear && flag_level < 0) flag_level = j;
          last_best = m1;
        }
        // every TMEM / stats read of this tile is done: hand the accumulator (and the stats slot) back before the
        // global stores
        ptx::tc_fence_before_sync();
        __syncwarp();
        trace_ev(p, warp, lane, tix, it, 255, 1);  // accumulator drained
        if (lane == 0) ptx::mbar_arrive(&B.acc_empty[buf]);
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
      }
    }
    if (p.inertia) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) inertia_acc += __shfl_xor_sync(MEVI_FULL_MASK, inertia_acc, o);
      if (lane == 0 && inertia_acc != 0.0) atomicAdd(p.inertia, inertia_acc);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  trace_clock(p, 1);
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

inline int make_x_tensormap5(mevi_ctx* ctx, const float* X, int64_t n, int d, CUtensorMap* out) {
  if (!ctx->tmap_encode_fn) {
    CUtensorMap dummy;
    int rc = v3::make_x_tensormap(ctx, X, n, d, &dummy);  // resolves the driver entry point
    if (rc != MEVI_OK) return rc;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)KC5, (cuuint32_t)TM5};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((v3::EncodeTiledFn)ctx->tmap_encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstride, box, estr,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mevi_set_error(ctx, MEVI_ERR_CUDA, "cuTensorMapEncodeTiled (128B swizzle, 128-row box) failed with %d", (int)r);
  return MEVI_OK;
}

}  // namespace v5
