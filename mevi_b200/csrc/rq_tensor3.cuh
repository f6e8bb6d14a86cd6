// rq_tensor3.cuh — K1, third generation of the tcgen05 RQ encode kernel (included by rq_tensor.cu).
//
// Same algorithm, error model and work-list protocol as rq_tensor_kernel (see rq_tensor.cu); what
// changes is how X reaches the converters.  Register-staged LDG prefetch stalls on scoreboard slots
// (measured: 4.5 TB/s with nothing but the loads running), so here the TMA engine does the streaming:
//   warp 0   one thread issues cp.async.bulk.tensor.2d boxes of [256 rows x 32 fp32] (32 KB) into a
//            3-stage shared-memory ring (96 KB in flight per SM, no registers, zero-fill past the end)
//   warp 3   one thread streams the pre-swizzled [C_hi|C_lo] chunk images (16 KB bulk copies)
//   warps 4-11 converters: conflict-free 16-byte shared loads of the fp32 stage -> scale, split into
//            fp16 hi|lo, write the UMMA K-major 64B-swizzle operand tiles (2-stage ring), row norms
//   warp 1   tcgen05.mma per 128-row half and 16-wide K step: A_hi.C_hi + A_hi.C_lo + A_lo.C_hi, all
//            accumulated into ONE 128-column fp32 accumulator per half (2 halves x 2 buffers = 512 TMEM cols)
//   warps 12-15 epilogue as before (one accumulator load per level instead of two)
// Tile = 256 rows, so the codebook image is re-read from L2 once per 256 rows (half the v2 traffic).
#pragma once
#include <cuda.h>

namespace v3 {

constexpr int TM3 = 256;
constexpr int KC3 = 32;
constexpr int NSX = 3, NSA3 = 2, NSB3 = 2;
static_assert(NSA3 == NSB3, "the A and B rings share their 'empty' barriers");
constexpr int X_STAGE = TM3 * KC3 * 4;  // 32 KB
constexpr int A_HALF = 128 * 64;        // one fp16 operand tile: 128 rows x 64 B
constexpr int A_STAGE3 = 4 * A_HALF;    // [half0 hi][half0 lo][half1 hi][half1 lo]
constexpr int THREADS3 = 512;
constexpr int TMEM_BUF3 = 256;

struct Smem3 {
  int x_off, a_off, b_off, gram_off, cn2_off, e1_off, lvl_off, stats_off, bar_off, holder_off, total;
};
__host__ __device__ inline Smem3 smem3_layout(int M, int K, int NT) {
  Smem3 L;
  L.x_off = 0;
  L.a_off = L.x_off + NSX * X_STAGE;
  L.b_off = L.a_off + NSA3 * A_STAGE3;
  L.gram_off = L.b_off + NSB3 * (2 * NT) * 64;
  int gram_pad = 0;
  for (int j = 1; j < M; ++j) gram_pad += j * K * (K + 1);
  L.cn2_off = L.gram_off + gram_pad * 4;
  L.e1_off = L.cn2_off + NT * 4;
  L.lvl_off = L.e1_off + NT * 4;
  L.stats_off = L.lvl_off + 64;
  L.bar_off = (L.stats_off + 2 * TM3 * 4 + 7) & ~7;
  L.holder_off = L.bar_off + 24 * 8;
  L.total = L.holder_off + 16;
  return L;
}

// Bimg32[chunk][row][32 halfs], rows 0..NT-1 = hi(c*sc), NT..2NT-1 = lo; 16-byte units XOR-swizzled by (row>>1)&3
__global__ void bimg32_kernel(const float* __restrict__ cb, int rows_valid, int d, int NT, const float* __restrict__ consts,
                              __half* __restrict__ Bimg) {
  const int units_per_row = d / 8;
  const int total = NT * units_per_row;
  const float sc = consts[C_SC];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / units_per_row, ug = i - r * units_per_row;
    const int chunk = ug / 4, u = ug & 3;
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = r < rows_valid ? cb[(size_t)r * d + ug * 8 + e] * sc : 0.f;
      hi[e] = __float2half_rn(t);
      lo[e] = __float2half_rn(t - __half2float(hi[e]));
    }
    const size_t base = (size_t)chunk * (2 * NT) * KC3;
    const int up = u ^ ((r >> 1) & 3);  // NT % 8 == 0, so row NT + r has the same swizzle phase
    *reinterpret_cast<uint4*>(Bimg + base + (size_t)r * KC3 + up * 8) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(Bimg + base + (size_t)(NT + r) * KC3 + up * 8) = *reinterpret_cast<const uint4*>(lo);
  }
}

template <bool SCALE>
__device__ __forceinline__ void converter_loop3(const Params& p, uint8_t* sX, uint8_t* sA, float* sStats, uint64_t* x_full,
                                                uint64_t* x_empty, uint64_t* a_full, uint64_t* a_empty, uint64_t* acc_empty,
                                                uint64_t* st_full, int cw, int lane) {
  const int rsub = lane >> 3, c4 = lane & 7;
  const float sx = p.consts[C_SX], inv_sx2 = p.consts[C_INV_SX2];
  const int nchunks = p.d / KC3;
  const uint32_t sX_u32 = ptx::smem_u32(sX), sA_u32 = ptx::smem_u32(sA);
  const int row0 = cw * 4 + rsub;  // row of i = 0; i adds 32 rows
  // byte offsets that do not depend on the chunk
  const uint32_t ld_off = (uint32_t)row0 * 128u + (uint32_t)c4 * 16u;
  uint32_t st_off[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + 32 * i, half = row >> 7, rl = row & 127;
    st_off[i] = (uint32_t)half * (2 * A_HALF) + (uint32_t)rl * 64u + ((uint32_t)((c4 >> 1) ^ ((rl >> 1) & 3)) << 4) +
                ((uint32_t)(c4 & 1) << 3);
  }
  float2 norm[8];  // (even, odd) partial sums: packed fp32 FMAs
#pragma unroll
  for (int i = 0; i < 8; ++i) norm[i] = make_float2(0.f, 0.f);
  const float2 sx2 = make_float2(sx, sx);
  uint32_t xs = 0, xph = 0, as = 0, aph = 0, it = 0;
  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    for (int c = 0; c < nchunks; ++c) {
      if (!ptx::mbar_wait(&x_full[xs], xph) || !ptx::mbar_wait(&a_empty[as], aph ^ 1)) { atomicExch(p.err_flag, 4); return; }
      const uint32_t src = sX_u32 + xs * X_STAGE + ld_off;
      const uint32_t dst = sA_u32 + as * A_STAGE3;
      // all eight 16-byte loads first (the volatile loads/stores keep program order, so interleaving
      // them with the stores would expose the shared-memory latency eight times per chunk)
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src + i * 32 * 128));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 p01 = make_float2(v[i].x, v[i].y), p23 = make_float2(v[i].z, v[i].w);
        if (SCALE) { p01 = ptx::f2_mul(p01, sx2); p23 = ptx::f2_mul(p23, sx2); }
        norm[i] = ptx::f2_fma(p01, p01, norm[i]);
        norm[i] = ptx::f2_fma(p23, p23, norm[i]);
        const __half2 h01 = __float22half2_rn(p01), h23 = __float22half2_rn(p23);
        const __half2 l01 = __float22half2_rn(ptx::f2_sub(p01, __half22float2(h01)));
        const __half2 l23 = __float22half2_rn(ptx::f2_sub(p23, __half22float2(h23)));
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + st_off[i]), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                     "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + st_off[i] + A_HALF),
                     "r"(*reinterpret_cast<const uint32_t*>(&l01)), "r"(*reinterpret_cast<const uint32_t*>(&l23)) : "memory");
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&a_full[as]);
        ptx::mbar_arrive(&x_empty[xs]);
      }
      if (++xs == NSX) { xs = 0; xph ^= 1; }
      if (++as == NSA3) { as = 0; aph ^= 1; }
    }
    // tile finished: publish squared row norms for the epilogue
    const uint32_t buf = it & 1, ph = (it >> 1) & 1;
    if (!ptx::mbar_wait(&acc_empty[buf], ph ^ 1)) { atomicExch(p.err_flag, 5); return; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = norm[i].x + norm[i].y;
      s += __shfl_xor_sync(MEVI_FULL_MASK, s, 4);
      s += __shfl_xor_sync(MEVI_FULL_MASK, s, 2);
      s += __shfl_xor_sync(MEVI_FULL_MASK, s, 1);
      if (c4 == 0) sStats[buf * TM3 + row0 + 32 * i] = SCALE ? s * inv_sx2 : s;
      norm[i] = make_float2(0.f, 0.f);
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&st_full[buf]);
  }
}

template <int M>
__global__ void __launch_bounds__(THREADS3, 1) rq_tensor3_kernel(Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int K = p.K, NT = p.NT;
  const Smem3 L = smem3_layout(M, K, NT);
  uint8_t* sX = smem + L.x_off;
  uint8_t* sA = smem + L.a_off;
  uint8_t* sB = smem + L.b_off;
  float* sGram = reinterpret_cast<float*>(smem + L.gram_off);
  float* sCn2 = reinterpret_cast<float*>(smem + L.cn2_off);
  float* sE1 = reinterpret_cast<float*>(smem + L.e1_off);
  float* sLvl = reinterpret_cast<float*>(smem + L.lvl_off);
  float* sStats = reinterpret_cast<float*>(smem + L.stats_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + NSX;
  uint64_t* a_full = x_empty + NSX;
  uint64_t* a_empty = a_full + NSA3;
  uint64_t* b_full = a_empty + NSA3;
  uint64_t* b_empty = b_full + NSB3;
  uint64_t* acc_full = b_empty + NSB3;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* st_full = acc_empty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L.holder_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b_stage_bytes = (uint32_t)(2 * NT) * 64u;
  const int nchunks = p.d / KC3;

  for (int i = tid; i < p.gram_floats; i += THREADS3) {
    const int r = i / K, c = i - r * K;
    sGram[r * (K + 1) + c] = p.gram[i];
  }
  for (int i = tid; i < NT; i += THREADS3) {
    sCn2[i] = p.cn2[i];
    sE1[i] = p.e1[i];
  }
  for (int i = tid; i < M * 4; i += THREADS3) sLvl[i] = p.lvl[i];
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSX; ++s) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], CONV_WARPS); }
    for (int s = 0; s < NSA3; ++s) { ptx::mbar_init(&a_full[s], CONV_WARPS); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NSB3; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 4); ptx::mbar_init(&st_full[b], CONV_WARPS); }
    ptx::mbar_fence_init();
  }
  if (warp == 0 && lane == 0) ptx::tma_prefetch_desc(&tmap);
  if (warp == 2) ptx::tmem_alloc(tmem_holder, TMEM_COLS);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===== X producer: TMA boxes of [256 rows x 32 floats] =====
      uint32_t s = 0, ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          if (!ptx::mbar_wait_backoff(&x_empty[s], ph ^ 1, 32)) { atomicExch(p.err_flag, 1); return; }
          ptx::mbar_arrive_expect_tx(&x_full[s], X_STAGE);
          ptx::tma_load_2d(sX + (size_t)s * X_STAGE, &tmap, c * KC3, (int)(tile * TM3), &x_full[s]);
          if (++s == NSX) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ===== codebook producer =====
      uint32_t s = 0, ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          if (!ptx::mbar_wait_backoff(&a_empty[s], ph ^ 1, 32)) { atomicExch(p.err_flag, 7); return; }
          ptx::mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
          ptx::bulk_g2s(sB + (size_t)s * b_stage_bytes, p.Bimg + (size_t)c * (2 * NT) * KC3, b_stage_bytes, &b_full[s]);
          if (++s == NSB3) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (all operands warp-uniform), one elected lane issues =====
    const uint32_t idesc = ptx::umma_idesc_f16_m128((uint32_t)NT);
    uint32_t as = 0, aph = 0, bs = 0, bph = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait(&acc_empty[buf], ph ^ 1))) {
        if (lane == 0) atomicExch(p.err_flag, 2);
        return;
      }
      ptx::tc_fence_after_sync();
      for (int c = 0; c < nchunks; ++c) {
        if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait(&a_full[as], aph) && ptx::mbar_wait(&b_full[bs], bph))) {
          if (lane == 0) atomicExch(p.err_flag, 3);
          return;
        }
        ptx::tc_fence_after_sync();
        const uint32_t a_base = ptx::smem_u32(sA + (size_t)as * A_STAGE3);
        const uint32_t b_hi = ptx::smem_u32(sB + (size_t)bs * b_stage_bytes);
        const uint32_t b_lo = b_hi + (uint32_t)NT * 64u;
        if (ptx::elect_one()) {
          if (!(p.debug & 2)) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t d_tmem = tmem_base + buf * TMEM_BUF3 + h * 128;
              const uint32_t a_hi = a_base + h * (2 * A_HALF), a_lo = a_hi + A_HALF;
#pragma unroll
              for (int ks = 0; ks < KC3 / 16; ++ks) {
                ptx::umma_f16(d_tmem, ptx::umma_desc_sw64(a_hi + ks * 32), ptx::umma_desc_sw64(b_hi + ks * 32), idesc, (c | ks) != 0 ? 1u : 0u);
                ptx::umma_f16(d_tmem, ptx::umma_desc_sw64(a_hi + ks * 32), ptx::umma_desc_sw64(b_lo + ks * 32), idesc, 1u);
                ptx::umma_f16(d_tmem, ptx::umma_desc_sw64(a_lo + ks * 32), ptx::umma_desc_sw64(b_hi + ks * 32), idesc, 1u);
              }
            }
          }
          // ONE commit per chunk: a tcgen05.commit stalls the MMA stream for ~600 cycles (tools/umma_rate.cu), a
          // second one back to back another ~245.  The A and B rings advance in lockstep, so the converters and
          // the codebook producer both wait on a_empty[as].
          ptx::umma_commit(&a_empty[as]);
          if (c == nchunks - 1) ptx::umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++as == NSA3) { as = 0; aph ^= 1; }
        if (++bs == NSB3) { bs = 0; bph ^= 1; }
      }
    }
  } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
    if (p.consts[C_SX] == 1.f)
      converter_loop3<false>(p, sX, sA, sStats, x_full, x_empty, a_full, a_empty, acc_empty, st_full, warp - CONV_WARP0, lane);
    else
      converter_loop3<true>(p, sX, sA, sStats, x_full, x_empty, a_full, a_empty, acc_empty, st_full, warp - CONV_WARP0, lane);
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue =====
    const int ew = warp - EPI_WARP0;
    const float m2inv = (p.metric == MEVI_METRIC_L2 ? -2.f : -1.f) * p.consts[C_INV];
    const bool l2 = p.metric == MEVI_METRIC_L2;
    double inertia_acc = 0.0;
    uint32_t it = 0;
    bool ok = true;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      if (!ptx::mbar_wait_backoff(&acc_full[buf], ph, 96) || !ptx::mbar_wait_backoff(&st_full[buf], ph, 32)) { atomicExch(p.err_flag, 6); ok = false; break; }
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int rl = h * 128 + ew * 32 + lane;
        const float xn2 = sStats[buf * TM3 + rl];
        const float xn = sqrtf(xn2), nxn = -xn;
        const uint32_t taddr = tmem_base + buf * TMEM_BUF3 + h * 128 + ((uint32_t)(ew * 32) << 16);
        const int64_t row = tile * TM3 + rl;
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
              float g = 0.f;
#pragma unroll
              for (int m = 0; m < j; ++m) g += grow[m][k0 + kk];
              base = l2 ? fmaf(2.f, g, base) : g;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_x_tensormap(mevi_ctx* ctx, const float* X, int64_t n, int d, CUtensorMap* out) {
  if (!ctx->tmap_encode_fn) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return mevi_set_error(ctx, MEVI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    }
    ctx->tmap_encode_fn = fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)KC3, (cuuint32_t)TM3};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)ctx->tmap_encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstride, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mevi_set_error(ctx, MEVI_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return MEVI_OK;
}

}  // namespace v3
