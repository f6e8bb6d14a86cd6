// rq_tensor6.cuh — K1, generation 6: ONE fp16 MMA per K step (hi.hi) + in-epilogue fp32 refinement.
//
// Why (profiles/r02_k1_hypothesis.md): the split-fp16 kernel (rq_tensor4.cuh) issues three MMAs per K step and, at
// M = 4 (N = 128), runs the SMs at ~1.2-1.3 GHz under the 1 kW power cap; with the two correction MMAs removed and the
// epilogue off the critical path the same pipeline streams at 94 % of the HBM roofline.  So the corrections are not
// computed for every (row, centroid) any more.  The tensor cores contract only the fp16 roundings xh.ch and the
// epilogue bounds what that drops, PER ROW, from the measured norm of the row's remainder:
//     | xh.ch/(sx sc) - x.c_k |  <=  a |c_k| + (|x| + a) cl_k + U_REL |x| |c_k|,   a = |x sx - xh| / sx,  cl_k = |c_k sc - ch_k| / sc
// (Cauchy-Schwarz on the two dropped terms; products of fp16 values are exact in the fp32 accumulator).  A (row, level)
// whose best candidate beats every other candidate's lower bound is decided (95-99 % of them on N(0,1) data with the
// reference-trained codebook); otherwise the warp re-reads the row (3 KB, normally still in L2 — the TMA loaded it a
// tile ago) and computes EXACT-input fp32 dot products with the 2-8 candidates the bound leaves open, (d/128 + 8)
// roundings per element <= U_REL, and decides with the same tight bound the split kernel uses.  Rows still inside that
// bound go to the work list for the fp32 direct-form kernel (rq_exact.cu), exactly as before.
//
// Pipeline (persistent, 1 CTA/SM, 640 threads, tile = 128 rows, K chunk = 32, all rings NS6 = 8 deep and in lockstep):
//   warp 0       TMA producer: [128 rows x 32 fp32] boxes, 128B-swizzled (128 KB in flight per SM)
//   warp 3       codebook producer: NT*64-byte bulk copies of the pre-swizzled C_hi chunk image (8 KB at M*K = 128)
//   warps 4-7    converter group 0 (even chunks), warps 8-11 group 1 (odd chunks), ONE THREAD PER ROW (= TMEM lane):
//                8 conflict-free 16-byte loads of the row's chunk, scale, round to fp16, accumulate |x|^2 and |xl|^2,
//                ONE tcgen05.st.x16 into the chunk's TMEM operand stage
//   warp 1       tcgen05.mma, A from TMEM: two M=128 x N=NT x K=16 MMAs per chunk into the tile's accumulator
//   warps 12-15  epilogue of even tiles (accumulator 0), warps 16-19 of odd tiles (accumulator 1): the epilogue of tile
//                t (argmin per level, bounds, refinement) runs under the MMAs of tiles t+1 and t+2
// TMEM map (512 columns): [0,128) accumulator 0, [128,256) accumulator 1, [256,384) 8 operand stages x 16 columns.
#pragma once

namespace v6 {

constexpr int TM6 = 128;
constexpr int NS6 = 8;
constexpr int X_STAGE6 = TM6 * KC32 * 4;  // 16 KB
constexpr int THREADS6 = 640;
constexpr uint32_t A_COL0_6 = 256;
constexpr int GROUP_WARPS6 = 4;
constexpr int MAX_REFINE = 8;  // more open candidates than this: the exact kernel decides the row (degenerate data)

struct Smem6 {
  int x_off, b_off, gram_off, cn2_off, e1_off, ea1_off, ea2_off, lvl_off, stats_off, bar_off, holder_off, total;
};
__host__ __device__ inline Smem6 smem6_layout(int M, int NT) {
  Smem6 L;
  L.x_off = 0;
  L.b_off = L.x_off + NS6 * X_STAGE6;
  L.gram_off = L.b_off + NS6 * NT * 64;
  int gram_pad = 0;
  for (int j = 1; j < M; ++j) gram_pad += j * 32 * 33;
  L.cn2_off = L.gram_off + gram_pad * 4;
  L.e1_off = L.cn2_off + NT * 4;
  L.ea1_off = L.e1_off + NT * 4;
  L.ea2_off = L.ea1_off + NT * 4;
  L.lvl_off = L.ea2_off + NT * 4;
  L.stats_off = L.lvl_off + 64;
  L.bar_off = (L.stats_off + 2 * 2 * 2 * TM6 * 4 + 7) & ~7;  // [accumulator][converter group][|x|^2, |xl|^2][row]
  L.holder_off = L.bar_off + 48 * 8;
  L.total = L.holder_off + 16;
  return L;
}

struct Bars6 {
  uint64_t *x_full, *x_empty, *a_full, *a_empty, *b_full, *acc_full, *acc_empty, *st_full;
};

// Converter group `grp` takes the chunks with (global chunk index & 1) == grp.
template <bool SCALE>
__device__ __forceinline__ void converter_loop6(const Params& p, uint8_t* sX, float* sStats, uint32_t tmem_base, const Bars6& B,
                                                int grp, int gw, int lane, int warp) {
  const int row = gw * 32 + lane;  // row inside the tile == TMEM lane == row of the TMA box
  const float sx = p.consts[C_SX], inv_sx2 = p.consts[C_INV_SX2];
  const float2 sx2 = make_float2(sx, sx);
  const int nchunks = p.d / KC32;
  const uint32_t src_row = ptx::smem_u32(sX) + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);  // 128B swizzle: 16-byte unit j of this row sits at j ^ (row & 7)
  const uint32_t t_lane = tmem_base + ((uint32_t)(gw * 32) << 16) + A_COL0_6;
  uint32_t s = 0, ph = 0, it = 0, g = 0, tix = 0;
  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    float2 norm2 = make_float2(0.f, 0.f), lon2 = make_float2(0.f, 0.f);
    for (int c = 0; c < nchunks; ++c, ++g) {
      if ((int)(g & 1u) == grp) {
        if (!ptx::mbar_wait(&B.x_full[s], ph)) { atomicExch(p.err_flag, 4); return; }
        trace_ev(p, warp, lane, tix, it, c, 0);  // X stage landed
        if (!ptx::mbar_wait(&B.a_empty[s], ph ^ 1)) { atomicExch(p.err_flag, 4); return; }
        trace_ev(p, warp, lane, tix, it, c, 1);  // operand stage free
        ptx::tc_fence_after_sync();
        uint32_t hi[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float2 p01, p23;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p01.x), "=f"(p01.y), "=f"(p23.x), "=f"(p23.y)
                       : "r"(src_row + s * X_STAGE6 + (((uint32_t)j ^ sw) << 4)));
          if (SCALE) { p01 = ptx::f2_mul(p01, sx2); p23 = ptx::f2_mul(p23, sx2); }
          norm2 = ptx::f2_fma(p01, p01, norm2);
          norm2 = ptx::f2_fma(p23, p23, norm2);
          const __half2 h01 = __float22half2_rn(p01), h23 = __float22half2_rn(p23);
          // what the fp16 operand drops (exact in fp32); its norm is the row's share of the prefilter's error bound
          const float2 l01 = ptx::f2_sub(p01, __half22float2(h01)), l23 = ptx::f2_sub(p23, __half22float2(h23));
          lon2 = ptx::f2_fma(l01, l01, lon2);
          lon2 = ptx::f2_fma(l23, l23, lon2);
          hi[2 * j] = *reinterpret_cast<const uint32_t*>(&h01);      // K elements 4j, 4j+1
          hi[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h23);  // K elements 4j+2, 4j+3
        }
        ptx::tmem_st16(t_lane + s * 16, hi);
        // the store consumed every value loaded from the X stage: hand it back to the TMA producer
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.x_empty[s]);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.a_full[s]);
        trace_ev(p, warp, lane, tix, it, c, 2);  // converted and published
      }
      if (++s == NS6) { s = 0; ph ^= 1; }
    }
    // tile finished: publish this group's partial sums (the epilogue adds the two groups)
    const uint32_t buf = it & 1, bph = (it >> 1) & 1;
    if (!ptx::mbar_wait(&B.acc_empty[buf], bph ^ 1)) { atomicExch(p.err_flag, 5); return; }  // stats slot of tile it-2 consumed
    float* st = sStats + (buf * 2 + grp) * 2 * TM6;
    st[row] = SCALE ? (norm2.x + norm2.y) * inv_sx2 : norm2.x + norm2.y;
    st[TM6 + row] = SCALE ? (lon2.x + lon2.y) * inv_sx2 : lon2.x + lon2.y;
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&B.st_full[buf]);
  }
}

// Phase 2 of the epilogue, out of line (it must not cost the common path registers): the rows of this warp that phase 1
// left open, one at a time, warp-cooperatively.  For the row of lane `src`, from its first open level on: recompute
// the level's distances from tensor memory under the row's CURRENT codes (all lanes issue the collective tcgen05.ld,
// lane src uses its own row), list the candidates the hi.hi bound cannot exclude, compute their exact-input fp32
// distances with all 32 lanes (row and centroids come through L2; phase 1 prefetched the row), decide with the tight
// bound.  A level whose decision changes forces the later levels to be recomputed, otherwise only the levels phase 1
// marked open are visited.  codes are packed one byte per level.
// Returns, for every lane, (packed codes, first level the exact kernel must re-decide or -1, bits of the last level's
// best distance, number of (row, level) refinements made for this lane's row).
template <int M>
__device__ __noinline__ int4 refine_rows6(const float* __restrict__ X, const float* __restrict__ cb, int d, int64_t row0, int lane,
                                          uint32_t taddr, bool l2, float m2inv, const float* sGram, const float* sCn2,
                                          const float* sE1, const float* sEA1, const float* sEA2, const float* sLvl, float xn,
                                          float na, unsigned openmask, unsigned packed, float last_best) {
  constexpr int K = 32;
  const float nxn = -xn, mf = l2 ? -2.f : -1.f;
  int flag_level = -1;
  int n_ref = 0;
  unsigned pend = __ballot_sync(MEVI_FULL_MASK, openmask != 0u);
  while (pend != 0u) {
    const int src = __ffs(pend) - 1;
    pend &= pend - 1;
    const bool my = lane == src;
    const unsigned omask = __shfl_sync(MEVI_FULL_MASK, openmask, src);
    const float* xr = X + (row0 + src) * d;
    bool changed = false;
    for (int j = __ffs(omask) - 1; j < M; ++j) {
      if (!changed && !((omask >> j) & 1u)) continue;
      // Gram rows of the codes chosen so far (lane src's; other lanes compute on their own rows and discard)
      const float* gj = sGram + (j * (j - 1) / 2) * K * (K + 1);
      const float* g0 = j > 0 ? gj + (0 * K + (int)(packed & 255u)) * (K + 1) : nullptr;
      const float* g1 = j > 1 ? gj + (1 * K + (int)((packed >> 8) & 255u)) * (K + 1) : nullptr;
      const float* g2 = j > 2 ? gj + (2 * K + (int)((packed >> 16) & 255u)) * (K + 1) : nullptr;
      auto base_of = [&](int k) {
        float g = 0.f;
        if (g0) g += g0[k];
        if (g1) g += g1[k];
        if (g2) g += g2[k];
        return l2 ? fmaf(2.f, g, sCn2[j * K + k]) : g;
      };
      unsigned cand = 0;
      int best = 0;
      float bestd = CUDART_INF_F;
      bool resolved = false;
      {
        uint32_t ra[32];
        ptx::tmem_ld32(taddr + j * K, ra);
        ptx::tmem_ld_wait();
        float u1 = CUDART_INF_F, u2 = CUDART_INF_F;
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
          const float dk = fmaf(__uint_as_float(ra[kk]), m2inv, base_of(kk));
          if (dk < bestd) { bestd = dk; best = kk; }
          const float u = fmaf(nxn, sEA2[j * K + kk], fmaf(na, sEA1[j * K + kk], dk));
          u2 = fminf(u2, fmaxf(u1, u));
          u1 = fminf(u1, u);
        }
        const float eb = fmaf(xn, sEA2[j * K + best], -na * sEA1[j * K + best]);
        const float ub = fmaf(nxn, sEA2[j * K + best], fmaf(na, sEA1[j * K + best], bestd));
        const float hi_best = bestd + eb + sLvl[j * 4 + 1];
        resolved = ((ub == u1) ? u2 : u1) > hi_best;
        if (!resolved) {
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const float dk = fmaf(__uint_as_float(ra[kk]), m2inv, base_of(kk));
            const float u = fmaf(nxn, sEA2[j * K + kk], fmaf(na, sEA1[j * K + kk], dk));
            if (u <= hi_best) cand |= 1u << kk;
          }
        }
      }
      unsigned cm = __shfl_sync(MEVI_FULL_MASK, resolved ? 0u : cand, src);
      const int ncand = __popc(cm);
      if (ncand >= 2 && ncand <= MAX_REFINE) {
        // refined distances of the open candidates, ascending index; (v1,v2) = two smallest tight lower bounds
        float rb = CUDART_INF_F, vb = CUDART_INF_F, v1 = CUDART_INF_F, v2 = CUDART_INF_F;
        int rbi = 0;
        while (cm != 0u) {
          // two candidates per round: the row and both centroid rows are requested in ONE batch of independent loads
          // (18 x 16 B per lane), so a round exposes one memory latency, not one per 128 columns
          const int ka = __ffs(cm) - 1;
          cm &= cm - 1;
          const bool two = cm != 0u;
          const int kb = two ? __ffs(cm) - 1 : ka;
          cm &= cm - 1;  // (0 & ... stays 0)
          const float* ca = cb + (size_t)(j * K + ka) * d;
          const float* cbp = cb + (size_t)(j * K + kb) * d;
          float4 aa = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int b0 = lane * 4; b0 < d; b0 += 768) {
            float4 xv[6], va[6], vb4[6];
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 6; ++t) {
              const int idx = b0 + t * 128;
              const bool in = idx < d;
              xv[t] = in ? __ldg(reinterpret_cast<const float4*>(xr + idx)) : z;
              va[t] = in ? __ldg(reinterpret_cast<const float4*>(ca + idx)) : z;
              vb4[t] = (in && two) ? __ldg(reinterpret_cast<const float4*>(cbp + idx)) : z;
            }
#pragma unroll
            for (int t = 0; t < 6; ++t) {
              aa.x = fmaf(xv[t].x, va[t].x, aa.x);
              aa.y = fmaf(xv[t].y, va[t].y, aa.y);
              aa.z = fmaf(xv[t].z, va[t].z, aa.z);
              aa.w = fmaf(xv[t].w, va[t].w, aa.w);
              ab.x = fmaf(xv[t].x, vb4[t].x, ab.x);
              ab.y = fmaf(xv[t].y, vb4[t].y, ab.y);
              ab.z = fmaf(xv[t].z, vb4[t].z, ab.z);
              ab.w = fmaf(xv[t].w, vb4[t].w, ab.w);
            }
          }
          float sa = (aa.x + aa.y) + (aa.z + aa.w), sb = (ab.x + ab.y) + (ab.z + ab.w);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(MEVI_FULL_MASK, sa, o);
            sb += __shfl_xor_sync(MEVI_FULL_MASK, sb, o);
          }
          auto take = [&](int k, float sdot) {
            const float dr = fmaf(sdot, mf, base_of(k));
            const float v = fmaf(nxn, sE1[j * K + k], dr);
            if (dr < rb) { rb = dr; rbi = k; vb = v; }  // ascending k, strict: lowest index among equals
            v2 = fminf(v2, fmaxf(v1, v));
            v1 = fminf(v1, v);
          };
          take(ka, sa);
          if (two) take(kb, sb);
        }
        resolved = ((vb == v1) ? v2 : v1) > rb + xn * sE1[j * K + rbi] + sLvl[j * 4 + 1];
        best = rbi;
        bestd = rb;
        if (my) ++n_ref;
      }
      if (my) {
        if (!resolved && flag_level < 0) flag_level = j;
        if (best != (int)((packed >> (8 * j)) & 255u)) changed = true;
        packed = (packed & ~(255u << (8 * j))) | ((unsigned)best << (8 * j));
        if (j == M - 1) last_best = bestd;
      }
      changed = __shfl_sync(MEVI_FULL_MASK, changed ? 1 : 0, src) != 0;
      if (__shfl_sync(MEVI_FULL_MASK, flag_level, src) >= 0) break;  // the exact kernel re-decides from there on anyway
    }
  }
  return make_int4((int)packed, flag_level, __float_as_int(last_best), n_ref);
}

template <int M>
__global__ void __launch_bounds__(THREADS6, 1) rq_tensor6_kernel(Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int K = 32;
  const int NT = p.NT;
  const Smem6 L = smem6_layout(M, NT);
  uint8_t* sX = smem + L.x_off;
  uint8_t* sB = smem + L.b_off;
  float* sGram = reinterpret_cast<float*>(smem + L.gram_off);
  float* sCn2 = reinterpret_cast<float*>(smem + L.cn2_off);
  float* sE1 = reinterpret_cast<float*>(smem + L.e1_off);
  float* sEA1 = reinterpret_cast<float*>(smem + L.ea1_off);
  float* sEA2 = reinterpret_cast<float*>(smem + L.ea2_off);
  float* sLvl = reinterpret_cast<float*>(smem + L.lvl_off);
  float* sStats = reinterpret_cast<float*>(smem + L.stats_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  Bars6 B;
  B.x_full = bars;
  B.x_empty = B.x_full + NS6;
  B.a_full = B.x_empty + NS6;
  B.a_empty = B.a_full + NS6;
  B.b_full = B.a_empty + NS6;
  B.acc_full = B.b_full + NS6;    // [2]
  B.acc_empty = B.acc_full + 2;   // [2]
  B.st_full = B.acc_empty + 2;    // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L.holder_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b_stage_bytes = (uint32_t)NT * 64u;
  const int nchunks = p.d / KC32;

  for (int i = tid; i < p.gram_floats; i += THREADS6) {
    const int r = i / K, c = i - r * K;
    sGram[r * (K + 1) + c] = p.gram[i];
  }
  for (int i = tid; i < NT; i += THREADS6) {
    sCn2[i] = p.cn2[i];
    sE1[i] = p.e1[i];
    sEA1[i] = p.ea1[i];
    sEA2[i] = p.ea2[i];
  }
  for (int i = tid; i < M * 4; i += THREADS6) sLvl[i] = p.lvl[i];
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS6; ++s) {
      ptx::mbar_init(&B.x_full[s], 1);
      ptx::mbar_init(&B.x_empty[s], GROUP_WARPS6);
      ptx::mbar_init(&B.a_full[s], GROUP_WARPS6);
      ptx::mbar_init(&B.a_empty[s], 1);
      ptx::mbar_init(&B.b_full[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&B.acc_full[b], 1);
      ptx::mbar_init(&B.acc_empty[b], GROUP_WARPS6);
      ptx::mbar_init(&B.st_full[b], 2 * GROUP_WARPS6);
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
          ptx::mbar_arrive_expect_tx(&B.x_full[s], X_STAGE6);
          ptx::tma_load_2d(sX + (size_t)s * X_STAGE6, &tmap, c * KC32, (int)(tile * TM6), &B.x_full[s]);
        }
        __syncwarp();
        if (++s == NS6) { s = 0; ph ^= 1; }
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
          // the hi rows of chunk c are the first NT*64 bytes of its [C_hi | C_lo] block
          ptx::mbar_arrive_expect_tx(&B.b_full[s], b_stage_bytes);
          ptx::bulk_g2s(sB + (size_t)s * b_stage_bytes, p.Bimg + (size_t)c * (2 * NT) * KC32, b_stage_bytes, &B.b_full[s]);
        }
        __syncwarp();
        if (++s == NS6) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::umma_idesc_f16_m128((uint32_t)NT);
    uint32_t s = 0, ph = 0, it = 0, tix = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      const uint32_t d_tmem = tmem_base + buf * 128;
      // the epilogue of tile it-2 must have drained this accumulator
      if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait(&B.acc_empty[buf], bph ^ 1))) {
        if (lane == 0) atomicExch(p.err_flag, 2);
        return;
      }
      ptx::tc_fence_after_sync();
      trace_ev(p, warp, lane, tix, it, 255, 3);  // accumulator free
      for (int c = 0; c < nchunks; ++c) {
        if (!__all_sync(MEVI_FULL_MASK, ptx::mbar_wait(&B.a_full[s], ph) && ptx::mbar_wait(&B.b_full[s], ph))) {
          if (lane == 0) atomicExch(p.err_flag, 3);
          return;
        }
        trace_ev(p, warp, lane, tix, it, c, 1);  // operand + codebook stage full
        ptx::tc_fence_after_sync();
        const uint32_t b_hi = ptx::smem_u32(sB + (size_t)s * b_stage_bytes);
        if (ptx::elect_one()) {
          if (!(p.debug & 2)) {
            const uint32_t a_hi = tmem_base + A_COL0_6 + s * 16;
#pragma unroll
            for (int ks = 0; ks < KC32 / 16; ++ks)
              ptx::umma_f16_ts(d_tmem, a_hi + ks * 8, ptx::umma_desc_sw64(b_hi + ks * 32), idesc, (c | ks) != 0 ? 1u : 0u);
          }
          // ONE commit per chunk releases the operand (TMEM) and the codebook (smem) stage
          ptx::umma_commit(&B.a_empty[s]);
          if (c == nchunks - 1) ptx::umma_commit(&B.acc_full[buf]);
        }
        __syncwarp();
        trace_ev(p, warp, lane, tix, it, c, 2);  // issued + committed
        if (++s == NS6) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
    const int cw = warp - CONV_WARP0;
    if (p.consts[C_SX] == 1.f)
      converter_loop6<false>(p, sX, sStats, tmem_base, B, cw >> 2, cw & 3, lane, warp);
    else
      converter_loop6<true>(p, sX, sStats, tmem_base, B, cw >> 2, cw & 3, lane, warp);
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;
    const int eg = ew >> 2, q = ew & 3;  // epilogue group (tile parity), 32-lane quarter of TMEM (== warp % 4)
    const bool l2 = p.metric == MEVI_METRIC_L2;
    const float mf = l2 ? -2.f : -1.f;
    const float m2inv = mf * p.consts[C_INV];
    const int d = p.d;
    double inertia_acc = 0.0;
    unsigned n_refined = 0;
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
      const int rl = q * 32 + lane;
      const float* st0 = sStats + (buf * 2 + 0) * 2 * TM6;
      const float* st1 = sStats + (buf * 2 + 1) * 2 * TM6;
      const float xn2 = st0[rl] + st1[rl];
      const float xn = sqrtf(xn2), nxn = -xn;
      const float na = -sqrtf(st0[TM6 + rl] + st1[TM6 + rl]);  // -a: norm of what the fp16 operand dropped
      const uint32_t taddr = tmem_base + buf * 128 + ((uint32_t)(q * 32) << 16);
      const int64_t row0 = tile * TM6 + q * 32;  // first row of this warp
      const int64_t row = row0 + lane;
      const bool valid = row < p.n;
      int code[M];
      int flag_level = -1;
      float last_best = 0.f;
      unsigned openmask = 0;  // levels the hi.hi bound could not decide (phase 1)
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
        uint32_t ra[32];
        ptx::tmem_ld32(taddr + j * K, ra);
        ptx::tmem_ld_wait();
        // c1 = best distance (lowest index among equals); (u1,u2) = two smallest LOWER bounds
        // u_k = d_k - a EA1_k - |x| EA2_k over all candidates
        float c1 = CUDART_INF_F, u1 = CUDART_INF_F, u2 = CUDART_INF_F;
        int best = 0;
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
          float base = l2 ? sCn2[j * K + kk] : 0.f;
          float g = 0.f;
#pragma unroll
          for (int m = 0; m < j; ++m) g += grow[m][kk];
          base = l2 ? fmaf(2.f, g, base) : g;
          // dist = |c|^2 - 2 (x.c - g)  (L2)   or   -(x.c - g)  (IP), g = sum of the Gram rows of the codes chosen so far
          const float dk = fmaf(__uint_as_float(ra[kk]), m2inv, base);
          if (dk < c1) { c1 = dk; best = kk; }
          const float u = fmaf(nxn, sEA2[j * K + kk], fmaf(na, sEA1[j * K + kk], dk));
          u2 = fminf(u2, fmaxf(u1, u));
          u1 = fminf(u1, u);
        }
        // the best's own lower bound, its upper bound + level slack, and the smallest lower bound among the others
        const float eb = fmaf(xn, sEA2[j * K + best], -na * sEA1[j * K + best]);
        const float ub = fmaf(nxn, sEA2[j * K + best], fmaf(na, sEA1[j * K + best], c1));
        const float hi_best = c1 + eb + sLvl[j * 4 + 1];
        const float other_lo = (ub == u1) ? u2 : u1;
        const bool open = valid && !(other_lo > hi_best);  // inf / NaN -> open
        code[j] = best;  // provisional where open: phase 2 re-decides
        last_best = c1;
        // rows that just became open are pulled towards L2 now (24 lines of 128 B), phase 2 reads them a few us later
        unsigned fresh = __ballot_sync(MEVI_FULL_MASK, open && openmask == 0u);
        if (open) openmask |= 1u << j;
        while (fresh != 0u) {
          const int src = __ffs(fresh) - 1;
          fresh &= fresh - 1;
          if (lane * 32 < d) ptx::prefetch_l2(p.X + (row0 + src) * d + lane * 32);
        }
      }
      if (p.open_rows != nullptr) {
        // two-kernel form: rows with an open level leave the pipeline here - their state and the tensor-core
        // accumulators of every level from the first open one on go to the open list (rq_refine6_kernel finishes them with
        // thousands of rows in flight instead of one per epilogue warp); the codes written below are provisional for them
        const unsigned om = __ballot_sync(MEVI_FULL_MASK, openmask != 0u);
        if (om != 0u) {
          const int leader = __ffs(om) - 1;
          unsigned long long base = 0;
          if (lane == leader) base = atomicAdd(p.open_count, (unsigned long long)__popc(om));
          base = __shfl_sync(MEVI_FULL_MASK, base, leader);
          const long long slot = (long long)base + __popc(om & ((1u << lane) - 1u));
          const bool dump = openmask != 0u && slot < p.open_cap;
          const int j0 = openmask != 0u ? __ffs(openmask) - 1 : M;
#pragma unroll
          for (int j = 0; j < M; ++j) {
            if (__ballot_sync(MEVI_FULL_MASK, dump && j >= j0) == 0u) continue;
            uint32_t ra[32];
            ptx::tmem_ld32(taddr + j * K, ra);
            ptx::tmem_ld_wait();
            if (dump && j >= j0) {
              uint4* dst = reinterpret_cast<uint4*>(p.open_t1 + (slot * M + j) * 32);
#pragma unroll
              for (int t = 0; t < 8; ++t) dst[t] = make_uint4(ra[4 * t], ra[4 * t + 1], ra[4 * t + 2], ra[4 * t + 3]);
            }
          }
          if (dump) {
            unsigned packed = 0;
#pragma unroll
            for (int j = 0; j < M; ++j) packed |= (unsigned)code[j] << (8 * j);
            p.open_rows[slot] = (int32_t)row;
            p.open_meta[slot] = make_int4((int)packed, j0, __float_as_int(xn), __float_as_int(na));
          } else if (openmask != 0u) {
            flag_level = j0;  // the open list is full: the exact kernel takes the row from its first open level
          }
        }
      } else if (__ballot_sync(MEVI_FULL_MASK, openmask != 0u) != 0u && !(p.debug & 64)) {
        unsigned packed = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) packed |= (unsigned)code[j] << (8 * j);
        const int4 r = refine_rows6<M>(p.X, p.cb, d, row0, lane, taddr, l2, m2inv, sGram, sCn2, sE1, sEA1, sEA2, sLvl, xn, na,
                                       openmask, packed, last_best);
#pragma unroll
        for (int j = 0; j < M; ++j) code[j] = (r.x >> (8 * j)) & 255;
        flag_level = r.y;
        last_best = __int_as_float(r.z);
        n_refined += (unsigned)r.w;
      } else if (openmask != 0u) {
        flag_level = __ffs(openmask) - 1;  // debug 64: no refinement, the exact kernel takes every open row
      }
      // every TMEM / stats read of this tile is done: hand the accumulator (and the stats slot) back before the
      // global stores
      ptx::tc_fence_before_sync();
      __syncwarp();
      trace_ev(p, warp, lane, tix, it, 255, 1);  // accumulator drained
      if (lane == 0) ptx::mbar_arrive(&B.acc_empty[buf]);
      if (valid) {
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_refined += __shfl_xor_sync(MEVI_FULL_MASK, n_refined, o);
    if (lane == 0 && n_refined != 0u) atomicAdd(p.refine_count, (unsigned long long)n_refined);
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

// Second kernel of the two-kernel form: one warp per open row, lane = candidate.  The row's tensor-core accumulators of
// the levels from its first open one on come from the dump (128 B per level, coalesced), the row itself is read once
// (3 KB, coalesced) and kept in registers; per level the loose bound decides or lists the open candidates, whose exact-input
// fp32 dot products are computed by the whole warp; the tight bound decides or sends the row to the exact kernel's work list.
// Thousands of rows are in flight per SM-wave, so the memory latency that stalls the in-epilogue refinement is hidden.
template <int M>
__global__ void __launch_bounds__(256) rq_refine6_kernel(Params p) {
  constexpr int K = 32;
  extern __shared__ __align__(16) float sm6[];
  float* sGram = sm6;                       // padded rows of K+1
  int gram_pad = 0;
  for (int j = 1; j < M; ++j) gram_pad += j * K * (K + 1);
  float* sCn2 = sGram + gram_pad;
  float* sE1 = sCn2 + M * K;
  float* sEA1 = sE1 + M * K;
  float* sEA2 = sEA1 + M * K;
  float* sLvl = sEA2 + M * K;
  for (int i = threadIdx.x; i < p.gram_floats; i += blockDim.x) sGram[(i / K) * (K + 1) + (i % K)] = p.gram[i];
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    sCn2[i] = p.cn2[i];
    sE1[i] = p.e1[i];
    sEA1[i] = p.ea1[i];
    sEA2[i] = p.ea2[i];
  }
  for (int i = threadIdx.x; i < M * 4; i += blockDim.x) sLvl[i] = p.lvl[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long n_open = (long long)(*p.open_count < (unsigned long long)p.open_cap ? *p.open_count : (unsigned long long)p.open_cap);
  const bool l2 = p.metric == MEVI_METRIC_L2;
  const float mf = l2 ? -2.f : -1.f;
  const float m2inv = mf * p.consts[C_INV];
  const int d = p.d, nd = d / 128;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  unsigned n_ref = 0;
  for (long long slot = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_open; slot += warps) {
    const int row = p.open_rows[slot];
    const int4 meta = p.open_meta[slot];
    unsigned packed = (unsigned)meta.x;
    const int j0 = meta.y;
    const float xn = __int_as_float(meta.z), na = __int_as_float(meta.w), nxn = -xn;
    const float* xr = p.X + (int64_t)row * d;
    float4 xv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) xv[t] = t < nd ? __ldg(reinterpret_cast<const float4*>(xr + lane * 4 + 128 * t)) : make_float4(0.f, 0.f, 0.f, 0.f);
    int flag = -1;
    for (int j = j0; j < M && flag < 0; ++j) {
      const float acc = p.open_t1[(slot * M + j) * 32 + lane];
      const float* gj = sGram + (j * (j - 1) / 2) * K * (K + 1);
      float g = 0.f;
      for (int m = 0; m < j; ++m) g += gj[(m * K + (int)((packed >> (8 * m)) & 255u)) * (K + 1) + lane];
      const float base = l2 ? fmaf(2.f, g, sCn2[j * K + lane]) : g;
      const float dk = fmaf(acc, m2inv, base);
      float c1 = dk;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c1 = fminf(c1, __shfl_xor_sync(MEVI_FULL_MASK, c1, o));
      const unsigned bm = __ballot_sync(MEVI_FULL_MASK, dk == c1);
      int best = bm ? __ffs(bm) - 1 : 0;  // lowest index among equals (NaN rows: no lane matches -> unresolved below)
      const float u = fmaf(nxn, sEA2[j * K + lane], fmaf(na, sEA1[j * K + lane], dk));
      float other = lane == best ? CUDART_INF_F : u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) other = fminf(other, __shfl_xor_sync(MEVI_FULL_MASK, other, o));
      const float hi_best = c1 + fmaf(xn, sEA2[j * K + best], -na * sEA1[j * K + best]) + sLvl[j * 4 + 1];
      bool resolved = bm != 0u && other > hi_best;
      if (!resolved) {
        unsigned cm = __ballot_sync(MEVI_FULL_MASK, u <= hi_best);
        const int nc = __popc(cm);
        if (bm != 0u && nc >= 2 && nc <= MAX_REFINE) {
          float rb = CUDART_INF_F, vb = CUDART_INF_F, v1 = CUDART_INF_F, v2 = CUDART_INF_F;
          int rbi = 0;
          while (cm != 0u) {
            const int k = __ffs(cm) - 1;
            cm &= cm - 1;
            const float* cr = p.cb + (size_t)(j * K + k) * d + lane * 4;
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              if (t < nd) {
                const float4 cv = __ldg(reinterpret_cast<const float4*>(cr + 128 * t));
                a4.x = fmaf(xv[t].x, cv.x, a4.x);
                a4.y = fmaf(xv[t].y, cv.y, a4.y);
                a4.z = fmaf(xv[t].z, cv.z, a4.z);
                a4.w = fmaf(xv[t].w, cv.w, a4.w);
              }
            }
            float sdot = (a4.x + a4.y) + (a4.z + a4.w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(MEVI_FULL_MASK, sdot, o);
            const float dr = fmaf(sdot, mf, __shfl_sync(MEVI_FULL_MASK, base, k));
            const float v = fmaf(nxn, sE1[j * K + k], dr);
            if (dr < rb) { rb = dr; rbi = k; vb = v; }  // ascending k, strict: lowest index among equals
            v2 = fminf(v2, fmaxf(v1, v));
            v1 = fminf(v1, v);
          }
          resolved = ((vb == v1) ? v2 : v1) > rb + xn * sE1[j * K + rbi] + sLvl[j * 4 + 1];
          best = rbi;
          ++n_ref;
        }
      }
      packed = (packed & ~(255u << (8 * j))) | ((unsigned)best << (8 * j));
      if (!resolved) flag = j;  // the exact kernel re-decides the row from this level on
    }
    if (lane == 0) {
      int32_t* dst = p.codes + (int64_t)row * p.codes_stride;
#pragma unroll
      for (int j = 0; j < M; ++j) dst[j] = (int32_t)((packed >> (8 * j)) & 255u);
      if (flag >= 0) {
        const unsigned long long w = atomicAdd(p.work_count, 1ull);
        p.work_rows[w] = row;
        p.work_levels[w] = flag;
      }
    }
  }
  if (lane == 0 && n_ref != 0u) atomicAdd(p.refine_count, (unsigned long long)n_ref);
}

}  // namespace v6
