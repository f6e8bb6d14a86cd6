// umma_rate.cu — how many SM cycles one tcgen05.mma (kind::f16, M=128, K=16) really takes on this part, per operand
// source / N / swizzle.  One CTA per SM, one issuing lane, operands resident in shared memory (zeros), no other
// traffic: this is the floor the encode / flat-search pipelines are built against (DESIGN.md cites the output).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../mevi_b200/csrc/ptx.cuh"

enum Mode { SS_N128_SW64 = 0, SS_N256_SW64, SS_N128_SW128, SS_N256_SW128, TS_N128_SW64, TS_N256_SW64, SS_N128_TWO_ACC, SS_N64_SW64, NMODES };
static const char* names[NMODES] = {"SS  N=128 SW64", "SS  N=256 SW64", "SS  N=128 SW128", "SS  N=256 SW128",
                                    "TS  N=128 SW64 (A in TMEM)", "TS  N=256 SW64 (A in TMEM)", "SS  N=128 SW64, two accumulators alternating",
                                    "SS  N=64  SW64"};

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int rounds, long long* cycles, int bg_traffic, const uint8_t* gsrc, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, cbar[4];
  __shared__ volatile int done;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&bar2, 1); for (int i = 0; i < 4; ++i) ptx::mbar_init(&cbar[i], 1); done = 0; ptx::mbar_fence_init(); }
  if (warp == 0) ptx::tmem_alloc(&holder, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tm = holder;
  const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 32 * 1024;
  if (warp == 0) {
    const int N = (mode == SS_N256_SW64 || mode == SS_N256_SW128 || mode == TS_N256_SW64) ? 256 : (mode == SS_N64_SW64 ? 64 : 128);
    const uint32_t idesc = ptx::umma_idesc_f16_m128((uint32_t)N);
    const bool sw128 = mode == SS_N128_SW128 || mode == SS_N256_SW128;
    const bool ts = mode == TS_N128_SW64 || mode == TS_N256_SW64;
    long long t0 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t koff = (i & 1) * 32;
          const uint32_t d = tm + ((mode == SS_N128_TWO_ACC && (i & 2)) ? 128u : 0u);
          const uint64_t bd = sw128 ? ptx::umma_desc_sw128(b0 + koff) : ptx::umma_desc_sw64(b0 + koff);
          if (ts) ptx::umma_f16_ts(d, tm + 256 + (i & 1) * 8, bd, idesc, 1u);
          else ptx::umma_f16(d, sw128 ? ptx::umma_desc_sw128(a0 + koff) : ptx::umma_desc_sw64(a0 + koff), bd, idesc, 1u);
          if (commit_every > 0 && (i % commit_every) == commit_every - 1) ptx::umma_commit(&cbar[(i / commit_every) & 3]);
          if (commit_every < 0 && (i % (-commit_every)) == -commit_every - 1) { ptx::umma_commit(&cbar[0]); ptx::umma_commit(&cbar[1]); }
        }
      }
      ptx::umma_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    if (lane == 0) { cycles[blockIdx.x] = clock64() - t0; done = 1; }
  } else if (bg_traffic < 0) {
    // background async-proxy traffic: warp 1 keeps 2 x 16 KB bulk copies global -> shared in flight (what the TMA
    // producers of the encode kernel do), until the MMA warp is finished
    if (warp == 1 && lane == 0) {
      uint32_t ph = 0;
      const uint8_t* src = gsrc + (size_t)blockIdx.x * (1 << 20);
      uint32_t off = 0;
      while (!done) {
        ptx::mbar_arrive_expect_tx(&bar2, 32768);
        ptx::bulk_g2s(smem + 64 * 1024, src + off, 16384, &bar2);
        ptx::bulk_g2s(smem + 80 * 1024, src + off + 16384, 16384, &bar2);
        ptx::mbar_wait(&bar2, ph);
        ph ^= 1;
        off = (off + 32768) & ((1 << 20) - 1);
      }
    }
  } else if (bg_traffic) {
    // background shared-memory traffic from the other three warps: 16-byte loads + stores over a 32 KB window
    uint32_t base = ptx::smem_u32(smem) + 64 * 1024 + threadIdx.x * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = 0; r < rounds * bg_traffic; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + (i * 1536 * 2) % 32768));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(base), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

// Commit cadence study: groups of `g` MMAs (SS, N=128) followed by `nc` tcgen05.commit's onto barriers rotating over
// `nbar` mbarriers; with `wait` the issuing warp waits for the group's barrier before the next group (= latency).
__global__ void __launch_bounds__(128, 1) commit_kernel(int g, int nc, int nbar, int wait, int groups, int ts, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars[i], 1); ptx::mbar_fence_init(); }
  if (warp == 0) ptx::tmem_alloc(&holder, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tm = holder;
  const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 32 * 1024;
  if (warp == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16_m128(128u);
    uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    int bi = 0;
    for (int r = 0; r < groups; ++r) {
      if (ptx::elect_one()) {
        for (int i = 0; i < g; ++i) {
          const uint32_t koff = (i & 1) * 32;
          if (ts) ptx::umma_f16_ts(tm + (i & 2) * 64, tm + 256 + (i & 1) * 8, ptx::umma_desc_sw64(b0 + koff), idesc, 1u);
          else ptx::umma_f16(tm + (i & 2) * 64, ptx::umma_desc_sw64(a0 + koff), ptx::umma_desc_sw64(b0 + koff), idesc, 1u);
        }
        for (int c = 0; c < nc; ++c) ptx::umma_commit(&bars[(bi + c) % nbar]);
      }
      __syncwarp();
      if (wait) {
        for (int c = 0; c < nc; ++c) {
          const int b = (bi + c) % nbar;
          ptx::mbar_wait(&bars[b], cnt[b] & 1);
          cnt[b]++;
        }
      } else {
        for (int c = 0; c < nc; ++c) cnt[(bi + c) % nbar]++;
      }
      bi = (bi + nc) % nbar;
    }
    // drain: wait for the last arrival on every barrier that was used
    for (int b = 0; b < nbar; ++b)
      if (cnt[b] && !wait) ptx::mbar_wait(&bars[b], (cnt[b] - 1) & 1);
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
    (void)phase;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

// Two issuing warps, each with its own accumulator and its own commits: is the commit bubble a property of the
// issuing THREAD (then two streams overlap each other's bubbles) or of the tensor pipe (then nothing is gained)?
__global__ void __launch_bounds__(128, 1) commit2_kernel(int g, int nw, int groups, int ts, long long* cycles, int dmode, int pollers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4], done_bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&bars[i], 1); ptx::mbar_init(&done_bar, nw); ptx::mbar_fence_init(); }
  if (warp == 0) ptx::tmem_alloc(&holder, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tm = holder;
  const uint32_t a0 = ptx::smem_u32(smem) + warp * 8192, b0 = ptx::smem_u32(smem) + 32 * 1024;
  long long t0 = clock64();
  if (warp < nw) {
    const uint32_t idesc = ptx::umma_idesc_f16_m128(128u);
    for (int r = 0; r < groups; ++r) {
      if (ptx::elect_one()) {
        for (int i = 0; i < g; ++i) {
          const uint32_t koff = (i & 1) * 32;
          // dmode 0: one accumulator per warp; 1: alternate accumulators every 2 MMAs; 2: first half of the group
          // into one accumulator, second half into the other (the encode kernel's pattern)
          const uint32_t dsel = dmode == 0 ? 0u : dmode == 1 ? (uint32_t)((i >> 1) & 1) : (uint32_t)(i >= g / 2);
          const uint32_t d = tm + ((warp * 2 + dsel) & 3) * 64;
          if (ts) ptx::umma_f16_ts(d, tm + 256 + warp * 32 + (i & 1) * 8, ptx::umma_desc_sw64(b0 + koff), idesc, 1u);
          else ptx::umma_f16(d, ptx::umma_desc_sw64(a0 + koff), ptx::umma_desc_sw64(b0 + koff), idesc, 1u);
        }
        ptx::umma_commit(&bars[warp]);
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(&bars[2 + warp]);  // fresh barrier: completes after everything this warp issued
    __syncwarp();
    ptx::mbar_wait(&bars[2 + warp], 0);
    if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(&done_bar);
  } else if (pollers == 1) {
    ptx::mbar_wait(&done_bar, 0);           // every lane of the idle warps spins on try_wait (what the converters do)
  } else if (pollers == 2) {
    ptx::mbar_wait_backoff(&done_bar, 0, 64);  // same with a nanosleep between polls
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  long long* d_cycles;
  cudaMalloc(&d_cycles, nsm * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int rounds = 2000;
  uint8_t* d_src;
  cudaMalloc(&d_src, (size_t)nsm << 20);
  cudaMemset(d_src, 0, (size_t)nsm << 20);
  printf("device %s, %d SMs; %d MMAs per CTA per run; cycles per tcgen05.mma (M=128, K=16, fp16 -> fp32)\n", prop.name, nsm, rounds * 8);
  cudaFuncSetAttribute(commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(commit2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("two issuing warps (own accumulator + own commit each):\n");
  for (int dmode = 0; dmode <= 2; ++dmode)
  for (int ts = 1; ts <= 1; ++ts)
    for (int nw : {1, 2})
      for (int g : {6, 12}) {
        const int groups = 2000;
        commit2_kernel<<<nsm, 128, 100 * 1024>>>(g, nw, groups, ts, d_cycles, 2, dmode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
        long long h[256];
        cudaMemcpy(h, d_cycles, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("2W pollers=%d %s warps=%d g=%2d : %8.1f cycles per round (%d MMAs; %6.1f per MMA)\n", dmode, ts ? "TS" : "SS", nw, g, (double)mx / groups, nw * g, (double)mx / groups / (nw * g));
      }
  printf("commit cadence (N=128): cycles per group of g MMAs\n");
  for (int ts = 0; ts <= 1; ++ts)
    for (int wait = 0; wait <= 1; ++wait)
      for (int g : {1, 4, 6, 12, 24})
        for (int nc : {1, 2})
          for (int nbar : {1, 2, 4, 8}) {
            if (nbar < nc) continue;
            if (wait && nbar != 2) continue;
            const int groups = 2000;
            commit_kernel<<<nsm, 128, 100 * 1024>>>(g, nc, nbar, wait, groups, ts, d_cycles);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256];
            cudaMemcpy(h, d_cycles, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%s g=%2d commits=%d barriers=%d %s : %8.1f cycles/group (%6.1f per MMA; bare %d)\n", ts ? "TS" : "SS", g, nc, nbar,
                   wait ? "wait-each" : "streaming", (double)mx / groups, (double)mx / groups / g, g * (ts ? 74 : 107));
          }
  for (int mode : {(int)SS_N128_SW64, (int)TS_N128_SW64, (int)SS_N128_TWO_ACC, (int)SS_N256_SW64})
    for (int ce : {0, 8, 4, 2, 1, -8, -4}) {
      rate_kernel<<<nsm, 128, 100 * 1024>>>(mode, rounds, d_cycles, 0, d_src, ce);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256];
      cudaMemcpy(h, d_cycles, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-48s commit %s every %d MMAs : %7.1f cycles/MMA\n", names[mode], ce < 0 ? "x2" : "x1", ce < 0 ? -ce : ce, (double)mx / (rounds * 8));
    }
  for (int bg = 0; bg <= 3; ++bg)
    for (int mode = 0; mode < NMODES; ++mode) {
      if (bg >= 2 && mode != SS_N128_SW64 && mode != TS_N128_SW64) continue;
      for (int grid : {1, nsm}) {
        rate_kernel<<<grid, 128, 100 * 1024>>>(mode, rounds, d_cycles, bg == 1 ? 6 : (bg >= 2 ? -1 : 0), d_src, bg == 3 ? 4 : 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-48s : CUDA ERROR %s\n", names[mode], cudaGetErrorString(e)); return 1; }
        long long h[256];
        cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-48s grid %3d %s : %7.1f cycles/MMA\n", names[mode], grid, bg == 1 ? "+LSU traffic" : bg == 2 ? "+bulk copies" : bg == 3 ? "+bulk, commit/4" : "            ", (double)mx / (rounds * 8));
      }
    }
  return 0;
}
