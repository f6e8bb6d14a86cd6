// tmem_ld_rate.cu — what does the tensor-memory READ path (tcgen05.ld) sustain per SM?  The wide-codebook PQ kernel
// (pq_tensor.cuh) reads 24.6 KB of accumulators per row, so this is its roofline.
// One CTA per SM, W warps, each warp loops over tcgen05.ld.32x32b.xN of its lane quarter (+ wait), N = 16/32/64,
// with `depth` loads issued per wait.  Prints bytes / clock / SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_rate tools/tmem_ld_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<16>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(t) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t t, uint32_t* r) {
  ld<16>(t, r);
  ld<16>(t + 16, r + 16);
}
#define R8(b) "%" #b
template <int N, int DEPTH, bool X32>
__global__ void __launch_bounds__(512, 1) rate_kernel(int iters, int col_stride, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * col_stride) & 255);
  uint32_t acc = 0;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[DEPTH][N];
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd) {
      const uint32_t a = base + (uint32_t)(((it * DEPTH + dd) * N) & 255);
      if (X32 && N == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[dd][0]), "=r"(r[dd][1]), "=r"(r[dd][2]), "=r"(r[dd][3]), "=r"(r[dd][4]), "=r"(r[dd][5]), "=r"(r[dd][6]),
              "=r"(r[dd][7]), "=r"(r[dd][8]), "=r"(r[dd][9]), "=r"(r[dd][10]), "=r"(r[dd][11]), "=r"(r[dd][12]), "=r"(r[dd][13]),
              "=r"(r[dd][14]), "=r"(r[dd][15]), "=r"(r[dd][16 % N]), "=r"(r[dd][17 % N]), "=r"(r[dd][18 % N]), "=r"(r[dd][19 % N]),
              "=r"(r[dd][20 % N]), "=r"(r[dd][21 % N]), "=r"(r[dd][22 % N]), "=r"(r[dd][23 % N]), "=r"(r[dd][24 % N]),
              "=r"(r[dd][25 % N]), "=r"(r[dd][26 % N]), "=r"(r[dd][27 % N]), "=r"(r[dd][28 % N]), "=r"(r[dd][29 % N]),
              "=r"(r[dd][30 % N]), "=r"(r[dd][31 % N])
            : "r"(a) : "memory");
      } else {
#pragma unroll
        for (int c = 0; c < N; c += 16) ld<16>(a + c, &r[dd][c]);
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd)
#pragma unroll
      for (int i = 0; i < N; ++i) acc ^= r[dd][i];
  }
  const unsigned long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(holder) : "memory");
}

template <int N, int DEPTH, bool X32>
void run(const char* tag, int warps, int col_stride) {
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 20000;
  rate_kernel<N, DEPTH, X32><<<148, warps * 32, 0>>>(iters, col_stride, cyc, sink);
  rate_kernel<N, DEPTH, X32><<<148, warps * 32, 0>>>(iters, col_stride, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)h[i]; mean /= 148;
  const double bytes = (double)warps * iters * DEPTH * 32 * N * 4;
  printf("%-44s warps %2d  col_stride %3d : %7.1f B/clk/SM   (%s)\n", tag, warps, col_stride, bytes / mean, cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int warps : {4, 8, 16}) {
    run<16, 1, false>("x16, one load per wait", warps, 0);
    run<16, 2, false>("x16, two loads per wait", warps, 0);
    run<32, 1, true>("x32, one load per wait", warps, 0);
    run<32, 2, true>("x32, two loads per wait", warps, 0);
    run<32, 1, false>("2 x x16 per wait (32 columns)", warps, 0);
    run<64, 1, false>("4 x x16 per wait (64 columns)", warps, 0);
  }
  run<32, 1, true>("x32, warps of a quarter 64 columns apart", 8, 64);
  run<32, 1, true>("x32, warps of a quarter 128 columns apart", 8, 128);
  run<32, 1, true>("x32, warps of a quarter 128 columns apart", 16, 128);
  return 0;
}
