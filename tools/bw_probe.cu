// bw_probe.cu — how much HBM bandwidth does the RQ kernel's access pattern allow?
// Persistent CTAs read [128 rows x W bytes] pieces of a row-major fp32 matrix (row = 3072 B) with 16-byte
// loads, D pieces in flight per thread (register ring), no other work.  Compared with whole-row reads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float4 ldna(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// W = piece width in bytes (multiple of 256), T = threads, D = pieces in flight
template <int W, int T, int D>
__global__ void __launch_bounds__(T) tile_read(const float* __restrict__ X, int64_t n_rows, float* out) {
  constexpr int ROWB = 3072;
  constexpr int LANES_PER_ROW = W / 16;          // threads covering one row piece
  constexpr int ROWS_PER_PASS = T / LANES_PER_ROW;
  constexpr int U = 128 / ROWS_PER_PASS;         // loads per thread per piece
  constexpr int PIECES = ROWB / W;               // pieces per tile row
  const int64_t n_tiles = n_rows / 128;
  const int lr = threadIdx.x / LANES_PER_ROW, lc = threadIdx.x % LANES_PER_ROW;
  float4 ring[D][U];
  float acc = 0.f;
  int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  int64_t total = my_tiles * PIECES;
  auto load = [&](int64_t g, float4 (&v)[U]) {
    if (g >= total) return;
    int64_t it = g / PIECES; int pc = (int)(g - it * PIECES);
    int64_t row0 = (blockIdx.x + it * gridDim.x) * 128;
    const float* base = X + (row0 + lr) * (ROWB / 4) + pc * (W / 4) + lc * 4;
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldna(base + (int64_t)u * ROWS_PER_PASS * (ROWB / 4));
  };
#pragma unroll
  for (int d = 0; d < D - 1; ++d) load(d, ring[d]);
  for (int64_t g = 0; g < total; g += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      load(g + d + D - 1, ring[(d + D - 1) % D]);
      if (g + d < total) {
#pragma unroll
        for (int u = 0; u < U; ++u) acc += ring[d][u].x + ring[d][u].y + ring[d][u].z + ring[d][u].w;
      }
    }
  }
  if (acc == 1234.5f) out[0] = acc;
}

template <int W, int T, int D>
int run(const float* X, int64_t n, float* out, const char* name, int ctas_per_sm = 1, int smem = 0) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * ctas_per_sm;
  CK(cudaFuncSetAttribute(tile_read<W, T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tile_read<W, T, D><<<grid, T, smem>>>(X, n, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < 3; ++i) tile_read<W, T, D><<<grid, T, smem>>>(X, n, out);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
  printf("%-34s W=%4d T=%3d D=%d ctas/SM=%d smem=%6d : %7.3f ms  %7.1f GB/s\n", name, W, T, D, ctas_per_sm, smem, ms, n * 3072.0 / ms / 1e6);
  return 0;
}

int main() {
  const int64_t n = 4000000 / 128 * 128;
  float *X, *out; CK(cudaMalloc(&X, n * 3072)); CK(cudaMalloc(&out, 16)); CK(cudaMemset(X, 0, n * 3072));
  run<256, 256, 4>(X, n, out, "smem sweep", 1, 0);
  run<256, 256, 4>(X, n, out, "smem sweep", 1, 64 * 1024);
  run<256, 256, 4>(X, n, out, "smem sweep", 1, 150 * 1024);
  run<256, 256, 4>(X, n, out, "smem sweep", 1, 200 * 1024);
  run<256, 256, 4>(X, n, out, "smem sweep", 1, 221 * 1024);
  run<256, 512, 4>(X, n, out, "smem sweep T=512", 1, 221 * 1024);
  run<256, 256, 2>(X, n, out, "tile pieces (current kernel)");
  run<256, 256, 3>(X, n, out, "tile pieces");
  run<256, 256, 4>(X, n, out, "tile pieces");
  run<256, 512, 2>(X, n, out, "tile pieces");
  run<256, 512, 3>(X, n, out, "tile pieces");
  run<256, 512, 4>(X, n, out, "tile pieces");
  run<256, 256, 3>(X, n, out, "tile pieces, 2 CTAs/SM", 2);
  run<512, 256, 2>(X, n, out, "512 B pieces");
  run<512, 512, 3>(X, n, out, "512 B pieces");
  run<1024, 512, 2>(X, n, out, "1 KB pieces");
  run<3072, 768, 1>(X, n, out, "whole rows");
  run<3072, 768, 2>(X, n, out, "whole rows");
  return 0;
}
