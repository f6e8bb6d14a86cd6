// l2_reread_probe.cu — does a per-SM tile that is read twice in a row come from L2 the second time?
// Every CTA (one per SM, persistent) walks its tiles round-robin like the encode kernel; for each tile it
// streams the tile once, then (mode 1) streams it again.  Time vs tile size tells how large the per-SM tile may
// be before the second read falls out of the 126 MB L2 (footprint = 148 x tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_reread_probe tools/l2_reread_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(1024, 1) probe(const float4* __restrict__ X, int64_t n_tiles, int tile_f4, int passes,
                                                 float* out) {
  float acc = 0.f;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const float4* base = X + t * tile_f4;
    for (int p = 0; p < passes; ++p) {
      for (int i = threadIdx.x; i < tile_f4; i += 4096) {
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a, d = a;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(base + i));
        if (i + 1024 < tile_f4) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(base + i + 1024));
        if (i + 2048 < tile_f4) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(base + i + 2048));
        if (i + 3072 < tile_f4) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(base + i + 3072));
        acc += a.x + b.y + c.z + d.w;
      }
      __syncthreads();
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const size_t bytes = (size_t)8 << 30;  // 8 GB
  float4* X;
  float* out;
  cudaMalloc(&X, bytes);
  cudaMalloc(&out, 4);
  cudaMemset(X, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  const int tiles_kb[] = {96, 192, 384, 576, 768, 1024, 1536};
  for (int ti = 0; ti < 7; ++ti) {
    const int tile_f4 = tiles_kb[ti] * 1024 / 16;
    const int64_t n_tiles = bytes / ((size_t)tile_f4 * 16);
    for (int passes = 1; passes <= 2; ++passes) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<<<sms, 1024>>>(X, n_tiles, tile_f4, passes, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      printf("tile %5d KB (footprint %6.1f MB) passes %d: %.3f ms  -> %.2f TB/s of unique bytes\n", tiles_kb[ti],
             tiles_kb[ti] * sms / 1024.0, passes, best, bytes / best / 1e9);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
