// umma_probe.cu — standalone hardware probe for the tcgen05 building blocks the RQ/flat kernels use.
// Not part of the library: it pins down, on a real B200, (1) which shared-memory matrix-descriptor
// encodings produce a correct D = A.B^T for K-major operands written by ordinary threads,
// (2) the accuracy of the split-fp16 (hi/lo, 3-term) contraction against float64, (3) bulk async
// copies with mbarrier completion.  Every wait has an iteration cap so a wrong guess cannot hang the GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

struct Cfg {
  int kind;        // 0 = f16, 1 = tf32
  int N;           // 128 or 256
  int Ktot;        // elements
  int swz;         // 0 none, 2 = 64B, 3 = 128B
  int layout_type; // descriptor bits 61..63
  uint32_t lbo;    // 16-byte units
  uint32_t sbo;    // 16-byte units
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo & 0x3FFF) << 16;
  d |= (uint64_t)(sbo & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ bool mbar_wait_capped(uint32_t bar, uint32_t parity, int cap) {
  for (int i = 0; i < cap; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// operands in global: A [128][Ktot], B [N][Ktot] row-major, element = half (kind 0) or float (kind 1)
template <int KIND>
__global__ void __launch_bounds__(128) probe_kernel(const void* __restrict__ Ag, const void* __restrict__ Bg, float* __restrict__ Dg,
                                                    Cfg cfg, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_holder;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int esz = KIND == 0 ? 2 : 4;
  const int atom_bytes = cfg.swz == 3 ? 128 : (cfg.swz == 2 ? 64 : 16);
  const int chunk_elems = atom_bytes / esz;  // K elements per chunk
  const int nchunks = cfg.Ktot / chunk_elems;
  uint8_t* sA = smem;                                   // [nchunks][128][atom_bytes]
  uint8_t* sB = smem + (size_t)nchunks * 128 * atom_bytes;  // [nchunks][N][atom_bytes]
  const int units = atom_bytes / 16;

  // ---- fill operands (generic-proxy stores), 16 bytes at a time
  auto fill = [&](uint8_t* dst, const uint8_t* src, int rows) {
    const int units_per_row = cfg.Ktot * esz / 16;
    for (int i = tid; i < rows * units_per_row; i += blockDim.x) {
      const int row = i / units_per_row, ug = i % units_per_row;
      const int chunk = ug / units, u = ug % units;
      int up = u;
      if (cfg.swz == 3) up = u ^ (row & 7);
      else if (cfg.swz == 2) up = u ^ ((row >> 1) & 3);
      const uint4 v = *reinterpret_cast<const uint4*>(src + ((size_t)row * units_per_row + ug) * 16);
      *reinterpret_cast<uint4*>(dst + ((size_t)chunk * rows + row) * atom_bytes + up * 16) = v;
    }
  };
  fill(sA, (const uint8_t*)Ag, 128);
  fill(sB, (const uint8_t*)Bg, cfg.N);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_holder)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> visible to the async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_holder;

  if (tid == 0) {
    uint32_t idesc = (1u << 4) | ((KIND == 0 ? 0u : 2u) << 7) | ((KIND == 0 ? 0u : 2u) << 10) | ((uint32_t)(cfg.N >> 3) << 17) | ((128u >> 4) << 24);
    const int kstep_bytes = 32;  // one UMMA K step = 16 halves or 8 tf32
    const int ksteps = atom_bytes >= 32 ? atom_bytes / kstep_bytes : 0;
    uint32_t accum = 0;
    if (cfg.swz != 0) {
      for (int c = 0; c < nchunks; ++c)
        for (int ks = 0; ks < ksteps; ++ks) {
          uint64_t da = make_desc(smem_u32(sA + (size_t)c * 128 * atom_bytes) + ks * kstep_bytes, cfg.lbo, cfg.sbo, cfg.layout_type);
          uint64_t db = make_desc(smem_u32(sB + (size_t)c * cfg.N * atom_bytes) + ks * kstep_bytes, cfg.lbo, cfg.sbo, cfg.layout_type);
          umma<KIND>(tmem, da, db, idesc, accum);
          accum = 1;
        }
    } else {
      // no swizzle: [K/8 units][rows][16 B]; one K step = 2 units; unit stride = rows*16 bytes
      const int total_steps = cfg.Ktot * esz / 32;
      for (int s = 0; s < total_steps; ++s) {
        uint64_t da = make_desc(smem_u32(sA + (size_t)(2 * s) * 128 * 16), cfg.lbo, cfg.sbo, cfg.layout_type);
        // for B the K-unit stride differs when N != 128: recompute lbo from rows if lbo==128*16/16
        uint32_t lbo_b = cfg.lbo == (128 * 16 / 16) ? (uint32_t)(cfg.N * 16 / 16) : cfg.lbo;
        uint32_t sbo_b = cfg.sbo == (128 * 16 / 16) ? (uint32_t)(cfg.N * 16 / 16) : cfg.sbo;
        uint64_t db = make_desc(smem_u32(sB + (size_t)(2 * s) * cfg.N * 16), lbo_b, sbo_b, cfg.layout_type);
        umma<KIND>(tmem, da, db, idesc, accum);
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  __syncwarp();
  const bool done = mbar_wait_capped(smem_u32(&mbar), 0, 1 << 22);
  if (!done && tid == 0) *status = -1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (done) {
    for (int c0 = 0; c0 < cfg.N; c0 += 32) {
      uint32_t r[32];
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                     "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                     "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 32; ++j) Dg[(size_t)row * cfg.N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// bulk async copy probe: global -> shared with mbarrier complete_tx
__global__ void bulk_copy_probe(const float* __restrict__ src, float* __restrict__ dst, int n_floats, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = n_floats * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(src), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
  }
  const bool done = mbar_wait_capped(smem_u32(&mbar), 0, 1 << 22);
  if (!done && threadIdx.x == 0) *status = -1;
  if (done)
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) dst[i] = reinterpret_cast<float*>(smem)[i];
}


// ---- A operand from tensor memory: each thread writes its row of A (fp16, K-major) into TMEM with
// tcgen05.st, two consecutive K elements per 32-bit column; B stays in shared memory (SW128). -------------
template <int PACK>
__global__ void __launch_bounds__(128) probe_ts_kernel(const __half* __restrict__ Ag, const __half* __restrict__ Bg,
                                                      float* __restrict__ Dg, int N, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_holder;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 64;
  // B tile [N][64 halfs] SW128
  for (int i = tid; i < N * 8; i += blockDim.x) {
    const int row = i / 8, u = i % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(Bg + (size_t)row * K + u * 8);
    *reinterpret_cast<uint4*>(smem + (size_t)row * 128 + ((u ^ (row & 7)) * 16)) = v;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_holder)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_holder;
  const uint32_t a_col = 256;
  {  // row tid of A -> 32 columns starting at a_col, lane = tid
    uint32_t r[32];
    const __half* arow = Ag + (size_t)tid * K;
    for (int c = 0; c < 32; ++c) {
      const uint32_t lo = __half_as_ushort(arow[2 * c]), hi = __half_as_ushort(arow[2 * c + 1]);
      r[c] = PACK == 0 ? (lo | (hi << 16)) : (hi | (lo << 16));
    }
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + a_col;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
                    "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                    "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t db = make_desc(smem_u32(smem) + ks * 32, 1, 64, 2);
      const uint32_t a_addr = tmem + a_col + ks * 8;
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(a_addr), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  __syncwarp();
  const bool done = mbar_wait_capped(smem_u32(&mbar), 0, 1 << 22);
  if (!done && tid == 0) *status = -1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (done) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                     "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                     "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) Dg[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int PACK>
static int run_ts(const char* name) {
  const int N = 128, K = 64;
  std::vector<__half> A((size_t)128 * K), B((size_t)N * K);
  std::vector<double> Ar(A.size()), Br(B.size());
  srand(7);
  for (size_t i = 0; i < A.size(); ++i) { float v = (float)((rand() % 2001) - 1000) / 1000.f; A[i] = __float2half_rn(v); Ar[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { float v = (float)((rand() % 2001) - 1000) / 1000.f; B[i] = __float2half_rn(v); Br[i] = __half2float(B[i]); }
  __half *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, 128 * N * 4)); CK(cudaMemset(dS, 0, 4));
  const size_t smem = (size_t)N * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_ts_kernel<PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_ts_kernel<PACK><<<1, 128, smem>>>(dA, dB, dD, N, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s : CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<float> D((size_t)128 * N); int st;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += Ar[(size_t)i * K + k] * Br[(size_t)j * K + k];
    double err = fabs(D[(size_t)i * N + j] - ref); if (!(err <= 1e30)) err = 1e30;
    if (err > maxerr) maxerr = err; if (err > 1e-2 * (1 + fabs(ref))) ++bad;
  }
  printf("%-44s : %s  status=%d max_abs_err=%.3e bad=%d\n", name, (bad == 0 && st == 0) ? "PASS" : "FAIL", st, maxerr, bad);
  return 0;
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int KIND>
static int run_cfg(const char* name, Cfg cfg, const std::vector<float>& A, const std::vector<float>& B, bool quiet = false, double* out_maxerr = nullptr, std::vector<float>* out_D = nullptr) {
  const int K = cfg.Ktot, N = cfg.N;
  const int esz = KIND == 0 ? 2 : 4;
  std::vector<uint8_t> Ah((size_t)128 * K * esz), Bh((size_t)N * K * esz);
  std::vector<double> Ar((size_t)128 * K), Br((size_t)N * K), Art, Brt;
  if (KIND == 0) {
    for (size_t i = 0; i < Ar.size(); ++i) { __half h = __float2half_rn(A[i]); ((__half*)Ah.data())[i] = h; Ar[i] = __half2float(h); }
    for (size_t i = 0; i < Br.size(); ++i) { __half h = __float2half_rn(B[i]); ((__half*)Bh.data())[i] = h; Br[i] = __half2float(h); }
  } else {
    Art.resize(Ar.size()); Brt.resize(Br.size());
    for (size_t i = 0; i < Ar.size(); ++i) { ((float*)Ah.data())[i] = A[i]; Ar[i] = tf32_trunc(A[i]); Art[i] = tf32_rn(A[i]); }
    for (size_t i = 0; i < Br.size(); ++i) { ((float*)Bh.data())[i] = B[i]; Br[i] = tf32_trunc(B[i]); Brt[i] = tf32_rn(B[i]); }
  }
  void *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, Ah.size())); CK(cudaMalloc(&dB, Bh.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, Ah.data(), Ah.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bh.data(), Bh.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)128 * N * 4)); CK(cudaMemset(dS, 0, 4));
  size_t smem = (size_t)(128 + N) * K * esz + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<KIND><<<1, 128, smem>>>(dA, dB, dD, cfg, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s : CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
  int st; std::vector<float> D((size_t)128 * N);
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxerr_rn = 0, maxref = 0; int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      double ref = 0, ref2 = 0;
      for (int k = 0; k < K; ++k) ref += Ar[(size_t)i * K + k] * Br[(size_t)j * K + k];
      if (KIND == 1) for (int k = 0; k < K; ++k) ref2 += Art[(size_t)i * K + k] * Brt[(size_t)j * K + k];
      double got = D[(size_t)i * N + j];
      double err = fabs(got - ref); if (!(err <= 1e30)) err = 1e30;
      if (err > maxerr) maxerr = err;
      if (KIND == 1) { double e2 = fabs(got - ref2); if (e2 > maxerr_rn) maxerr_rn = e2; }
      if (fabs(ref) > maxref) maxref = fabs(ref);
      if (err > 1e-2 * (1 + fabs(ref))) ++bad;
    }
  if (!quiet) {
    if (KIND == 0) printf("%-44s : %s  status=%d max_abs_err=%.3e max|ref|=%.2f bad=%d\n", name, (bad == 0 && st == 0) ? "PASS" : "FAIL", st, maxerr, maxref, bad);
    else printf("%-44s : %s  status=%d err_vs_trunc=%.3e err_vs_rn=%.3e max|ref|=%.2f bad=%d\n", name, (bad == 0 && st == 0) ? "PASS" : "FAIL", st, maxerr, maxerr_rn, maxref, bad);
  }
  if (out_maxerr) *out_maxerr = maxerr;
  if (out_D) *out_D = D;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return (bad == 0 && st == 0) ? 0 : 1;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s cc %d.%d SMs %d smem/block optin %zu\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  srand(1);
  auto rnd = [] { return (float)((rand() % 2001) - 1000) / 1000.0f; };
  const int K = 128;
  std::vector<float> A((size_t)128 * K), B((size_t)256 * K);
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();

  run_ts<0>("A in TMEM (tcgen05.st, k even in low half)");
  run_ts<1>("A in TMEM (tcgen05.st, k even in high half)");
  // ---- 1. descriptor encodings, fp16
  run_cfg<0>("f16 SW128 N=128 lbo=1 sbo=64 lt=2", Cfg{0, 128, K, 3, 2, 1, 64}, A, B);
  run_cfg<0>("f16 SW128 N=256 lbo=1 sbo=64 lt=2", Cfg{0, 256, K, 3, 2, 1, 64}, A, B);
  run_cfg<0>("f16 SW128 N=128 lbo=0 sbo=64 lt=2", Cfg{0, 128, K, 3, 2, 0, 64}, A, B);
  run_cfg<0>("f16 SW64  N=128 lbo=1 sbo=32 lt=4", Cfg{0, 128, K, 2, 4, 1, 32}, A, B);
  run_cfg<0>("f16 SW64  N=256 lbo=1 sbo=32 lt=4", Cfg{0, 256, K, 2, 4, 1, 32}, A, B);
  run_cfg<0>("f16 NOSWZ N=128 lbo=128 sbo=8 lt=0", Cfg{0, 128, K, 0, 0, 128, 8}, A, B);
  // (lbo=8, sbo=128 — the two strides swapped — faults with an illegal memory access on B200: not run)
  run_cfg<0>("f16 NOSWZ N=256 lbo=128 sbo=8 lt=0", Cfg{0, 256, K, 0, 0, 128, 8}, A, B);
  // ---- 2. tf32 (operands are raw fp32 in smem): truncation or rounding of the low 13 bits?
  {
    std::vector<float> A2((size_t)128 * 64), B2((size_t)256 * 64);
    for (auto& v : A2) v = rnd() * 1.2345678f;
    for (auto& v : B2) v = rnd() * 0.7654321f;
    run_cfg<1>("tf32 SW128 N=128 K=64 lbo=1 sbo=64 lt=2", Cfg{1, 128, 64, 3, 2, 1, 64}, A2, B2);
    run_cfg<1>("tf32 SW128 N=256 K=64 lbo=1 sbo=64 lt=2", Cfg{1, 256, 64, 3, 2, 1, 64}, A2, B2);
  }
  // ---- 3. split-fp16 accuracy at K=768 (hi.hi + hi.lo + lo.hi), N(0,1) rows vs small-norm centroids
  {
    const int K7 = 768;
    std::vector<float> X((size_t)128 * K7), C((size_t)128 * K7);
    auto gauss = [] { double u1 = (rand() + 1.0) / (RAND_MAX + 2.0), u2 = (rand() + 1.0) / (RAND_MAX + 2.0); return (float)(sqrt(-2 * log(u1)) * cos(6.283185307179586 * u2)); };
    for (auto& v : X) v = gauss();
    for (auto& v : C) v = 0.06f * gauss();
    // power-of-two scaling into the fp16 sweet spot, then hi/lo split
    const float sx = 256.f, sc = 4096.f;
    std::vector<float> Xhi(X.size()), Xlo(X.size()), Chi(C.size()), Clo(C.size());
    for (size_t i = 0; i < X.size(); ++i) { float t = X[i] * sx; float h = __half2float(__float2half_rn(t)); Xhi[i] = h; Xlo[i] = __half2float(__float2half_rn(t - h)); }
    for (size_t i = 0; i < C.size(); ++i) { float t = C[i] * sc; float h = __half2float(__float2half_rn(t)); Chi[i] = h; Clo[i] = __half2float(__float2half_rn(t - h)); }
    std::vector<float> D1, D2, D3; double e;
    Cfg c7{0, 128, K7, 3, 2, 1, 64};
    // A operand smem = 128*768*2 = 196 KB + B 196 KB > 227 KB: run K in two halves of 384 and add on the host in double
    auto half_run = [&](const std::vector<float>& P, const std::vector<float>& Qm, std::vector<double>& acc) {
      for (int h = 0; h < 2; ++h) {
        std::vector<float> Ph((size_t)128 * 384), Qh((size_t)128 * 384);
        for (int i = 0; i < 128; ++i) for (int k = 0; k < 384; ++k) { Ph[(size_t)i * 384 + k] = P[(size_t)i * K7 + h * 384 + k]; Qh[(size_t)i * 384 + k] = Qm[(size_t)i * K7 + h * 384 + k]; }
        std::vector<float> Dh; Cfg ch{0, 128, 384, 3, 2, 1, 64};
        run_cfg<0>("", ch, Ph, Qh, true, &e, &Dh);
        for (size_t i = 0; i < Dh.size(); ++i) acc[i] += Dh[i];
      }
    };
    std::vector<double> hh(128 * 128, 0.0), hl(128 * 128, 0.0), lh(128 * 128, 0.0);
    half_run(Xhi, Chi, hh); half_run(Xhi, Clo, hl); half_run(Xlo, Chi, lh);
    double worst1 = 0, worst3 = 0, worst3f = 0;
    for (int i = 0; i < 128; ++i) {
      double xn = 0; for (int k = 0; k < K7; ++k) xn += (double)X[(size_t)i * K7 + k] * X[(size_t)i * K7 + k];
      for (int j = 0; j < 128; ++j) {
        double cn = 0, ref = 0;
        for (int k = 0; k < K7; ++k) { cn += (double)C[(size_t)j * K7 + k] * C[(size_t)j * K7 + k]; ref += (double)X[(size_t)i * K7 + k] * C[(size_t)j * K7 + k]; }
        const double scale = sqrt(xn) * sqrt(cn), inv = 1.0 / ((double)sx * sc);
        double a1 = hh[i * 128 + j] * inv, a3 = (hh[i * 128 + j] + hl[i * 128 + j] + lh[i * 128 + j]) * inv;
        float a3f = ((float)hh[i * 128 + j] + ((float)hl[i * 128 + j] + (float)lh[i * 128 + j])) * (float)inv;
        worst1 = fmax(worst1, fabs(a1 - ref) / scale); worst3 = fmax(worst3, fabs(a3 - ref) / scale); worst3f = fmax(worst3f, fabs((double)a3f - ref) / scale);
      }
    }
    printf("split-fp16 K=768 (two K=384 tensor passes per term): max |err|/(|x||c|): 1-term 2^%.2f, 3-term 2^%.2f, 3-term fp32-combined 2^%.2f\n",
           log2(worst1), log2(worst3), log2(worst3f));
  }
  // ---- 4. bulk copy
  {
    const int n = 8192; std::vector<float> h(n); for (int i = 0; i < n; ++i) h[i] = (float)i;
    float *s, *d; int* st; CK(cudaMalloc(&s, n * 4)); CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&st, 4));
    CK(cudaMemcpy(s, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(d, 0, n * 4)); CK(cudaMemset(st, 0, 4));
    CK(cudaFuncSetAttribute(bulk_copy_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4 + 1024));
    bulk_copy_probe<<<1, 128, n * 4 + 1024>>>(s, d, n, st);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(n); int sth = 0; cudaMemcpy(o.data(), d, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&sth, st, 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < n; ++i) bad += (o[i] != h[i]);
    printf("cp.async.bulk 32 KB global->smem + mbarrier complete_tx : %s (err=%s status=%d bad=%d)\n", (e == cudaSuccess && bad == 0 && sth == 0) ? "PASS" : "FAIL", cudaGetErrorString(e), sth, bad);
  }
  return 0;
}
