// Hardware probe (not product code): tcgen05.mma.cta_group::2 mechanics and rate on B200.
//   D[256 x N] = A[256 x 64] . B[N x 64]^T   fp16 operands, fp32 accumulate, N = 128 or 256
// A CTA pair (cluster of 2): CTA r holds rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 (r+1)) of B in its own
// shared memory (K-major, SWIZZLE_128B, same offsets in both CTAs).  Both CTAs issue their TMA loads with the LEADER's
// mbarrier as completion target; the leader (rank 0) issues the MMAs; tcgen05.commit multicasts the completion to both
// CTAs; each CTA reads its 128 accumulator rows from its own TMEM.
// Part 1 checks the numbers, part 2 times bursts of MMAs (cycles per MMA) for cta_group::1 N=128 and cta_group::2
// N=128 / N=256.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_2cta_probe umma_2cta_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int KC = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 2000000000ll) return false;  // ~1 s: never hang the box
  }
  return true;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// mode 0: cta_group::1 reference (each CTA its own 128 x N GEMM against the FULL B), mode 1: cta_group::2
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_half,
             const __grid_constant__ CUtensorMap map_b_full, int mode, int N, int burst, float* out /*[256][N]*/,
             long long* cycles, int* status) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_rank();
  const uint32_t sA = smem_u32(smem);        // 128 x 64 fp16 = 16 KB
  const uint32_t sB = sA + 16384;            // up to 256 x 64 fp16 = 32 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool two = mode == 1;
  const uint32_t a_bytes = 16384, b_bytes = (uint32_t)(two ? N / 2 : N) * 128;
  if (threadIdx.x == 0) {
    // leader's bar_full: one arrival (its own expect_tx) covering the bytes of BOTH CTAs in 2-CTA mode
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (two) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();  // both CTAs' barriers exist before anybody signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;
  if (threadIdx.x == 0) {
    if (two) {
      const uint32_t leader_bar = mapa(smem_u32(&bar_full), 0);
      if (rank == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"(2 * (a_bytes + b_bytes)) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sA),
          "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(leader_bar), "r"(0), "r"((int)rank * 128)
          : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sB),
          "l"(reinterpret_cast<uint64_t>(&map_b_half)), "r"(leader_bar), "r"(0), "r"((int)rank * (N / 2))
          : "memory");
    } else {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"(a_bytes + b_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sA),
                   "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(smem_u32(&bar_full)), "r"(0), "r"((int)rank * 128)
                   : "memory");
      for (int h = 0; h < N / 128; ++h)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sB + h * 16384),
                     "l"(reinterpret_cast<uint64_t>(&map_b_full)), "r"(smem_u32(&bar_full)), "r"(0), "r"(h * 128)
                     : "memory");
    }
    if (!two || rank == 0) {
      ok = mbar_wait(&bar_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t M = two ? 256 : 128;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint64_t ad = desc_sw128(sA), bd = desc_sw128(sB);
      const long long t0 = clock64();
      for (int rep = 0; rep < burst && ok; ++rep) {
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (rep | k) != 0;
          if (two)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
                         : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
                         : "memory");
        }
      }
      if (two)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma)),
                     "h"((uint16_t)3)
                     : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
      ok = ok && mbar_wait(&bar_mma, 0);
      cycles[rank] = clock64() - t0;
    }
  }
  if (threadIdx.x != 0 || (two && rank == 1)) ok = mbar_wait(&bar_mma, 0);
  if (!ok) atomicExch(status, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok && burst == 1) {
    for (int c = 0; c < N; c += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) out[((size_t)rank * 128 + warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();  // the peer may still be reading operands of / signalling into this CTA
  if (warp == 0) {
    if (two) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                       \
  do {                                                              \
    cudaError_t e = (x);                                            \
    if (e != cudaSuccess) {                                         \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));         \
      return 1;                                                     \
    }                                                               \
  } while (0)

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int NMAX = 256;
  std::vector<__half> hA((size_t)256 * KC), hB((size_t)NMAX * KC);
  std::vector<float> fA(hA.size()), fB(hB.size());
  for (int r = 0; r < 256; ++r)
    for (int c = 0; c < KC; ++c) {
      fA[(size_t)r * KC + c] = (float)(((r * 7 + c * 3) % 17) - 8) * 0.125f;
      hA[(size_t)r * KC + c] = __float2half(fA[(size_t)r * KC + c]);
    }
  for (int n = 0; n < NMAX; ++n)
    for (int c = 0; c < KC; ++c) {
      fB[(size_t)n * KC + c] = (float)(((n * 5 + c * 11) % 13) - 6) * 0.25f;
      hB[(size_t)n * KC + c] = __float2half(fB[(size_t)n * KC + c]);
    }
  __half *dA, *dB;
  float* dO;
  long long* dC;
  int* dS;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, 256 * NMAX * 4));
  CK(cudaMalloc(&dC, 16));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int N : {128, 256}) {
    CUtensorMap mapA, mapBh, mapBf;
    const cuuint64_t strides[1] = {KC * 2};
    const cuuint32_t es[2] = {1, 1};
    {
      const cuuint64_t dims[2] = {KC, 256};
      const cuuint32_t box[2] = {KC, 128};
      if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
      const cuuint64_t dimsb[2] = {KC, (cuuint64_t)N};
      const cuuint32_t boxh[2] = {KC, (cuuint32_t)(N / 2)}, boxf[2] = {KC, 128};
      if (enc(&mapBh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsb, strides, boxh, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
      if (enc(&mapBf, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsb, strides, boxf, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    }
    for (int mode = 0; mode < 2; ++mode) {
      for (int burst : {1, 16, 64}) {
        CK(cudaMemset(dO, 0xff, 256 * NMAX * 4));
        CK(cudaMemset(dC, 0, 16));
        CK(cudaMemset(dS, 0, 4));
        probe_kernel<<<2, 128, smem>>>(mapA, mapBh, mapBf, mode, N, burst, dO, dC, dS);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d N %d burst %d: kernel error %s\n", mode, N, burst, cudaGetErrorString(e));
          return 2;
        }
        long long cyc[2];
        int st;
        CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        if (st) {
          printf("mode %d N %d burst %d: TIMEOUT waiting for a barrier\n", mode, N, burst);
          continue;
        }
        if (burst == 1) {
          std::vector<float> hO((size_t)256 * N);
          CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
          double maxerr = 0;
          int bad = 0;
          for (int r = 0; r < 256; ++r)
            for (int n = 0; n < N; ++n) {
              double ref = 0;
              for (int c = 0; c < KC; ++c) ref += (double)fA[(size_t)r * KC + c] * fB[(size_t)n * KC + c];
              const double err = fabs(ref - hO[(size_t)r * N + n]);
              if (!(err < 1e-3)) ++bad;
              if (err > maxerr || err != err) maxerr = err;
            }
          printf("%s N=%d: max err %.4g  bad %d / %d\n", mode ? "cta_group::2 (M=256)" : "cta_group::1 (2 x M=128)", N, maxerr,
                 bad, 256 * N);
        } else {
          printf("%s N=%d: %d MMAs issued + retired in %lld cycles = %.1f cycles per MMA (K=16)\n",
                 mode ? "cta_group::2" : "cta_group::1", N, 4 * burst, cyc[0], (double)cyc[0] / (4 * burst));
        }
      }
    }
  }
  return 0;
}
