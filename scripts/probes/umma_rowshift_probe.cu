// Hardware probe (not product code): can a tcgen05.mma A-operand descriptor start at an arbitrary
// ROW offset inside a long K-major fp16 matrix that TMA wrote to shared memory?
//   mode 0: SWIZZLE_128B layout, descriptor base_offset field = 0
//   mode 1: SWIZZLE_128B layout, base_offset = (start_addr >> 7) & 7
//   mode 2: no swizzle, [16-byte K chunk][row][8 halfs] layout, LBO = rows*16, SBO = 128
//   mode 3: as 2 with LBO / SBO exchanged
// Prints max |D - expected| for a list of row offsets.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_rowshift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int ROWS = 384;  // rows of A resident in smem
constexpr int KC = 64;     // K
constexpr int N = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int mode,
             int row_off, float* out /*[128][64]*/) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sA = smem_u32(smem);                    // ROWS * 128 bytes
  const uint32_t sW = sA + ROWS * 128;                   // 64 x 64 fp16, SW128 (8 KB)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"((uint32_t)(ROWS * 128 + 8192)) : "memory");
    if (mode < 2) {
      for (int b = 0; b < ROWS / 128; ++b)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sA + b * 16384),
                     "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(smem_u32(&bar_full)), "r"(0), "r"(b * 128)
                     : "memory");
    } else {
      // 3-D view (8 halfs, rows, 8 chunks): one box per 128 rows would scatter chunks with pitch 128*16, so
      // load the whole ROWS in one box per 128-row group but chunk-major over ALL rows needs pitch ROWS*16:
      // issue one box per (chunk): box = (8 halfs, ROWS rows, 1 chunk)  [ROWS > 256 is not a legal box -> split]
      for (int c = 0; c < 8; ++c)
        for (int b = 0; b < ROWS / 128; ++b)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(sA + c * (ROWS * 16) + b * 2048),
                       "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(smem_u32(&bar_full)), "r"(0), "r"(b * 128), "r"(c)
                       : "memory");
    }
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sW),
                 "l"(reinterpret_cast<uint64_t>(&map_w)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0)
                 : "memory");
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t wdesc = (uint64_t)((sW >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    for (int k = 0; k < 4; ++k) {
      uint64_t adesc;
      if (mode < 2) {
        const uint32_t a = sA + row_off * 128;
        adesc = (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        if (mode == 1) adesc |= (uint64_t)((a >> 7) & 7) << 49;
        adesc += 2 * k;  // +32 bytes along K
      } else {
        // chunk-major layout: element (row, chunk) at chunk*ROWS*16 + row*16; one MMA consumes K=16 = 2 chunks
        const uint32_t a = sA + row_off * 16 + (2 * k) * (ROWS * 16);
        const uint64_t lbo = (uint64_t)((ROWS * 16) >> 4), sbo = (uint64_t)(128 >> 4);
        adesc = (uint64_t)((a >> 4) & 0x3FFF) | ((mode == 2 ? lbo : sbo) << 16) | ((mode == 2 ? sbo : lbo) << 32) |
                ((uint64_t)1 << 46);
      }
      const uint32_t acc = k != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(adesc), "l"(wdesc + 2 * k), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                       \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int GROWS = 1024;
  std::vector<__half> hA((size_t)GROWS * KC), hW((size_t)N * KC);
  std::vector<float> fA(hA.size()), fW(hW.size());
  for (int r = 0; r < GROWS; ++r)
    for (int c = 0; c < KC; ++c) {
      fA[(size_t)r * KC + c] = (float)(((r * 7 + c * 3) % 17) - 8) * 0.125f;
      hA[(size_t)r * KC + c] = __float2half(fA[(size_t)r * KC + c]);
    }
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < KC; ++c) {
      fW[(size_t)n * KC + c] = (float)(((n * 5 + c * 11) % 13) - 6) * 0.25f;
      hW[(size_t)n * KC + c] = __float2half(fW[(size_t)n * KC + c]);
    }
  __half *dA, *dW;
  float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dO, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap mapA_sw, mapA_ns, mapW;
  {
    const cuuint64_t dims[2] = {KC, GROWS};
    const cuuint64_t strides[1] = {KC * 2};
    const cuuint32_t box[2] = {KC, 128}, es[2] = {1, 1};
    CUresult r = enc(&mapA_sw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A sw failed %d\n", (int)r); return 1; }
    const cuuint64_t dimsw[2] = {KC, N};
    const cuuint32_t boxw[2] = {KC, N};
    r = enc(&mapW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dW, dimsw, strides, boxw, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode W failed %d\n", (int)r); return 1; }
    // 3-D view of A: (8 halfs, rows, 8 chunks)
    const cuuint64_t dims3[3] = {8, GROWS, 8};
    const cuuint64_t strides3[2] = {KC * 2, 16};
    const cuuint32_t box3[3] = {8, 128, 1}, es3[3] = {1, 1, 1};
    r = enc(&mapA_ns, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dA, dims3, strides3, box3, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A ns failed %d\n", (int)r); return 1; }
  }
  const size_t smem = ROWS * 128 + 8192 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int offs[] = {0, 8, 64, 128, 1, 2, 3, 4, 5, 7, 9, 33, 127, 250};
  std::vector<float> hO(128 * N);
  for (int mode = 0; mode < 4; ++mode) {
    for (int off : offs) {
      CK(cudaMemset(dO, 0xff, 128 * N * 4));
      probe_kernel<<<1, 128, smem>>>(mode < 2 ? mapA_sw : mapA_ns, mapW, mode, off, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d off %d: kernel error %s\n", mode, off, cudaGetErrorString(e));
        return 2;
      }
      CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int c = 0; c < KC; ++c) ref += (double)fA[(size_t)(r + off) * KC + c] * fW[(size_t)n * KC + c];
          const double err = fabs(ref - hO[r * N + n]);
          if (!(err < 1e-3)) ++bad;
          if (err > maxerr || err != err) maxerr = err;
        }
      printf("mode %d row_off %3d: max err %.4g  bad %d / %d\n", mode, off, maxerr, bad, 128 * N);
    }
  }
  return 0;
}
