// Hardware probe (not product code): tcgen05.mma kind::f16 M=128 K=16 issue/execute rate as a function of N
// and of how many independent TMEM accumulators the instruction stream rotates over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, int NACC, int ROWSHIFT, int TS, int NOISE>
__global__ void __launch_bounds__(128 + 512, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the operand area (values irrelevant)
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint32_t tmem = __reduce_or_sync(0xffffffffu, tmem_s);
    const uint32_t sb = __reduce_or_sync(0xffffffffu, smem_u32(smem));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = (uint64_t)(((sb + ROWSHIFT * 128) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t bdesc = (uint64_t)(((sb + 32768) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    long long t0 = clock64();
    if (elect()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (TS)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + a * N),
                           "r"(tmem + 448), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + a * N),
                           "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop_flag = 1; }
  }
  if (warp >= 4 && NOISE == 1) {
    // LDS.128 + STS.128 traffic on a private 64 KB region (rows swizzled like the epilogue)
    uint32_t base = smem_u32(smem) + 65536 + (threadIdx.x - 128) * 128;
    uint4 acc = make_uint4(0, 0, 0, 0);
    while (!stop_flag) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + ((c ^ (threadIdx.x & 7)) * 16)));
        acc.x += v.x; acc.y ^= v.y;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + ((c ^ (threadIdx.x & 7)) * 16)), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
      }
    }
    if (acc.x == 0x12345) out[3] = acc.y;
  }
  if (warp >= 4 && NOISE == 3) {
    // ALU/MUFU-heavy loop (like the gate epilogue): competes for issue slots with the MMA warp
    float a = threadIdx.x * 0.001f, b2 = 1.0f, c = 0.5f, e = 0.25f;
    while (!stop_flag) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a = __expf(fminf(a, 4.0f)) * 0.01f + c;
        b2 = fmaf(b2, 0.999f, a);
        c = fmaf(c, 0.5f, e);
        e = __fdividef(e + 1.0f, b2 + 2.0f);
      }
    }
    if (a + b2 + c + e == 0.12345f) out[3] = 1;
  }
  if (warp >= 4 && NOISE == 2) {
    // tcgen05.ld traffic on columns 256.. (not touched by the MMAs)
    uint32_t acc = 0;
    while (!stop_flag) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
            "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31])
          : "r"(tmem_s + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 1) * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) acc += r[i];
    }
    if (acc == 0x12345) out[3] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(512u) : "memory");
}

// the exact MMA1 operand pattern of iaf_flow_tc_kernel: hi / lo activation planes 80 KB apart, row-shifted
// windows, 3 taps x (lo.Wh, hi.Wl, hi.Wh) x 4 k-steps = 36 MMAs per task, one commit per task
template <int VARIANT>
__global__ void __launch_bounds__(640, 1) pattern_kernel(long long* out, int tasks) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[8];
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int c = 0; c < 8; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[c])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (208 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= 4 && VARIANT >= 10) {
    // warps parked in an mbarrier.try_wait spin (like the role warps of the flow kernel)
    uint32_t ok = 0;
    while (!ok) {
      if (VARIANT == 10 || (threadIdx.x & 31) == 0)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&dummy[7])), "r"(0u) : "memory");
      ok = __shfl_sync(0xffffffffu, ok, 0);
    }
  }
  if (warp == 0) {
    const uint32_t tmem = __reduce_or_sync(0xffffffffu, tmem_s);
    const uint32_t sb = __reduce_or_sync(0xffffffffu, smem_u32(smem));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [](uint32_t a) -> uint64_t {
      return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    };
    const uint32_t OFF_LO = 81920, OFF_WDH = 163840, OFF_WDL = OFF_WDH + 24576;
    long long t0 = clock64();
    for (int t = 0; t < tasks; ++t) {
      const int k = 3 - (t & 3);
      const int d = 4;
      if (elect()) {
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
          const int tap = 2 - seg;
          const uint32_t arow = (uint32_t)((1 + k) * 128 - seg * d) * 128u;
          const uint64_t alo = desc(sb + (VARIANT == 1 ? 0 : OFF_LO) + arow), ahi = desc(sb + arow);
          const uint64_t wh = desc(sb + OFF_WDH + tap * 8192), wl = desc(sb + (VARIANT == 2 ? OFF_WDH : OFF_WDL) + tap * 8192);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                         "l"(alo + 2 * kk), "l"(wh + 2 * kk), "r"(idesc), "r"(1u) : "memory");
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                         "l"(ahi + 2 * kk), "l"(wl + 2 * kk), "r"(idesc), "r"(1u) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                         "l"(ahi + 2 * kk), "l"(wh + 2 * kk), "r"(idesc), "r"(1u) : "memory");
          }
        }
        if (VARIANT >= 3 && VARIANT < 10) {
          const int nc = VARIANT == 3 ? 1 : (VARIANT == 4 ? 3 : 5);
          for (int c = 0; c < nc; ++c)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[c])) : "memory");
        }
      }
      __syncwarp();
    }
    if (elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t2 - t0; }
    if (VARIANT >= 10 && threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&dummy[7])) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(512u) : "memory");
}
template <int VARIANT>
void run_pattern_grid(const char* name, long long* d, int grid, int tasks) {
  const size_t smem = 208 * 1024 + 2048;
  cudaFuncSetAttribute(pattern_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  pattern_kernel<VARIANT><<<grid, 128, smem>>>(d, tasks);
  cudaEventRecord(e0);
  pattern_kernel<VARIANT><<<grid, 128, smem>>>(d, tasks);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s grid %3d tasks %6d: %.1f cyc per task (one CTA's clock64), %.3f ms -> %.0f MHz effective, %.1f ns per task\n", name, grid, tasks,
         (double)h[0] / tasks, ms, (double)h[0] / (ms * 1e3), ms * 1e6 / tasks);
}
template <int VARIANT>
void run_pattern(const char* name, long long* d) {
  const size_t smem = 208 * 1024 + 2048;
  cudaFuncSetAttribute(pattern_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    pattern_kernel<VARIANT><<<1, VARIANT >= 10 ? 640 : 128, smem>>>(d, 64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-50s: %.1f cyc per 36-MMA task = %.1f cyc/MMA\n", name, (double)h[0] / 64, (double)h[0] / 64 / 36);
}

// TMEM contention: 8 warps loop over tcgen05.ld.x32 (+ optional tcgen05.st.x32) on their own columns while
// warp 0 issues the flow kernel's MMA pattern (or just spins for the same time): iterations per warp tell how
// much the epilogues' TMEM traffic is slowed down by the accumulating MMAs and vice versa
template <int MMA_ON, int WITH_ST>
__global__ void __launch_bounds__(128 + 256, 1) tmem_contention_kernel(long long* out, int tasks) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop_flag = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (208 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint32_t tmem = __reduce_or_sync(0xffffffffu, tmem_s);
    const uint32_t sb = __reduce_or_sync(0xffffffffu, smem_u32(smem));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [](uint32_t a) -> uint64_t {
      return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    };
    const uint32_t OFF_LO = 81920, OFF_WDH = 163840, OFF_WDL = OFF_WDH + 24576;
    long long t0 = clock64();
    if (MMA_ON) {
      for (int t = 0; t < tasks; ++t) {
        const int k = 3 - (t & 3);
        if (elect()) {
#pragma unroll
          for (int seg = 0; seg < 3; ++seg) {
            const uint32_t arow = (uint32_t)((1 + k) * 128 - seg * 4) * 128u;
            const uint64_t alo = desc(sb + OFF_LO + arow), ahi = desc(sb + arow);
            const uint64_t wh = desc(sb + OFF_WDH + (2 - seg) * 8192), wl = desc(sb + OFF_WDL + (2 - seg) * 8192);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                           "l"(alo + 2 * kk), "l"(wh + 2 * kk), "r"(idesc), "r"(1u) : "memory");
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                           "l"(ahi + 2 * kk), "l"(wl + 2 * kk), "r"(idesc), "r"(1u) : "memory");
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (t & 1) * 64),
                           "l"(ahi + 2 * kk), "l"(wh + 2 * kk), "r"(idesc), "r"(1u) : "memory");
            }
          }
        }
        __syncwarp();
      }
      if (elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      __syncwarp();
      mbar_wait(&bar, 0);
    } else {
      while (clock64() - t0 < (long long)tasks * 1736) {}
    }
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t2 - t0; stop_flag = 1; }
  }
  if (warp >= 4) {
    long long iters = 0;
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = i;
    const uint32_t addr = tmem_s + ((uint32_t)((warp & 3) * 32) << 16) + 192 + ((warp >> 2) & 1) * 32;
    while (!stop_flag) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
            "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31])
          : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (WITH_ST) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
            "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
            "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++iters;
    }
    if (threadIdx.x == 128) out[2] = iters;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(512u) : "memory");
}
template <int MMA_ON, int WITH_ST>
void run_tmem(const char* name, long long* d) {
  const size_t smem = 208 * 1024 + 2048;
  cudaFuncSetAttribute(tmem_contention_kernel<MMA_ON, WITH_ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[3];
  for (int rep = 0; rep < 2; ++rep) {
    tmem_contention_kernel<MMA_ON, WITH_ST><<<1, 384, smem>>>(d, 64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-46s: %.1f cyc per 36-MMA task; each of 8 warps did one 4 KB TMEM ld%s every %.1f cycles\n", name, (double)h[0] / 64,
         WITH_ST ? "+st" : "", (double)h[0] / (double)(h[2] > 0 ? h[2] : 1));
}

__global__ void __launch_bounds__(128, 1) depth_kernel(long long* out, int n) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint32_t tmem = __reduce_or_sync(0xffffffffu, tmem_s);
    const uint32_t sb = __reduce_or_sync(0xffffffffu, smem_u32(smem));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t bdesc = (uint64_t)(((sb + 32768) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    long long t0 = clock64(), t1 = 0;
    if (elect()) {
      for (int i = 0; i < n; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(adesc + 2 * (i & 3)), "l"(bdesc + 2 * (i & 3)), "r"(idesc), "r"(1u) : "memory");
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      out[0] = t1 - t0;
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(512u) : "memory");
}

template <int N, int NACC, int ROWSHIFT, int TS, int NOISE = 0>
void run(const char* name, long long* d) {
  const int iters = 64;
  const size_t smem = 64 * 1024 + 1024 + 1024 + (NOISE == 1 ? 65536 : 0);
  cudaFuncSetAttribute(rate_kernel<N, NACC, ROWSHIFT, TS, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<N, NACC, ROWSHIFT, TS, NOISE><<<1, NOISE == 3 ? 640 : (NOISE ? 384 : 128), smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int n = iters * 4 * NACC;
  printf("%-34s N=%3d acc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%d MMAs)\n", name, N, NACC, (double)h[0] / n, (double)h[1] / n, n);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run_tmem<0, 0>("TMEM ld loop, no MMAs", d);
  run_tmem<1, 0>("TMEM ld loop + MMA pattern", d);
  run_tmem<0, 1>("TMEM ld+st loop, no MMAs", d);
  run_tmem<1, 1>("TMEM ld+st loop + MMA pattern", d);
  {
    const size_t smem = 64 * 1024 + 2048;
    cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ns[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
    for (int n : ns) {
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) { depth_kernel<<<1, 128, smem>>>(d, n); cudaDeviceSynchronize(); }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("queue depth probe: %2d SS MMAs (N=64) from an idle pipe: issue returns after %5lld cycles, all complete after %5lld\n", n, h[0], h[1]);
    }
  }
  run_pattern<0>("flow kernel MMA1 pattern", d);
  run_pattern<1>("pattern, lo plane aliased onto hi plane", d);
  run_pattern<2>("pattern, Wl aliased onto Wh", d);
  run_pattern_grid<0>("pattern, whole chip", d, 148, 20000);
  run_pattern_grid<0>("pattern, one CTA", d, 1, 20000);
  run_pattern_grid<0>("pattern, 74 CTAs", d, 74, 20000);
  run_pattern<10>("pattern + 16 warps (all lanes) in try_wait", d);
  run_pattern<11>("pattern + 16 warps (lane 0) in try_wait", d);
  run_pattern<3>("pattern + 1 commit per task", d);
  run_pattern<4>("pattern + 3 commits per task", d);
  run_pattern<5>("pattern + 5 commits per task", d);
  run<64, 1, 0, 0>("SS same accumulator", d);
  run<64, 2, 0, 0>("SS 2 accumulators", d);
  run<64, 4, 0, 0>("SS 4 accumulators", d);
  run<64, 1, 3, 0>("SS row-shifted A (3 rows)", d);
  run<128, 1, 0, 0>("SS same accumulator", d);
  run<128, 2, 0, 0>("SS 2 accumulators", d);
  run<256, 1, 0, 0>("SS same accumulator", d);
  run<32, 1, 0, 0>("SS same accumulator", d);
  run<64, 1, 0, 0, 1>("SS + 8 warps LDS/STS noise", d);
  run<64, 1, 0, 0, 2>("SS + 8 warps tcgen05.ld noise", d);
  run<64, 1, 0, 0, 3>("SS + 16 warps ALU/MUFU noise", d);
  run<64, 1, 0, 1, 3>("TS + 16 warps ALU/MUFU noise", d);
  run<64, 1, 0, 1>("TS (A from TMEM) same acc", d);
  run<64, 2, 0, 1>("TS (A from TMEM) 2 acc", d);
  run<128, 1, 0, 1>("TS (A from TMEM) same acc", d);
  return 0;
}
