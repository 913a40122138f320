// Hardware probe (not product code): latency of a cross-SM value exchange through L2 tagged words,
// the mechanism fastgen_kernel uses once per layer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o xchg_probe xchg_probe.cu
// mode 0: ping-pong between CTA 0 and CTA k (one-way latency = round / 2)
// mode 1: all-to-all over NC CTAs, 768 tagged 8-byte entries per round (6 per CTA, R replicas),
//         polled by 128 threads x 3 x 16 B exactly like the product kernel, no compute in between.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

enum { ST_RELAXED = 0, ST_RELEASE, ST_FENCE, ST_ATOM, ST_RED, ST_VOLATILE, ST_RELAXED_SYS, ST_N };
enum { LD_RELAXED = 0, LD_ACQUIRE, LD_VOLATILE, LD_ATOM, LD_CG, LD_N };
static const char* st_names[] = {"st.relaxed.gpu", "st.release.gpu", "st.relaxed+fence", "atom.exch", "red.max", "st.volatile", "st.relaxed.sys"};
static const char* ld_names[] = {"ld.relaxed.gpu", "ld.acquire.gpu", "ld.volatile", "atom.or0", "ld.cg"};

__device__ __forceinline__ void put(unsigned long long* p, uint32_t v, uint32_t tag, int kind) {
  const unsigned long long w = ((unsigned long long)tag << 32) | v;
  switch (kind) {
    case ST_RELAXED: asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); break;
    case ST_RELEASE: asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); break;
    case ST_FENCE:
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      break;
    case ST_ATOM: {
      unsigned long long old;
      asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(w) : "memory");
      break;
    }
    case ST_RED: asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); break;
    case ST_VOLATILE: asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); break;
    default: asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); break;
  }
}
__device__ __forceinline__ uint4 get16(const unsigned long long* p, int kind) {
  uint4 r;
  switch (kind) {
    case LD_RELAXED: asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); break;
    case LD_ACQUIRE: {
      unsigned long long a, b;
      asm volatile("ld.acquire.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
      r.x = (uint32_t)a; r.y = (uint32_t)(a >> 32); r.z = (uint32_t)b; r.w = (uint32_t)(b >> 32);
      break;
    }
    case LD_VOLATILE: asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); break;
    case LD_ATOM: {
      unsigned long long a, b;
      asm volatile("atom.relaxed.gpu.global.or.b64 %0, [%1], 0;" : "=l"(a) : "l"(p) : "memory");
      asm volatile("atom.relaxed.gpu.global.or.b64 %0, [%1], 0;" : "=l"(b) : "l"(p + 1) : "memory");
      r.x = (uint32_t)a; r.y = (uint32_t)(a >> 32); r.z = (uint32_t)b; r.w = (uint32_t)(b >> 32);
      break;
    }
    default: asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); break;
  }
  return r;
}

__global__ void pingpong(unsigned long long* buf, int partner, int iters, int stk, int ldk, long long* out) {
  if (threadIdx.x != 0) return;
  const int c = blockIdx.x;
  if (c != 0 && c != partner) return;
  unsigned long long* mine = buf + (c == 0 ? 0 : 64);
  const unsigned long long* theirs = buf + (c == 0 ? 64 : 0);
  const long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (c == 0) { put(mine, i, i, stk); put(mine + 1, i, i, stk); }
    for (;;) {
      const uint4 r = get16(theirs, ldk);
      if (r.y == (uint32_t)i && r.w == (uint32_t)i) break;
    }
    if (c != 0) { put(mine, i, i, stk); put(mine + 1, i, i, stk); }
  }
  if (c == 0) out[0] = clock64() - t0;
}

// all-to-all: XS entries per slot, every CTA owns XS/NC of them; slots rotate over NSLOT rounds
constexpr int XS = 768, NSLOT = 34, MAXREP = 128;
__global__ void __launch_bounds__(256, 1)
alltoall(unsigned long long* buf, int iters, int stk, int ldk, int nrep, int npoll, int work, long long* out) {
  const int tid = threadIdx.x, c = blockIdx.x, NC = gridDim.x;
  const int per = XS / NC;  // entries this CTA publishes
  __shared__ float sink[1024];
  long long t0 = 0;
  float acc = 0.f;
  for (int i = 1; i <= iters; ++i) {
    if (i == 17) t0 = clock64();
    unsigned long long* slot = buf + (size_t)(i % NSLOT) * MAXREP * XS;
    if (tid < 128) {
      // "compute" stand-in, then publish (first `per` warps' lanes < nrep, like the product kernel)
      for (int k = 0; k < work; ++k) acc = fmaf(acc, 1.0001f, 0.5f);
      const int w = tid >> 5, lane = tid & 31;
      if (lane < nrep)
        for (int e = w; e < per; e += 4) put(slot + (size_t)lane * XS + c * per + e, (uint32_t)(c + (acc > 1e30f)), (uint32_t)i, stk);
    } else if (tid - 128 < npoll) {
      const unsigned long long* rp = slot + (size_t)(c & (nrep - 1)) * XS;
      const int k = tid - 128;
      // npoll threads cover 384 16-byte words
      for (int j = k; j < XS / 2; j += npoll) {
        for (;;) {
          const uint4 r = get16(rp + 2 * j, ldk);
          if (r.y == (uint32_t)i && r.w == (uint32_t)i) { sink[(2 * j) & 1023] = __uint_as_float(r.x); break; }
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) out[c] = clock64() - t0;
  if (acc == 123.f) out[0] = 0;
}

// variant: the 3 loads of a poll thread are issued together (like the product kernel)
__global__ void __launch_bounds__(256, 1)
alltoall3(unsigned long long* buf, int iters, int stk, int ldk, int nrep, int work, long long* out) {
  const int tid = threadIdx.x, c = blockIdx.x, NC = gridDim.x;
  const int per = XS / NC;
  __shared__ float sink[1024];
  long long t0 = 0;
  float acc = 0.f;
  for (int i = 1; i <= iters; ++i) {
    if (i == 17) t0 = clock64();
    unsigned long long* slot = buf + (size_t)(i % NSLOT) * MAXREP * XS;
    if (tid < 128) {
      for (int k = 0; k < work; ++k) acc = fmaf(acc, 1.0001f, 0.5f);
      const int w = tid >> 5, lane = tid & 31;
      for (int rep = lane; rep < nrep; rep += 32)
        for (int e = w; e < per; e += 4) put(slot + (size_t)rep * XS + c * per + e, (uint32_t)(c + (acc > 1e30f)), (uint32_t)i, stk);
    } else {
      const unsigned long long* rp = slot + (size_t)(c & (nrep - 1)) * XS;
      const int k = tid - 128;
      bool ok0 = false, ok1 = false, ok2 = false;
      uint4 r0, r1, r2;
      for (;;) {
        if (!ok0) r0 = get16(rp + 2 * k, ldk);
        if (!ok1) r1 = get16(rp + 256 + 2 * k, ldk);
        if (!ok2) r2 = get16(rp + 512 + 2 * k, ldk);
        ok0 = ok0 || (r0.y == (uint32_t)i && r0.w == (uint32_t)i);
        ok1 = ok1 || (r1.y == (uint32_t)i && r1.w == (uint32_t)i);
        ok2 = ok2 || (r2.y == (uint32_t)i && r2.w == (uint32_t)i);
        if (ok0 && ok1 && ok2) break;
      }
      sink[2 * k] = __uint_as_float(r0.x ^ r1.x ^ r2.x);
    }
    __syncthreads();
  }
  if (tid == 0) out[c] = clock64() - t0;
  if (acc == 123.f) out[0] = 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// variant: the slot is fetched with ONE cp.async.bulk (48 line requests) per poll and the tags are checked in smem
__global__ void __launch_bounds__(256, 1)
alltoall_bulk(unsigned long long* buf, int iters, int stk, int nrep, int work, int wide, long long* out) {
  const int tid = threadIdx.x, c = blockIdx.x, NC = gridDim.x;
  const int per = XS / NC;
  __shared__ __align__(128) unsigned long long inbox[XS];
  __shared__ unsigned long long mbar;
  __shared__ int flags[4];
  __shared__ float sink[1024];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = 0;
  float acc = 0.f;
  uint32_t par = 0;
  long long polls = 0;
  for (int i = 1; i <= iters; ++i) {
    if (i == 17) t0 = clock64();
    unsigned long long* slot = buf + (size_t)(i % NSLOT) * MAXREP * XS;
    if (tid < 128) {
      for (int k = 0; k < work; ++k) acc = fmaf(acc, 1.0001f, 0.5f);
      const int w = tid >> 5, lane = tid & 31;
      if (wide) {
        // one 16-byte store carries two tagged entries (per is even)
        for (int rep = lane; rep < nrep; rep += 32)
          for (int e = 2 * w; e < per; e += 8) {
            const uint32_t v = (uint32_t)(c + (acc > 1e30f));
            asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot + (size_t)rep * XS + c * per + e), "r"(v), "r"((uint32_t)i), "r"(v), "r"((uint32_t)i) : "memory");
          }
      } else {
        for (int rep = lane; rep < nrep; rep += 32)
          for (int e = w; e < per; e += 4) put(slot + (size_t)rep * XS + c * per + e, (uint32_t)(c + (acc > 1e30f)), (uint32_t)i, stk);
      }
    } else {
      const unsigned long long* rp = slot + (size_t)(c & (nrep - 1)) * XS;
      const int k = tid - 128;
      for (;;) {
        if (tid == 128) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(XS * 8) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(inbox)), "l"(rp), "r"(XS * 8), "r"(smem_u32(&mbar)) : "memory");
          ++polls;
        }
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(par) : "memory");
        par ^= 1;
        const uint4 r0 = *reinterpret_cast<const uint4*>(inbox + 2 * k);
        const uint4 r1 = *reinterpret_cast<const uint4*>(inbox + 256 + 2 * k);
        const uint4 r2 = *reinterpret_cast<const uint4*>(inbox + 512 + 2 * k);
        const uint32_t ii = (uint32_t)i;
        const bool good = r0.y == ii && r0.w == ii && r1.y == ii && r1.w == ii && r2.y == ii && r2.w == ii;
        const bool wgood = __all_sync(0xffffffffu, good);
        if ((tid & 31) == 0) flags[(tid >> 5) - 4] = wgood;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const bool all = flags[0] && flags[1] && flags[2] && flags[3];
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (all) { sink[2 * k] = __uint_as_float(r0.x ^ r1.x ^ r2.x); break; }
      }
    }
    __syncthreads();
  }
  if (tid == 0) out[c] = clock64() - t0;
  if (tid == 128) out[128 + c] = polls;
  if (acc == 123.f) out[0] = 0;
}

int main() {
  unsigned long long* buf;
  long long* out;
  const size_t nb = (size_t)NSLOT * MAXREP * XS * 8;
  cudaMalloc(&buf, nb);
  cudaMalloc(&out, 256 * 8);
  long long h[256];
  const int iters = 4000;
  printf("== ping-pong (one-way cycles) ==\n");
  for (int partner : {1, 2, 64, 127}) {
    for (int stk = 0; stk < ST_N; ++stk)
      for (int ldk = 0; ldk < LD_N; ++ldk) {
        if (partner != 64 && (stk > 1 || ldk > 0)) continue;
        cudaMemset(buf, 0, nb);
        pingpong<<<128, 32>>>(buf, partner, iters, stk, ldk, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("partner %3d %-18s %-16s one-way %.0f cycles\n", partner, st_names[stk], ld_names[ldk], (double)h[0] / iters / 2);
      }
  }
  printf("== all-to-all, 768 entries, cycles per round (max over CTAs) ==\n");
  auto run_a2a = [&](int nc, int stk, int ldk, int nrep, int npoll, int work, int three) {
    cudaMemset(buf, 0, nb);
    void* args7[] = {&buf, (void*)&iters, &stk, &ldk, &nrep, &npoll, &work, &out};
    void* args6[] = {&buf, (void*)&iters, &stk, &ldk, &nrep, &work, &out};
    cudaError_t e = three ? cudaLaunchCooperativeKernel((void*)alltoall3, dim3(nc), dim3(256), args6, 0, 0)
                          : cudaLaunchCooperativeKernel((void*)alltoall, dim3(nc), dim3(256), args7, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, out, nc * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nc; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("NC %3d %-18s %-16s nrep %d npoll %3d work %4d %s: %.0f cycles/round\n", nc, st_names[stk], ld_names[ldk], nrep,
           three ? 128 : npoll, work, three ? "3-at-once" : "sequential", (double)mx / (iters - 16));
  };
  auto run_bulk = [&](int nc, int stk, int nrep, int work, int wide) {
    cudaMemset(buf, 0, nb);
    void* args[] = {&buf, (void*)&iters, &stk, &nrep, &work, &wide, &out};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)alltoall_bulk, dim3(nc), dim3(256), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, out, 256 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, pl = 0;
    for (int i = 0; i < nc; ++i) { mx = h[i] > mx ? h[i] : mx; pl += h[128 + i]; }
    printf("NC %3d bulk-poll %-18s nrep %3d work %4d wide %d: %.0f cycles/round, %.2f polls/round\n", nc, st_names[stk], nrep, work, wide,
           (double)mx / (iters - 16), (double)pl / nc / iters);
  };
  for (int nrep : {8, 16, 32, 64, 128}) run_a2a(128, ST_RELAXED, LD_RELAXED, nrep, 128, 0, 1);
  for (int nrep : {32, 128}) run_a2a(128, ST_RED, LD_RELAXED, nrep, 128, 0, 1);
  for (int nrep : {8, 32, 128}) run_bulk(128, ST_RELAXED, nrep, 0, 0);
  for (int nrep : {8, 32, 128}) run_bulk(128, ST_RELAXED, nrep, 0, 1);
  for (int nrep : {8, 32, 128}) run_bulk(128, ST_RED, nrep, 0, 0);
  for (int work : {250, 500}) run_bulk(128, ST_RELAXED, 32, work, 1);
  for (int work : {250, 500}) run_a2a(128, ST_RELAXED, LD_RELAXED, 128, 128, work, 1);
  return 0;
}
