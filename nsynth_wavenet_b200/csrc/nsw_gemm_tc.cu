// tcgen05 engine of the conv-GEMM (see nsw_gemm.cuh): split-fp16 operands
// (x = hi + lo, both fp16), fp32 accumulation in TMEM, three tensor-core products
// per k-block (lo*hi + hi*lo + hi*hi; lo*lo ~ 2^-18 relative is dropped).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 3D for X, 2D for W; SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  epilogue: tcgen05.ld (32x32b.x32) -> bias/activation -> global stores
// smem ring: 3 stages x {X_hi, X_lo [128 x 64], W_hi, W_lo [128 x 64]} fp16 = 64 KB/stage.
// TMEM: two accumulator stages of 128 fp32 columns each, so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include "nsw_gemm.cuh"

#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <vector>

namespace nsw {

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 64;
constexpr int STAGES = 3;
constexpr int TC_THREADS = 192;
constexpr uint32_t TILE_BYTES = TBM * TBK * 2;          // 16 KB, one operand plane
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;        // 64 KB
constexpr uint32_t TMEM_COLS = 2 * TBN;                 // 256
constexpr long long WATCHDOG_CYCLES = 4000000000ll;     // ~2 s

struct TcSmemTail {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};
constexpr uint32_t TC_OFF_SCR = STAGES * STAGE_BYTES + 512;  // 4 epilogue warps x 4 KB
static_assert(sizeof(TcSmemTail) <= 512, "barrier block grew");
constexpr size_t TC_SMEM_BYTES = TC_OFF_SCR + 4 * 4096 + 1024;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a wedged pipeline traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > WATCHDOG_CYCLES) {
      printf("nsw conv_gemm_tc: mbarrier watchdog (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (64 fp16 = 128 B per row,
// 8-row groups 1024 B apart); field layout per the sm_100 UMMA descriptor.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);     // start address
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=TBN
__device__ __forceinline__ uint32_t umma_idesc() {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TBN >> 3) << 17) |
         ((uint32_t)(TBM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait (several in flight, then one tmem_ld_wait())
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(saddr)
               : "memory");
  return v;
}

// coalesced read of a [32 rows x 32 floats] block (row pitch `ld` floats) into one row per thread, in two halves so
// that the global loads can be issued long before their values are needed (rows_issue) and transposed through the
// warp's scratch later (rows_finish)
__device__ __forceinline__ void rows_issue(const float* base, size_t row0, int ld, int col, int rows_valid, int lane,
                                           float4 (&t)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    t[i] = r < rows_valid
               ? __ldg(reinterpret_cast<const float4*>(base + (row0 + r) * (size_t)ld + col) + (lane & 7))
               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void rows_finish(const float4 (&t)[8], uint32_t scr, int lane, float (&out)[32]);
__device__ __forceinline__ void load_rows32(const float* base, size_t row0, int ld, int col, int rows_valid,
                                            uint32_t scr, int lane, float (&out)[32]) {
  float4 t[8];
  rows_issue(base, row0, ld, col, rows_valid, lane, t);
  rows_finish(t, scr, lane, out);
}
__device__ __forceinline__ void rows_finish(const float4 (&t)[8], uint32_t scr, int lane, float (&out)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    sts128(scr + (uint32_t)(r * 8 + ((lane & 7) ^ (r & 7))) * 16u, t[i]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 q = lds128(scr + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u);
    out[4 * j] = q.x; out[4 * j + 1] = q.y; out[4 * j + 2] = q.z; out[4 * j + 3] = q.w;
  }
  __syncwarp();
}
// coalesced write of one row per thread: nchunks 16-byte chunks per row (8 = 32 floats, 4 = 32 halves,
// 2 = 16 halves), destination row pitch given in bytes
template <int NCH>
__device__ __forceinline__ void store_rows(unsigned char* base, size_t row0, size_t pitch_bytes,
                                           int rows_valid, uint32_t scr, int lane, const uint4 (&in)[NCH]) {
  constexpr int RPI = 32 / NCH;  // rows per instruction
#pragma unroll
  for (int j = 0; j < NCH; ++j)
    sts128(scr + (uint32_t)(lane * NCH + (j ^ ((lane * NCH / 8) & (NCH - 1)))) * 16u,
           make_float4(__uint_as_float(in[j].x), __uint_as_float(in[j].y), __uint_as_float(in[j].z),
                       __uint_as_float(in[j].w)));
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int r = RPI * i + lane / NCH;
    const int c = lane % NCH;
    const float4 q = lds128(scr + (uint32_t)(r * NCH + (c ^ ((r * NCH / 8) & (NCH - 1)))) * 16u);
    if (r < rows_valid)
      reinterpret_cast<float4*>(base + (row0 + r) * pitch_bytes)[c] = q;
  }
  __syncwarp();
}

__device__ __forceinline__ uint4 pack8_f16(const float* a) {
  uint4 r;
  __half2 p0 = __floats2half2_rn(a[0], a[1]), p1 = __floats2half2_rn(a[2], a[3]);
  __half2 p2 = __floats2half2_rn(a[4], a[5]), p3 = __floats2half2_rn(a[6], a[7]);
  r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
  r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
  return r;
}

// The per-row input an EPI_ROWS / EPI_GATE epilogue reads back from global memory (at most one is prefetched): the
// rows it accumulates onto, else the addend rows.
struct RowSrc {
  const float* base;
  int ld;
};
__device__ __forceinline__ RowSrc primary_rows(const EpiParams& e) {
  if (e.mode == EPI_GATE) return RowSrc{e.addend, e.ld_add};
  if (e.mode == EPI_ROWS) return e.accumulate ? RowSrc{e.out_f32, e.ld_out} : RowSrc{e.addend, e.ld_add};
  return RowSrc{nullptr, 0};
}

// EPI_ROWS / EPI_GATE (see nsw_gemm.cuh) for 32 rows x 32 columns; `pre` = the primary_rows() block of this tile when
// the caller has already loaded it (nullptr: loaded here)
__device__ __forceinline__ void store_tile32_rows(const ConvGemm& g, const EpiParams& e, int clip, int m0,
                                                  int n, const float* v, uint32_t scr, int lane,
                                                  const float* pre = nullptr) {
  const int rows_valid = min(32, g.mclip - m0);
  if (rows_valid <= 0) return;
  const size_t row0 = (size_t)clip * g.mclip + m0;
  float f[32];
  if (e.mode == EPI_GATE) {
    float c[32];
    if (pre) {
#pragma unroll
      for (int j = 0; j < 32; ++j) c[j] = pre[j];
    } else if (e.addend) {
      load_rows32(e.addend, row0, e.ld_add, n, rows_valid, scr, lane, c);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) c[j] = 0.f;
    }
    if (e.bias) {  // per-column bias in the same gate-interleaved order as the columns
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        c[4 * j] += b.x; c[4 * j + 1] += b.y; c[4 * j + 2] += b.z; c[4 * j + 3] += b.w;
      }
    }
    float gh[16], gl[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // sigmoid(s) * tanh(t) = (1 - b) / ((1 + a)(1 + b)), a = e^-s, b = e^-2t: 3 MUFU ops (the flow kernel's gate)
      const float a = __expf(fminf(-(v[2 * j] + c[2 * j]), 40.0f));
      const float b = __expf(fminf(-2.0f * (v[2 * j + 1] + c[2 * j + 1]), 40.0f));
      const float gv = __fdividef(1.0f - b, (1.0f + a) * (1.0f + b));
      gh[j] = __half2float(__float2half_rn(gv));
      gl[j] = gv - gh[j];
    }
    const uint4 hi[2] = {pack8_f16(gh), pack8_f16(gh + 8)};
    const uint4 lo[2] = {pack8_f16(gl), pack8_f16(gl + 8)};
    const size_t pb = (size_t)e.ld_split * 2;
    store_rows<2>(reinterpret_cast<unsigned char*>(e.out_hi + (n >> 1)), row0, pb, rows_valid, scr, lane, hi);
    store_rows<2>(reinterpret_cast<unsigned char*>(e.out_lo + (n >> 1)), row0, pb, rows_valid, scr, lane, lo);
    return;
  }
  // EPI_ROWS
  const float4* b4 = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = __ldg(b4 + j);
    f[4 * j] = v[4 * j] + b.x; f[4 * j + 1] = v[4 * j + 1] + b.y;
    f[4 * j + 2] = v[4 * j + 2] + b.z; f[4 * j + 3] = v[4 * j + 3] + b.w;
  }
  if (e.addend) {
    if (pre && !e.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += pre[j];
    } else {
      float c[32];
      load_rows32(e.addend, row0, e.ld_add, n, rows_valid, scr, lane, c);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += c[j];
    }
  }
  if (e.accumulate) {
    if (pre) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += pre[j];
    } else {
      float c[32];
      load_rows32(e.out_f32, row0, e.ld_out, n, rows_valid, scr, lane, c);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += c[j];
    }
  }
  if (e.relu_out) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (e.out_f32) {
    uint4 w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      w[j] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                        __float_as_uint(f[4 * j + 3]));
    store_rows<8>(reinterpret_cast<unsigned char*>(e.out_f32 + n), row0, (size_t)e.ld_out * 4, rows_valid, scr,
                  lane, w);
  }
  if (e.out_hi) {
    if (n >= e.relu_split_from) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    float h[32], l[32];
    uint32_t rmx = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      range_track(rmx, f[j]);
      h[j] = __half2float(__float2half_rn(f[j]));
      l[j] = f[j] - h[j];
    }
    range_commit(rmx);
    const uint4 hi[4] = {pack8_f16(h), pack8_f16(h + 8), pack8_f16(h + 16), pack8_f16(h + 24)};
    const uint4 lo[4] = {pack8_f16(l), pack8_f16(l + 8), pack8_f16(l + 16), pack8_f16(l + 24)};
    const size_t pb = (size_t)e.ld_split * 2;
    store_rows<4>(reinterpret_cast<unsigned char*>(e.out_hi + n), row0, pb, rows_valid, scr, lane, hi);
    store_rows<4>(reinterpret_cast<unsigned char*>(e.out_lo + n), row0, pb, rows_valid, scr, lane, lo);
  }
}

// EPI_ROWS whose only output is the fp16 split pair (no fp32 rows, no addend: the teacher's residual/skip update), for 32
// rows x 64 columns: every store instruction writes whole 128-byte lines of a plane (4 rows x 128 B) instead of 64-byte
// halves (8 rows x 64 B with the 32-column path).
__device__ __forceinline__ void store_tile64_split(const ConvGemm& g, const EpiParams& e, int clip, int m0, int n,
                                                   const float* v, uint32_t scr, int lane) {
  const int rows_valid = min(32, g.mclip - m0);
  if (rows_valid <= 0 || n >= g.N) return;
  const size_t row0 = (size_t)clip * g.mclip + m0;
  const float4* b4 = reinterpret_cast<const float4*>(e.bias + n);
  const bool relu = n >= e.relu_split_from;
  uint4 hi[8], lo[8];
  uint32_t rmx = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b0 = __ldg(b4 + 2 * j), b1 = __ldg(b4 + 2 * j + 1);
    float f[8] = {v[8 * j] + b0.x, v[8 * j + 1] + b0.y, v[8 * j + 2] + b0.z, v[8 * j + 3] + b0.w,
                  v[8 * j + 4] + b1.x, v[8 * j + 5] + b1.y, v[8 * j + 6] + b1.z, v[8 * j + 7] + b1.w};
    float h[8], l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (relu) f[k] = fmaxf(f[k], 0.f);
      range_track(rmx, f[k]);
      h[k] = __half2float(__float2half_rn(f[k]));
      l[k] = f[k] - h[k];
    }
    hi[j] = pack8_f16(h);
    lo[j] = pack8_f16(l);
  }
  range_commit(rmx);
  const size_t pb = (size_t)e.ld_split * 2;
  store_rows<8>(reinterpret_cast<unsigned char*>(e.out_hi + n), row0, pb, rows_valid, scr, lane, hi);
  store_rows<8>(reinterpret_cast<unsigned char*>(e.out_lo + n), row0, pb, rows_valid, scr, lane, lo);
}

// Epilogue of one warp for 32 rows x 32 consecutive columns (n % 32 == 0).  Thread `lane` holds
// row m0 + lane in v[32].  Values go through a swizzled 4 KB smem scratch so that global memory
// is written with full 128 B lines (8 lanes per row) instead of 32 scattered 16 B pieces per
// instruction (row-per-thread stores made L1TEX the limiter; profiles/r01).
__device__ __forceinline__ void store_tile32(const ConvGemm& g, const EpiParams& e, int clip, int m0,
                                             int n, const float* v, uint32_t scr, int lane,
                                             const float* pre = nullptr) {
  if (n >= g.N) return;  // warp-uniform
  if (e.mode == EPI_ROWS || e.mode == EPI_GATE) {
    store_tile32_rows(g, e, clip, m0, n, v, scr, lane, pre);
    return;
  }
  float f[32];
  int rr = 0, co = n;
  if (e.mode == EPI_PLANES) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      f[4 * j] = v[4 * j] + b.x; f[4 * j + 1] = v[4 * j + 1] + b.y;
      f[4 * j + 2] = v[4 * j + 2] + b.z; f[4 * j + 3] = v[4 * j + 3] + b.w;
    }
  } else {
    rr = n / e.cout;
    co = n - rr * e.cout;
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + co);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      f[4 * j] = apply_act(v[4 * j] + b.x, e.act);
      f[4 * j + 1] = apply_act(v[4 * j + 1] + b.y, e.act);
      f[4 * j + 2] = apply_act(v[4 * j + 2] + b.z, e.act);
      f[4 * j + 3] = apply_act(v[4 * j + 3] + b.w, e.act);
    }
  }
  // destination of row r (0..31) of this warp: pointer to its 32-float chunk, or NULL
  auto row_f32 = [&](int r) -> float4* {
    const int m = m0 + r;
    if (m >= g.mclip || e.out_f32 == nullptr) return nullptr;
    if (e.mode == EPI_PLANES) {
      const size_t M = (size_t)g.nclips * g.mclip;
      return reinterpret_cast<float4*>(e.out_f32 + ((size_t)(n >> 6) * M + (size_t)clip * g.mclip + m) * 64 + (n & 63));
    }
    int c = clip, mm = m;
    if (e.flat_rows > 0) { c = m / e.flat_rows; mm = m - c * e.flat_rows; }  // flattened input (see EpiParams)
    const int o = mm * e.s + rr - e.p;
    if (o < 0 || o >= e.Lout) return nullptr;
    return reinterpret_cast<float4*>(e.out_f32 + ((size_t)c * (e.out_clip_rows > 0 ? e.out_clip_rows : e.Lout) + o) * e.cout + co);
  };
  if (e.mode == EPI_PLANES && (n >> 6) < e.tiled_planes) {
    // row-interleaved plane (tc3 layer kernel): lane's row, float4 j -> [j][lane]; 512 B per store
    const size_t M = (size_t)g.nclips * g.mclip;
    const size_t grow0 = (size_t)clip * g.mclip + m0;  // m0 % 32 == 0, mclip % 128 == 0
    float4* dst = reinterpret_cast<float4*>(e.out_f32 + (size_t)(n >> 6) * M * 64 +
                                            ((grow0 >> 7) * 8 + ((grow0 & 127) >> 5) * 2 + ((n & 63) >> 5)) * 1024);
    if (m0 + lane < g.mclip) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j * 32 + lane] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
    return;
  }
  if (e.out_f32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(scr + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u,
             make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3);
      float4* dst = row_f32(r);
      const float4 val = lds128(scr + (uint32_t)(r * 8 + ((lane & 7) ^ (r & 7))) * 16u);
      if (dst) dst[lane & 7] = val;
    }
    __syncwarp();
  }
  if (e.mode == EPI_DECONV && e.out_hi) {
    {
      uint32_t rmx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) range_track(rmx, f[j]);
      range_commit(rmx);
    }
    // fp16 split planes: 32 values = 64 B per row per plane; both planes share one pass
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = f[8 * j + 2 * k], b = f[8 * j + 2 * k + 1];
          const float ah = __half2float(__float2half_rn(a));
          const float bh = __half2float(__float2half_rn(b));
          __half2 pk = pl == 0 ? __floats2half2_rn(ah, bh) : __floats2half2_rn(a - ah, b - bh);
          w[k] = *reinterpret_cast<uint32_t*>(&pk);
        }
        sts128(scr + (uint32_t)(lane * 4 + (j ^ ((lane >> 1) & 3))) * 16u,
               make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])));
      }
      __syncwarp();
      __half* base = pl == 0 ? e.out_hi : e.out_lo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        const int m = m0 + r;
        int c = clip, mm = m;
        if (e.flat_rows > 0) { c = m / e.flat_rows; mm = m - c * e.flat_rows; }
        const int o = mm * e.s + rr - e.p;
        const float4 val = lds128(scr + (uint32_t)(r * 4 + ((lane & 3) ^ ((r >> 1) & 3))) * 16u);
        if (m < g.mclip && o >= 0 && o < e.Lout)
          reinterpret_cast<float4*>(base + ((size_t)c * (e.out_clip_rows > 0 ? e.out_clip_rows : e.Lout) + o) * e.cout + co)[lane & 3] = val;
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_xh,
                    const __grid_constant__ CUtensorMap map_xl,
                    const __grid_constant__ CUtensorMap map_wh,
                    const __grid_constant__ CUtensorMap map_wl, ConvGemm g, EpiParams e,
                    int tiles_per_clip, int n_tiles, int total_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B operand tiles need 1024 B alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tap = g.cin / TBK;
  const int num_kb = g.ntaps * kb_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tail->tmem_full[a], 1);
      mbar_init(&tail->tmem_empty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tail->tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int clip = mt / tiles_per_clip;
        const int m0 = (mt - clip * tiles_per_clip) * TBM;
        const int n0 = nt * TBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          const int tap = kb / kb_per_tap;
          const int c0 = (kb - tap * kb_per_tap) * TBK;
          const int frame0 = m0 + g.a_off + tap * g.tap_stride;  // may be negative: TMA zero-fills
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          mbar_expect_tx(&tail->full[stage], STAGE_BYTES);
          tma_load_3d(sbase, &map_xh, &tail->full[stage], c0, frame0, clip);
          tma_load_3d(sbase + TILE_BYTES, &map_xl, &tail->full[stage], c0, frame0, clip);
          tma_load_2d(sbase + 2 * TILE_BYTES, &map_wh, &tail->full[stage], kb * TBK, n0);
          tma_load_2d(sbase + 3 * TILE_BYTES, &map_wl, &tail->full[stage], kb * TBK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    const uint32_t idesc = umma_idesc();
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tail->tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * TBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&tail->full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t xh = umma_desc_sw128(sbase);
          const uint64_t xl = umma_desc_sw128(sbase + TILE_BYTES);
          const uint64_t wh = umma_desc_sw128(sbase + 2 * TILE_BYTES);
          const uint64_t wl = umma_desc_sw128(sbase + 3 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < TBK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);  // 32 B per UMMA_K step
            umma_f16(d_tmem, xl + adv, wh + adv, idesc, (kb | k) != 0);
            umma_f16(d_tmem, xh + adv, wl + adv, idesc, 1);
            umma_f16(d_tmem, xh + adv, wh + adv, idesc, 1);
          }
          umma_commit(&tail->empty[stage]);  // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tail->tmem_full[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue ---------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t scr = smem_u32(smem) + TC_OFF_SCR + (uint32_t)(warp - 2) * 4096u;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int clip = mt / tiles_per_clip;
      const int m0 = (mt - clip * tiles_per_clip) * TBM + q * 32;  // first of this warp's 32 rows
      const int n0 = nt * TBN;
      const int as = it & 1;
      mbar_wait(&tail->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TBN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TBN + c0), v);
        store_tile32(g, e, clip, m0, n0 + c0, v, scr, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NSW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    NSW_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, NSW_ECUDA,
              "cuTensorMapEncodeTiled is not available from this driver");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return NSW_OK;
}

int make_map(EncodeTiledFn enc, CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
             const uint64_t* strides_bytes, const uint32_t* box) {
  cuuint64_t gd[3], gs[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSW_CHECK(r == CUDA_SUCCESS, NSW_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NSW_OK;
}


// ---------------------------------------------------------------------------------------------
// Mel-conditioning projection with the activation tile RESIDENT in shared memory.
//
// conv_gemm_tc_kernel re-fetches the [128 x 256] activation tile (hi + lo = 128 KB) for every
// 128-column block of the output, so per output tile it pulls 256 KB through L2 for ~3100 cycles
// of tensor work: it ran at the chip's L2 -> SM limit (~6300 B/clk), not at the tensor limit.  Here
// a CTA loads the activation tile once and streams only the weights (128 KB per 128 columns)
// through a 3-stage ring.  Output: row-interleaved planes (EpiParams::tiled_planes layout) for ALL
// planes, written straight from registers (512 contiguous bytes per warp store).
//   warp 0 TMA producer, warp 1 MMA issuer (uniform control flow, elected lane), warps 2..9 epilogue
// ---------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 320;
constexpr int CP_WSTAGES = 3;
constexpr uint32_t CP_X_BYTES = 8 * TILE_BYTES;       // hi/lo x 4 k-blocks = 128 KB
constexpr uint32_t CP_W_STAGE = 2 * TILE_BYTES;       // W hi + lo [128 n x 64 k] = 32 KB
constexpr uint32_t CP_OFF_W = CP_X_BYTES;
constexpr uint32_t CP_OFF_BARS = CP_OFF_W + CP_WSTAGES * CP_W_STAGE;
constexpr size_t CP_SMEM_BYTES = CP_OFF_BARS + 1024 + 1024;

struct CpBars {
  uint64_t x_full[4], x_free[4];  // per 64-wide k-block of the resident activation tile
  uint64_t w_full[CP_WSTAGES], w_empty[CP_WSTAGES];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

#ifndef NSW_COND_STAGGER
#define NSW_COND_STAGGER 1
#endif

struct CondProj {
  int nclips, mclip, a_off, N;
  int tiles_per_clip, n_tiles;  // n-tiles of 128 columns
  int cluster;                  // 1, or 2: CTA pairs work on two m-tiles in lock step and each CTA fetches half
                                // of every weight stage, multicast into both (halves the L2 reads of W)
  const float* bias;
  float* out;  // [N/64] planes of nclips*mclip rows, row-interleaved
  int exp;     // TIMING EXPERIMENTS ONLY (NSW_COND_EXP): 1 = no global stores, 2 = no MMAs (results are wrong)
  long long* dbg;  // NSW_COND_DEBUG: where the MMA issuer of pair 0 waited (cycles): accumulator, weights, activations, total
};

// Column-tile visiting order of m-tile group `mtg`: rotated by a group-dependent offset, so that the CTAs (which all
// sweep the same 32 weight tiles) are not all asking the same L2 lines for the same tile at the same moment.  A
// bijection per m-tile group, so the static split of the (m-tile, n-tile) item list still covers every pair once.
__device__ __forceinline__ int cp_ntile(int k, int mtg, int n_tiles) {
  if (NSW_COND_STAGGER == 0) return k;
  int nt = k + (mtg * 13) % n_tiles;
  return nt >= n_tiles ? nt - n_tiles : nt;
}

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ bool tc_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(CP_THREADS, 1)
cond_proj_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                    const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                    const __grid_constant__ CUtensorMap map_wh2, const __grid_constant__ CUtensorMap map_wl2,
                    CondProj g) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  CpBars* B = reinterpret_cast<CpBars*>(smem + CP_OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < 4; ++kb) { mbar_init(&B->x_full[kb], 1); mbar_init(&B->x_free[kb], 1); }
    for (int s = 0; s < CP_WSTAGES; ++s) { mbar_init(&B->w_full[s], 1); mbar_init(&B->w_empty[s], (uint32_t)g.cluster); }
    for (int a = 0; a < 2; ++a) { mbar_init(&B->tmem_full[a], 1); mbar_init(&B->tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = B->tmem_base;
  // work items = (m-tile, n-tile) pairs in m-major order; every CTA (or CTA pair: then an "m-tile" of the
  // item list is a PAIR of m-tiles, one per CTA) takes an equal contiguous share and (re)loads the
  // activation tile whenever its share crosses into a new m-tile
  const int CS = g.cluster;
  const int crank = CS == 2 ? (int)(blockIdx.x & 1) : 0;
  const int n_groups = (int)gridDim.x / CS, group = (int)blockIdx.x / CS;
  const long long items = (long long)(g.nclips * g.tiles_per_clip / CS) * g.n_tiles;
  const int i0 = (int)(items * group / n_groups), i1 = (int)(items * (group + 1) / n_groups);
  if (CS == 2) {
    // barriers of both CTAs must be initialised before the peer multicasts into / arrives on them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  if (warp == 0) {
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      int xit = 0, cur_mt = -1;
      for (int i = i0; i < i1; ++i) {
        const int mtg = i / g.n_tiles, nt = cp_ntile(i - mtg * g.n_tiles, mtg, g.n_tiles);
        const int mt = mtg * CS + crank;
        const bool new_x = mt != cur_mt;
        int clip = 0, m0 = 0;
        if (new_x) {
          cur_mt = mt;
          clip = mt / g.tiles_per_clip;
          m0 = (mt - clip * g.tiles_per_clip) * TBM;
          ++xit;
        }
        for (int kb = 0; kb < 4; ++kb) {
          if (new_x) {
            // k-block kb of the activation tile may be replaced as soon as the previous tile's last
            // column block has consumed it: the reload overlaps the tail of the previous tile
            if (xit > 1) mbar_wait(&B->x_free[kb], (uint32_t)((xit - 2) & 1));
            mbar_expect_tx(&B->x_full[kb], 2 * TILE_BYTES);
            tma_load_3d(sbase + (2 * kb) * TILE_BYTES, &map_xh, &B->x_full[kb], kb * TBK, m0 + g.a_off, clip);
            tma_load_3d(sbase + (2 * kb + 1) * TILE_BYTES, &map_xl, &B->x_full[kb], kb * TBK, m0 + g.a_off, clip);
          }
          mbar_wait(&B->w_empty[ws], wphase ^ 1);
          const uint32_t dst = sbase + CP_OFF_W + ws * CP_W_STAGE;
          mbar_expect_tx(&B->w_full[ws], CP_W_STAGE);
          if (CS == 1) {
            tma_load_2d(dst, &map_wh, &B->w_full[ws], kb * TBK, nt * TBN);
            tma_load_2d(dst + TILE_BYTES, &map_wl, &B->w_full[ws], kb * TBK, nt * TBN);
          } else {
            // this CTA fetches 64 of the 128 weight rows of both planes and multicasts them to the pair;
            // the peer's half arrives the same way and completes the same transaction count
            const uint32_t half = (uint32_t)crank * (TILE_BYTES / 2);
            tma_load_2d_mc(dst + half, &map_wh2, &B->w_full[ws], kb * TBK, nt * TBN + crank * 64, 3);
            tma_load_2d_mc(dst + TILE_BYTES + half, &map_wl2, &B->w_full[ws], kb * TBK, nt * TBN + crank * 64, 3);
          }
          if (++ws == CP_WSTAGES) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc();
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
    const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
    int ws = 0;
    uint32_t wphase = 0;
    int xit = 0, cur_mt = -1, acc_it = 0;
    for (int i = i0; i < i1; ++i, ++acc_it) {
      const int mt = i / g.n_tiles;  // m-tile (pair) index: only used to detect a change of activation tile
      const bool first_of_x = mt != cur_mt;
      if (first_of_x) {
        cur_mt = mt;
        ++xit;
      }
      const bool last_of_x = (i + 1 == i1) || ((i + 1) / g.n_tiles != mt);
      const int as = acc_it & 1;
      mbar_wait(&B->tmem_empty[as], (uint32_t)(((acc_it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + (uint32_t)(as * TBN);
      for (int kb = 0; kb < 4; ++kb) {
        if (first_of_x) mbar_wait(&B->x_full[kb], (uint32_t)((xit - 1) & 1));
        mbar_wait(&B->w_full[ws], wphase);
        tc_fence_after();
        const uint64_t xh = umma_desc_sw128(sbase_u + (2 * kb) * TILE_BYTES);
        const uint64_t xl = umma_desc_sw128(sbase_u + (2 * kb + 1) * TILE_BYTES);
        const uint64_t wh = umma_desc_sw128(sbase_u + CP_OFF_W + ws * CP_W_STAGE);
        const uint64_t wl = umma_desc_sw128(sbase_u + CP_OFF_W + ws * CP_W_STAGE + TILE_BYTES);
        if (tc_elect()) {
#pragma unroll
          for (int k = 0; k < TBK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
            if (g.exp == 2 && k > 0) continue;
            umma_f16(d_tmem, xl + adv, wh + adv, idesc, (kb | k) != 0);
            if (g.exp == 2) continue;
            umma_f16(d_tmem, xh + adv, wl + adv, idesc, 1);
            umma_f16(d_tmem, xh + adv, wh + adv, idesc, 1);
          }
          if (CS == 1) umma_commit(&B->w_empty[ws]);
          else umma_commit_mc(&B->w_empty[ws], 3);  // the stage is refilled by BOTH CTAs of the pair
          if (last_of_x) umma_commit(&B->x_free[kb]);
          if (kb == 3) umma_commit(&B->tmem_full[as]);
        }
        __syncwarp();
        if (++ws == CP_WSTAGES) { ws = 0; wphase ^= 1; }
      }
    }
  } else {
    // epilogue: warps w and w+4 share a TMEM lane quarter and take 64 columns (one plane) each
    const int q = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    const size_t M = (size_t)g.nclips * g.mclip;
    int acc_it = 0;
    for (int i = i0; i < i1; ++i, ++acc_it) {
      const int mtg = i / g.n_tiles, nt = cp_ntile(i - mtg * g.n_tiles, mtg, g.n_tiles);
      const int mt = mtg * CS + crank;
      const size_t tile_base = ((size_t)mt * 8 + q * 2) * 1024;  // + h * 1024 + j * 128 + lane * 4
      const int as = acc_it & 1;
      mbar_wait(&B->tmem_full[as], (uint32_t)((acc_it >> 1) & 1));
      tc_fence_after();
      const int n_plane0 = nt * TBN + colhalf * 64;  // first column of this warp's plane
      if (n_plane0 < g.N) {
        float* plane = g.out + (size_t)(n_plane0 >> 6) * M * 64 + tile_base;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TBN + colhalf * 64 + h * 32), v);
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + n_plane0 + h * 32);
          float4* dst = reinterpret_cast<float4*>(plane + h * 1024) + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            const float4 o = make_float4(v[4 * j] + b.x, v[4 * j + 1] + b.y, v[4 * j + 2] + b.z, v[4 * j + 3] + b.w);
            if (g.exp != 1 || o.x == 123456.789f) dst[j * 32] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS == 2) {
    // do not exit while the peer may still multicast into this CTA's shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// The same projection with cta_group::2 MMAs (M = 256 over a CTA pair).  Each CTA of the pair keeps its own
// [128 x 256] activation tile resident and holds only HALF of every weight stage (64 of the 128 weight rows: the
// tensor cores of the two SMs exchange the B operand between themselves), so the per-SM weight stream -- what bounded
// the 1-CTA kernel: 3 stages x 32 KB in flight against ~1 us of L2 latency, tensor pipe 50 % busy, and still 0.29 of
// 0.39 ms with 1/12 of the MMAs (profiles/r02) -- is halved and the ring is twice as deep in the same shared memory.
//   leader (cluster rank 0): issues every tcgen05.mma.cta_group::2; its x_full / w_full barriers count the bytes of
//   BOTH CTAs (the peer's TMA loads name the leader's barrier); commits are multicast to both CTAs; the peer's
//   epilogue warps release an accumulator stage with a remote arrive on the leader's tmem_empty barrier.
// ---------------------------------------------------------------------------------------------
// TN = columns per work item.  TN = 256 is the shape that takes the SM's shared-memory bandwidth out of the picture: an
// SS MMA reads its A and B operands from shared memory at 128 B/clk, and per 16-deep k-step a CTA reads
//   1-CTA, N = 128:  4 KB (A) + 4 KB (B) per  64 tensor cycles = 128 B/clk  (exactly the limit, before the TMA fills)
//   pair,  N = 128:  4 KB     + 2 KB     per  64               =  96 B/clk
//   pair,  N = 256:  4 KB     + 4 KB     per 128               =  64 B/clk
// The two accumulator stages of TN = 256 fill the 512 TMEM columns.
constexpr uint32_t CP2_W_BYTES = 6 * TILE_BYTES;      // weight ring: 96 KB next to the 128 KB activation tile
constexpr int CP2_MAX_STAGES = 6;
constexpr uint32_t CP2_OFF_BARS = CP_OFF_W + CP2_W_BYTES;
constexpr size_t CP2_SMEM_BYTES = CP2_OFF_BARS + 1024 + 1024;

struct Cp2Bars {
  uint64_t x_full[4], x_free[4];
  uint64_t w_full[CP2_MAX_STAGES], w_empty[CP2_MAX_STAGES];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cp2_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cp2_tma_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                           int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void cp2_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cp2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cp2_commit(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void cp2_arrive_remote(uint32_t bar_cluster_addr) {
  // default semantics (release at CTA scope): the accumulator stage this releases lives in TMEM, ordered by
  // tcgen05.fence::before_thread_sync; a cluster-scope release waits for the warp's outstanding global stores
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int TN>
__global__ void __launch_bounds__(CP_THREADS, 1)
cond_proj_tc2_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const __grid_constant__ CUtensorMap map_wh2, const __grid_constant__ CUtensorMap map_wl2, CondProj g) {
  // per CTA and stage: TN/2 weight rows x 64 k, hi plane then lo plane
  constexpr uint32_t W_PLANE = (uint32_t)(TN / 2) * TBK * 2;
  constexpr uint32_t W_STAGE = 2 * W_PLANE;
  constexpr int WSTAGES = (int)(CP2_W_BYTES / W_STAGE);
  constexpr uint32_t TM_COLS = 2 * TN;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  Cp2Bars* B = reinterpret_cast<Cp2Bars*>(smem + CP2_OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g.dbg != nullptr && threadIdx.x == 0) {  // NSW_COND_DEBUG: entry / exit time of every CTA
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g.dbg[8 + 2 * blockIdx.x] = (long long)t;
  }
  const uint32_t sbase = smem_u32(smem);
  const int crank = (int)(blockIdx.x & 1);
  const bool leader = crank == 0;
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < 4; ++kb) { mbar_init(&B->x_full[kb], 1); mbar_init(&B->x_free[kb], 1); }
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(&B->w_full[s], 1); mbar_init(&B->w_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&B->tmem_full[a], 1); mbar_init(&B->tmem_empty[a], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)),
                 "r"(TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  // barriers of both CTAs must exist before the peer's TMA completes on / arrives at them
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = B->tmem_base;
  const int n_groups = (int)gridDim.x / 2, group = (int)blockIdx.x / 2;
  const long long items = (long long)(g.nclips * g.tiles_per_clip / 2) * g.n_tiles;
  const int i0 = (int)(items * group / n_groups), i1 = (int)(items * (group + 1) / n_groups);

  if (warp == 0) {
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      int xit = 0, cur_mt = -1;
      for (int i = i0; i < i1; ++i) {
        const int mtg = i / g.n_tiles, nt = cp_ntile(i - mtg * g.n_tiles, mtg, g.n_tiles);
        const int mt = mtg * 2 + crank;
        const bool new_x = mt != cur_mt;
        int clip = 0, m0 = 0;
        if (new_x) {
          cur_mt = mt;
          clip = mt / g.tiles_per_clip;
          m0 = (mt - clip * g.tiles_per_clip) * TBM;
          ++xit;
        }
        for (int kb = 0; kb < 4; ++kb) {
          if (new_x) {
            if (xit > 1) mbar_wait(&B->x_free[kb], (uint32_t)((xit - 2) & 1));
            if (leader) mbar_expect_tx(&B->x_full[kb], 4 * TILE_BYTES);  // hi + lo of BOTH CTAs
            const uint32_t xb = cp2_mapa(smem_u32(&B->x_full[kb]), 0);
            cp2_tma_3d(sbase + (2 * kb) * TILE_BYTES, &map_xh, xb, kb * TBK, m0 + g.a_off, clip);
            cp2_tma_3d(sbase + (2 * kb + 1) * TILE_BYTES, &map_xl, xb, kb * TBK, m0 + g.a_off, clip);
          }
          mbar_wait(&B->w_empty[ws], wphase ^ 1);
          const uint32_t dst = sbase + CP_OFF_W + ws * W_STAGE;
          if (leader) mbar_expect_tx(&B->w_full[ws], 2 * W_STAGE);
          const uint32_t wb = cp2_mapa(smem_u32(&B->w_full[ws]), 0);
          // this CTA's half of the item's weight rows, hi then lo plane.  The last item of a row pair may have fewer
          // than TN valid columns (704 = 2 x 256 + 192): its MMAs run with N = the valid width, so the pair splits
          // THOSE rows in half (the box still loads TN/2 rows; the extra ones are not read)
          const int nthis = min(TN, g.N - nt * TN);
          cp2_tma_2d(dst, &map_wh2, wb, kb * TBK, nt * TN + crank * (nthis / 2));
          cp2_tma_2d(dst + W_PLANE, &map_wl2, wb, kb * TBK, nt * TN + crank * (nthis / 2));
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // instruction descriptor: D = f32, A = B = f16, K-major, M = 256 (128 per CTA), N = the item's valid columns
      const uint32_t idesc_base = (1u << 4) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
      int ws = 0;
      uint32_t wphase = 0;
      int xit = 0, cur_mt = -1, acc_it = 0;
      const bool dbg = g.dbg != nullptr && blockIdx.x == 0;
      long long w_acc = 0, w_x = 0, w_w = 0;
      const long long t_begin = clock64();
      unsigned long long ns_begin = 0;
      if (dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin));
      for (int i = i0; i < i1; ++i, ++acc_it) {
        const int mt = i / g.n_tiles;
        const bool first_of_x = mt != cur_mt;
        if (first_of_x) {
          cur_mt = mt;
          ++xit;
        }
        const bool last_of_x = (i + 1 == i1) || ((i + 1) / g.n_tiles != mt);
        const int as = acc_it & 1;
        const int mtg_i = i / g.n_tiles;
        const int nthis = min(TN, g.N - cp_ntile(i - mtg_i * g.n_tiles, mtg_i, g.n_tiles) * TN);
        const uint32_t idesc = idesc_base | ((uint32_t)(nthis >> 3) << 17);
        long long tw = dbg ? clock64() : 0;
        mbar_wait(&B->tmem_empty[as], (uint32_t)(((acc_it >> 1) & 1) ^ 1));
        if (dbg) { const long long now = clock64(); w_acc += now - tw; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(as * TN);
        for (int kb = 0; kb < 4; ++kb) {
          tw = dbg ? clock64() : 0;
          if (first_of_x) mbar_wait(&B->x_full[kb], (uint32_t)((xit - 1) & 1));
          if (dbg) { const long long now = clock64(); w_x += now - tw; tw = now; }
          mbar_wait(&B->w_full[ws], wphase);
          if (dbg) { const long long now = clock64(); w_w += now - tw; }
          tc_fence_after();
          const uint64_t xh = umma_desc_sw128(sbase_u + (2 * kb) * TILE_BYTES);
          const uint64_t xl = umma_desc_sw128(sbase_u + (2 * kb + 1) * TILE_BYTES);
          const uint64_t wh = umma_desc_sw128(sbase_u + CP_OFF_W + ws * W_STAGE);
          const uint64_t wl = umma_desc_sw128(sbase_u + CP_OFF_W + ws * W_STAGE + W_PLANE);
          if (tc_elect()) {
#pragma unroll
            for (int k = 0; k < TBK / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              cp2_mma(d_tmem, xl + adv, wh + adv, idesc, (kb | k) != 0);
              cp2_mma(d_tmem, xh + adv, wl + adv, idesc, 1);
              cp2_mma(d_tmem, xh + adv, wh + adv, idesc, 1);
            }
            cp2_commit(&B->w_empty[ws]);
            if (last_of_x) cp2_commit(&B->x_free[kb]);
            if (kb == 3) cp2_commit(&B->tmem_full[as]);
          }
          __syncwarp();
          if (++ws == WSTAGES) { ws = 0; wphase ^= 1; }
        }
      }
      if (dbg && lane == 0) {
        unsigned long long ns_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
        g.dbg[0] = w_acc; g.dbg[1] = w_w; g.dbg[2] = w_x; g.dbg[3] = clock64() - t_begin; g.dbg[4] = i1 - i0;
        g.dbg[5] = (long long)(ns_end - ns_begin);
      }
    }
  } else {
    // epilogue: warps w and w+4 share a TMEM lane quarter and take half of the TN columns (TN/128 planes) each; the
    // accumulator stage is released on the LEADER's barrier
    const int q = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    constexpr int PLANES = TN / 128;  // planes per warp
    const size_t M = (size_t)g.nclips * g.mclip;
    int acc_it = 0;
    for (int i = i0; i < i1; ++i, ++acc_it) {
      const int mtg = i / g.n_tiles, nt = cp_ntile(i - mtg * g.n_tiles, mtg, g.n_tiles);
      const int mt = mtg * 2 + crank;
      const size_t tile_base = ((size_t)mt * 8 + q * 2) * 1024;
      const int as = acc_it & 1;
      mbar_wait(&B->tmem_full[as], (uint32_t)((acc_it >> 1) & 1));
      const long long te0 = (g.dbg != nullptr && blockIdx.x == 0 && warp == 2) ? clock64() : 0;
      tc_fence_after();
      // read this warp's TN/2 columns out in one go and hand the accumulator stage back before the stores: the issuer
      // was waiting on the stage for 30 % of its cycles while the epilogue pushed 128 KB per item through the LSU
      uint32_t acc[2 * PLANES][32];
#pragma unroll
      for (int c = 0; c < 2 * PLANES; ++c)
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TN + colhalf * (TN / 2) + c * 32), acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&B->tmem_empty[as]);
        else cp2_arrive_remote(cp2_mapa(smem_u32(&B->tmem_empty[as]), 0));
      }
#pragma unroll
      for (int pl = 0; pl < PLANES; ++pl) {
        const int col0 = colhalf * (TN / 2) + pl * 64;  // first accumulator column of this plane
        const int n_plane0 = nt * TN + col0;
        if (n_plane0 >= g.N) continue;
        float* plane = g.out + (size_t)(n_plane0 >> 6) * M * 64 + tile_base;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + n_plane0 + h * 32);
          float4* dst = reinterpret_cast<float4*>(plane + h * 1024) + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            dst[j * 32] = make_float4(__uint_as_float(acc[2 * pl + h][4 * j]) + b.x, __uint_as_float(acc[2 * pl + h][4 * j + 1]) + b.y,
                                      __uint_as_float(acc[2 * pl + h][4 * j + 2]) + b.z, __uint_as_float(acc[2 * pl + h][4 * j + 3]) + b.w);
          }
        }
      }
      if (te0 != 0 && lane == 0) g.dbg[6] += clock64() - te0;  // NSW_COND_DEBUG: epilogue busy cycles of one warp
    }
  }

  tc_fence_before();
  __syncthreads();
  // do not exit (or free TMEM) while the peer's MMAs / TMA may still touch this CTA
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
  }
  if (g.dbg != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g.dbg[8 + 2 * blockIdx.x + 1] = (long long)t;
  }
}


// ---------------------------------------------------------------------------------------------
// The general conv-GEMM with cta_group::2 MMAs: M = 256 over a CTA pair (one 128-row m-tile per CTA), N = 256 per
// work item, each CTA streaming its own A rows and HALF of the B rows (the tensor cores of the pair exchange B).
// conv_gemm_tc_kernel above (1 CTA, 128 x 128) reads 4 KB of A and 4 KB of B from shared memory per 64 tensor cycles:
// exactly the 128 B/clk the SM's shared memory delivers, before the TMA fills -- it ran at 48 % of the dense peak on
// the teacher's layers.  Here a k-step is 4 KB + 4 KB per 128 tensor cycles.  Same ring (3 x 64 KB per CTA), same
// epilogues (store_tile32), 2 x 256 accumulator columns = all of TMEM.
//   leader (cluster rank 0) issues; its `full` barriers count the TMA bytes of BOTH CTAs; commits are multicast; the
//   peer's epilogue warps release accumulator stages with a remote arrive on the leader's barrier.
// ---------------------------------------------------------------------------------------------
constexpr int TN2 = 256;
// 8 epilogue warps: warps w and w + 4 share a TMEM lane quarter and take 128 of the 256 columns each.  The epilogues
// with per-row addends (EPI_ROWS, EPI_GATE) are chains of global load -> shared-memory transpose -> math -> transpose ->
// store; with one warp per scheduler they, not the tensor pipe, set the pace (the 128 x 128 kernel and a first version
// of this one with 4 epilogue warps measured the same 13.7 ms on the teacher forward).
constexpr int TC2_THREADS = 64 + 8 * 32;
constexpr size_t TC2_SMEM_BYTES = TC_OFF_SCR + 8 * 4096 + 1024;

// MODE (EpiMode) and SPLIT (ConvGemm::split_acc) are compile-time: with every epilogue inlined behind run-time
// switches the kernel was 52 k SASS instructions and kept the accumulator chunks in local memory.
template <int MODE, int SPLIT>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                     const __grid_constant__ CUtensorMap map_x2h, const __grid_constant__ CUtensorMap map_x2l,
                     const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl,
                     const __grid_constant__ CUtensorMap map_id, ConvGemm g, EpiParams e_in, int tiles_per_clip, int n_tiles,
                     int total_items) {
  EpiParams e = e_in;
  e.mode = MODE;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tap = g.cin / TBK;
  // the k-blocks of the second source come FIRST: tcgen05's fp32 accumulation truncates, an error proportional to the
  // accumulator's magnitude per instruction, so the extra instructions are cheapest while the sum is still small
  const int num_kb2 = g.cin2 / TBK;
  const int num_kbw = num_kb2 + g.ntaps * kb_per_tap;     // k-blocks against the weights
  const int num_kb = num_kbw + (g.acc3 ? TN2 / TBK : 0);  // + the identity k-blocks of ConvGemm::acc3
  const int crank = (int)(blockIdx.x & 1);
  const bool leader = crank == 0;
  const int n_pairs = (int)gridDim.x / 2, pair = (int)blockIdx.x / 2;
  const int total_mt = g.nclips * tiles_per_clip;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tail->tmem_full[a], 1);
      mbar_init(&tail->tmem_empty[a], 16);  // 8 epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "r"(2u * TN2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  // barriers of both CTAs must exist before the peer's TMA completes on / arrives at them
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  const uint32_t sbase0 = smem_u32(smem);

  // item -> (pair of m-tiles, 256-column block); items are dealt round-robin to the pairs, n fastest, so the pairs
  // working at the same time share the A rows through L2
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < total_items; item += n_pairs) {
        const int mp = item / n_tiles, nt = item - mp * n_tiles;
        const int mt = 2 * mp + crank;  // may be one past the end (odd tile count): TMA zero-fills, nothing is stored
        const int clip = mt / tiles_per_clip;
        const int m0 = (mt - clip * tiles_per_clip) * TBM;
        const int nrow = nt * TN2 + crank * (TN2 / 2);  // this CTA's half of the B rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          const uint32_t dst = sbase0 + (uint32_t)stage * STAGE_BYTES;
          if (leader) mbar_expect_tx(&tail->full[stage], kb >= num_kbw ? 2 * (2 * TILE_BYTES + 32 * TBK * 2) : 2 * STAGE_BYTES);
          const uint32_t fb = cp2_mapa(smem_u32(&tail->full[stage]), 0);
          if (kb >= num_kbw) {
            // accumulate source: columns [nt * 256 + j * 64, +64) of Y against the 64 x 64 identity (N = 64 MMAs into
            // that column slice of the accumulator; this CTA holds 32 of the identity's rows: 4 KB per k-block)
            const int j = kb - num_kbw;
            cp2_tma_3d(dst, &map_yh, fb, nt * TN2 + j * TBK, m0, clip);
            cp2_tma_3d(dst + TILE_BYTES, &map_yl, fb, nt * TN2 + j * TBK, m0, clip);
            cp2_tma_2d(dst + 2 * TILE_BYTES, &map_id, fb, 0, crank * 32);
          } else if (kb >= num_kb2) {
            const int tap = (kb - num_kb2) / kb_per_tap;
            const int c0 = (kb - num_kb2 - tap * kb_per_tap) * TBK;
            const int frame0 = m0 + g.a_off + tap * g.tap_stride;  // may be negative: TMA zero-fills
            cp2_tma_3d(dst, &map_xh, fb, c0, frame0, clip);
            cp2_tma_3d(dst + TILE_BYTES, &map_xl, fb, c0, frame0, clip);
          } else {
            cp2_tma_3d(dst, &map_x2h, fb, kb * TBK, m0 + g.a_off2, clip);
            cp2_tma_3d(dst + TILE_BYTES, &map_x2l, fb, kb * TBK, m0 + g.a_off2, clip);
          }
          if (kb < num_kbw) {
            cp2_tma_2d(dst + 2 * TILE_BYTES, &map_wh, fb, kb * TBK, nrow);
            cp2_tma_2d(dst + 3 * TILE_BYTES, &map_wl, fb, kb * TBK, nrow);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // D = f32, A = B = f16, K-major, M = 256 (128 per CTA), N = 256
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      constexpr bool split = SPLIT != 0;  // one accumulator stage: [hi*hi | small products]
      for (int item = pair; item < total_items; item += n_pairs, ++it) {
        const int as = split ? 0 : (it & 1);
        const uint32_t use = split ? (uint32_t)it : (uint32_t)(it >> 1);  // how often this stage has been used before
        mbar_wait(&tail->tmem_empty[as], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(as * TN2);
        const uint32_t d_small = split ? tmem_u + (uint32_t)TN2 : d_tmem;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint32_t sb = sbase_u + (uint32_t)stage * STAGE_BYTES;
          const uint64_t xh = umma_desc_sw128(sb);
          const uint64_t xl = umma_desc_sw128(sb + TILE_BYTES);
          const uint64_t wh = umma_desc_sw128(sb + 2 * TILE_BYTES);
          const uint64_t wl = umma_desc_sw128(sb + 3 * TILE_BYTES);
          if (kb >= num_kbw) {
            // identity k-block: D[:, 64 j .. 64 j + 63] += Y_lo * I + Y_hi * I (exact), M = 256, N = 64
            const uint32_t idesc64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t dj = (uint32_t)((kb - num_kbw) * 64);
            if (tc_elect()) {
#pragma unroll
              for (int k = 0; k < TBK / 16; ++k) {
                const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                cp2_mma(d_small + dj, xl + adv, wh + adv, idesc64, 1);
                cp2_mma(d_tmem + dj, xh + adv, wh + adv, idesc64, 1);
              }
              cp2_commit(&tail->empty[stage]);
              if (kb == num_kb - 1) cp2_commit(&tail->tmem_full[as]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (tc_elect()) {
#pragma unroll
            for (int k = 0; k < TBK / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              cp2_mma(d_small, xl + adv, wh + adv, idesc, (kb | k) != 0);
              cp2_mma(d_small, xh + adv, wl + adv, idesc, 1);
              cp2_mma(d_tmem, xh + adv, wh + adv, idesc, split ? (uint32_t)((kb | k) != 0) : 1u);
            }
            cp2_commit(&tail->empty[stage]);
            if (kb == num_kb - 1) cp2_commit(&tail->tmem_full[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    const uint32_t scr = sbase0 + TC_OFF_SCR + (uint32_t)(warp - 2) * 4096u;
    // The rows an EPI_ROWS / EPI_GATE epilogue reads back (accumulate / addend) are requested one 32-column chunk
    // ahead -- the first chunk of an item before its accumulator is even complete -- so their latency overlaps the
    // MMAs and the previous chunk's math and stores instead of sitting at the head of every chunk.
    const RowSrc ps = primary_rows(e);
    const int cbeg = colhalf * (TN2 / 2), cend = (colhalf + 1) * (TN2 / 2);
    float4 pre[8];
    auto issue = [&](int item, int c0) {  // returns through `pre`; false: nothing to load for this chunk
      const int mp = item / n_tiles, nt = item - mp * n_tiles;
      const int mt = 2 * mp + crank;
      const int clip = mt / tiles_per_clip;
      const int m0 = (mt - clip * tiles_per_clip) * TBM + q * 32;
      const int n = nt * TN2 + c0;
      const int rows_valid = min(32, g.mclip - m0);
      if (ps.base == nullptr || mt >= total_mt || n >= g.N || rows_valid <= 0) return;
      rows_issue(ps.base, (size_t)clip * g.mclip + m0, ps.ld, n, rows_valid, lane, pre);
    };
    int it = 0;
    if (pair < total_items) issue(pair, cbeg);
    for (int item = pair; item < total_items; item += n_pairs, ++it) {
      const int mp = item / n_tiles, nt = item - mp * n_tiles;
      const int mt = 2 * mp + crank;
      const int clip = mt / tiles_per_clip;
      const int m0 = (mt - clip * tiles_per_clip) * TBM + q * 32;
      const int n0 = nt * TN2;
      constexpr bool split = SPLIT != 0;
      const int as = split ? 0 : (it & 1);
      mbar_wait(&tail->tmem_full[as], split ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (split) {
        // one accumulator stage: read this warp's 4 x 32 columns of both accumulators out in one go, hand the stage
        // back to the MMA issuer, and only then do the math and the stores (they overlap the next item's MMAs)
        constexpr int NCH = TN2 / 2 / 32;
        uint32_t acc[NCH][32];  // (raw bits, compile-time indices only: stays in registers)
        const bool rows_live = mt < total_mt && m0 < g.mclip;
        if (rows_live) {
          const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
#pragma unroll
          for (int c = 0; c < NCH; ++c) tmem_ld32_issue(t0 + (uint32_t)(c * 32), acc[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t w[32];
            tmem_ld32_issue(t0 + (uint32_t)(TN2 + c * 32), w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c][j] = __float_as_uint(__uint_as_float(acc[c][j]) + __uint_as_float(w[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tail->tmem_empty[0]);
          else cp2_arrive_remote(cp2_mapa(smem_u32(&tail->tmem_empty[0]), 0));
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int c0 = cbeg + c * 32;
          float pc[32];
          const bool live = rows_live && n0 + c0 < g.N;
          if (live && ps.base) rows_finish(pre, scr, lane, pc);
          if (c + 1 < NCH) issue(item, c0 + 32);
          else if (item + n_pairs < total_items) issue(item + n_pairs, cbeg);
          if (live) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[c][j]);
            store_tile32(g, e, clip, m0, n0 + c0, v, scr, lane, ps.base ? pc : nullptr);
          }
        }
        continue;
      }
      if (MODE == EPI_ROWS && e.out_f32 == nullptr && e.out_hi != nullptr && ps.base == nullptr && e.ld_split % 64 == 0 &&
          (e.relu_split_from % 64 == 0)) {
        // split-only rows: 64 columns at a time, whole 128-byte lines per store (store_tile64_split)
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 64) {
          if (!(mt < total_mt && n0 + c0 < g.N && m0 < g.mclip)) continue;
          float v[64];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TN2 + c0);
          tmem_ld32(ta, v);
          tmem_ld32(ta + 32, v + 32);
          store_tile64_split(g, e, clip, m0, n0 + c0, v, scr, lane);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tail->tmem_empty[as]);
          else cp2_arrive_remote(cp2_mapa(smem_u32(&tail->tmem_empty[as]), 0));
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        float v[32], pc[32];
        const bool live = mt < total_mt && n0 + c0 < g.N && m0 < g.mclip;
        if (live) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TN2 + c0), v);
          if (ps.base) rows_finish(pre, scr, lane, pc);
        }
        if (c0 + 32 < cend) issue(item, c0 + 32);
        else if (item + n_pairs < total_items) issue(item + n_pairs, cbeg);
        if (live) store_tile32(g, e, clip, m0, n0 + c0, v, scr, lane, ps.base ? pc : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tail->tmem_empty[as]);
        else cp2_arrive_remote(cp2_mapa(smem_u32(&tail->tmem_empty[as]), 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  // do not exit (or free TMEM) while the peer's MMAs / TMA may still touch this CTA
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TN2) : "memory");
  }
}

}  // namespace

bool conv_gemm_tc_supported(const ConvGemm& g) {
  return g.cin % TBK == 0 && g.N % 64 == 0 && g.ntaps >= 1;
}

int conv_gemm_tc(const ConvGemm& g, const __half* X_hi, const __half* X_lo,
                 const __half* Bt_hi, const __half* Bt_lo, const EpiParams& e,
                 cudaStream_t stream, const __half* X2_hi, const __half* X2_lo, const __half* Y_hi, const __half* Y_lo) {
  NSW_CHECK(conv_gemm_tc_supported(g), NSW_EINVAL, "conv_gemm_tc: unsupported shape cin=%d N=%d",
            g.cin, g.N);
  NSW_CHECK(X_hi && X_lo && Bt_hi && Bt_lo, NSW_EINVAL, "conv_gemm_tc: null operand");
  if (e.mode == EPI_PLANES && e.tiled_planes > 0)
    NSW_CHECK(g.mclip % 128 == 0 && e.out_f32, NSW_EINVAL, "conv_gemm_tc: tiled planes need mclip %% 128 == 0");
  if (e.mode == EPI_DECONV)
    NSW_CHECK(e.cout % 32 == 0, NSW_EINVAL, "conv_gemm_tc: cout %d must be a multiple of 32", e.cout);
  if (e.mode == EPI_ROWS)
    NSW_CHECK(e.ld_out % 4 == 0 && e.ld_add % 4 == 0 && e.ld_split % 8 == 0 && e.bias, NSW_EINVAL,
              "conv_gemm_tc: EPI_ROWS leading dimensions must keep 16-byte alignment");
  if (e.mode == EPI_GATE)
    NSW_CHECK((e.addend || e.bias) && e.out_hi && e.out_lo && e.ld_add % 4 == 0 && e.ld_split % 8 == 0, NSW_EINVAL,
              "conv_gemm_tc: EPI_GATE needs cond rows or a bias, and split outputs");
  const bool second = g.cin2 > 0;
  if (second)
    NSW_CHECK(X2_hi && X2_lo && g.cin2 % TBK == 0 && g.L2 > 0, NSW_EINVAL, "conv_gemm_tc: bad second source (cin2=%d)", g.cin2);
  EncodeTiledFn enc;
  NSW_TRY(get_encode_fn(&enc));
  const int K = g.ntaps * g.cin + g.cin2;
  CUtensorMap mxh, mxl, mwh, mwl, mx2h, mx2l, myh, myl, mid;
  if (g.acc3) {
    NSW_CHECK(Y_hi && Y_lo && g.ld3 % 8 == 0 && g.ld3 >= g.N, NSW_EINVAL, "conv_gemm_tc: bad accumulate source (ld3=%d)", g.ld3);
    // 64 x 64 fp16 identity, one per device, built on first use
    static __half* ident[64] = {nullptr};
    static std::mutex ident_mu;
    int dv = 0;
    NSW_CUDA(cudaGetDevice(&dv));
    std::lock_guard<std::mutex> ident_lock(ident_mu);
    if (!ident[dv & 63]) {
      std::vector<__half> id((size_t)64 * 64, __float2half(0.f));
      for (int i = 0; i < 64; ++i) id[(size_t)i * 64 + i] = __float2half(1.f);
      __half* d = nullptr;
      NSW_CUDA(cudaMalloc(&d, id.size() * sizeof(__half)));
      NSW_CUDA(cudaMemcpy(d, id.data(), id.size() * sizeof(__half), cudaMemcpyHostToDevice));
      ident[dv & 63] = d;
    }
    {
      const uint64_t dims[2] = {64, 64};
      const uint64_t strides[1] = {64 * 2};
      const uint32_t box[2] = {TBK, 32};
      NSW_TRY(make_map(enc, &mid, ident[dv & 63], 2, dims, strides, box));
    }
    const uint64_t dims[3] = {(uint64_t)g.ld3, (uint64_t)g.mclip, (uint64_t)g.nclips};
    const uint64_t strides[2] = {(uint64_t)g.ld3 * 2, (uint64_t)g.mclip * g.ld3 * 2};
    const uint32_t box[3] = {TBK, TBM, 1};
    NSW_TRY(make_map(enc, &myh, Y_hi, 3, dims, strides, box));
    NSW_TRY(make_map(enc, &myl, Y_lo, 3, dims, strides, box));
  }
  if (second) {
    const uint64_t dims[3] = {(uint64_t)g.cin2, (uint64_t)g.L2, (uint64_t)g.nclips};
    const uint64_t strides[2] = {(uint64_t)g.cin2 * 2, (uint64_t)g.L2 * g.cin2 * 2};
    const uint32_t box[3] = {TBK, TBM, 1};
    NSW_TRY(make_map(enc, &mx2h, X2_hi, 3, dims, strides, box));
    NSW_TRY(make_map(enc, &mx2l, X2_lo, 3, dims, strides, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.cin, (uint64_t)g.L, (uint64_t)g.nclips};
    const uint64_t pitch = (uint64_t)(g.x_pitch > 0 ? g.x_pitch : g.cin);
    const uint64_t strides[2] = {pitch * 2, (uint64_t)g.L * pitch * 2};
    const uint32_t box[3] = {TBK, TBM, 1};
    NSW_TRY(make_map(enc, &mxh, X_hi, 3, dims, strides, box));
    NSW_TRY(make_map(enc, &mxl, X_lo, 3, dims, strides, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)g.N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {TBK, TBN};
    NSW_TRY(make_map(enc, &mwh, Bt_hi, 2, dims, strides, box));
    NSW_TRY(make_map(enc, &mwl, Bt_lo, 2, dims, strides, box));
  }
  static std::atomic<uint64_t> attr_done{0};  // per device (the attribute is per device)
  NSW_TRY(ensure_dynamic_smem((const void*)conv_gemm_tc_kernel, (int)TC_SMEM_BYTES, attr_done));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_per_clip = (g.mclip + TBM - 1) / TBM;
  // CTA-pair kernel (M = 256, N = 256 per item) whenever there are more than 128 columns and at least two m-tiles;
  // NSW_GEMM_1CTA=1 keeps the 128 x 128 kernel (A/B runs)
  static const bool force_1cta = getenv("NSW_GEMM_1CTA") != nullptr;
  const int total_mt = g.nclips * tiles_per_clip;
  const bool pair_ok = g.N > TBN && total_mt >= 2 && sms >= 2;
  NSW_CHECK(!(second || g.acc3) || pair_ok, NSW_EINVAL,
            "conv_gemm_tc: a second / accumulate source needs the pair kernel (N > 128, >= 2 m-tiles)");
  if (!second) { mx2h = mxh; mx2l = mxl; }
  if (!g.acc3) { myh = mxh; myl = mxl; mid = mwh; }
  NSW_CHECK(!(g.one_cta && (second || g.acc3 || g.split_acc)), NSW_EINVAL,
            "conv_gemm_tc: the one-CTA kernel has no second / accumulate source and no split accumulators");
  if (!g.one_cta && (!force_1cta || second || g.acc3) && pair_ok) {
    typedef void (*Tc2Fn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                          CUtensorMap, CUtensorMap, ConvGemm, EpiParams, int, int, int);
    static const Tc2Fn fns[4][2] = {{conv_gemm_tc2_kernel<EPI_PLANES, 0>, conv_gemm_tc2_kernel<EPI_PLANES, 1>},
                                    {conv_gemm_tc2_kernel<EPI_DECONV, 0>, conv_gemm_tc2_kernel<EPI_DECONV, 1>},
                                    {conv_gemm_tc2_kernel<EPI_ROWS, 0>, conv_gemm_tc2_kernel<EPI_ROWS, 1>},
                                    {conv_gemm_tc2_kernel<EPI_GATE, 0>, conv_gemm_tc2_kernel<EPI_GATE, 1>}};
    NSW_CHECK(e.mode >= 0 && e.mode < 4, NSW_EINVAL, "conv_gemm_tc: bad epilogue mode %d", e.mode);
    const int sp = g.split_acc ? 1 : 0;
    const Tc2Fn kern = fns[e.mode][sp];
    static std::atomic<uint64_t> attr2_done[4][2];
    NSW_TRY(ensure_dynamic_smem((const void*)kern, (int)TC2_SMEM_BYTES, attr2_done[e.mode][sp]));
    const int n_tiles2 = (g.N + TN2 - 1) / TN2;
    const int items = ((total_mt + 1) / 2) * n_tiles2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1);
    cfg.blockDim = dim3(TC2_THREADS);
    cfg.dynamicSmemBytes = TC2_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs[64] = {0};  // per device
    if (max_pairs[dev & 63] == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = sms / 2;
      max_pairs[dev & 63] = n;
    }
    cfg.gridDim = dim3(2 * (unsigned)std::min(items, std::min(max_pairs[dev & 63], sms / 2)));
    NSW_CUDA(cudaLaunchKernelEx(&cfg, kern, mxh, mxl, mwh, mwl, mx2h, mx2l, myh, myl, mid, g, e, tiles_per_clip, n_tiles2,
                                items));
    count_launch();
    NSW_CUDA(cudaGetLastError());
    return NSW_OK;
  }
  const int n_tiles = (g.N + TBN - 1) / TBN;
  const int total = g.nclips * tiles_per_clip * n_tiles;
  const int grid = std::min(total, sms);
  conv_gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(mxh, mxl, mwh, mwl, g, e,
                                                                  tiles_per_clip, n_tiles, total);
  count_launch();
  NSW_CUDA(cudaGetLastError());
  return NSW_OK;
}

// X-resident mel-conditioning projection: out planes (ALL row-interleaved) = X[clip, m + a_off, :] . W + bias.
// X_hi/X_lo [nclips, L, 256]; Bt_hi/Bt_lo [N, 256]; N % 64 == 0; mclip % 128 == 0.
int cond_proj_tc(int nclips, int L, int mclip, int a_off, int N, const __half* X_hi, const __half* X_lo,
                 const __half* Bt_hi, const __half* Bt_lo, const float* bias, float* out_tiled,
                 cudaStream_t stream) {
  NSW_CHECK(mclip % TBM == 0 && N % 64 == 0 && nclips >= 1, NSW_EINVAL,
            "cond_proj_tc: needs mclip %% 128 == 0 and N %% 64 == 0 (got %d, %d)", mclip, N);
  EncodeTiledFn enc;
  NSW_TRY(get_encode_fn(&enc));
  const int K = 4 * TBK;
  CUtensorMap mxh, mxl, mwh, mwl;
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)L, (uint64_t)nclips};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)L * K * 2};
    const uint32_t box[3] = {TBK, TBM, 1};
    NSW_TRY(make_map(enc, &mxh, X_hi, 3, dims, strides, box));
    NSW_TRY(make_map(enc, &mxl, X_lo, 3, dims, strides, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {TBK, TBN};
    NSW_TRY(make_map(enc, &mwh, Bt_hi, 2, dims, strides, box));
    NSW_TRY(make_map(enc, &mwl, Bt_lo, 2, dims, strides, box));
  }
  CUtensorMap mwh2, mwl2;  // half-height weight boxes for the CTA-pair multicast
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {TBK, TBN / 2};
    NSW_TRY(make_map(enc, &mwh2, Bt_hi, 2, dims, strides, box));
    NSW_TRY(make_map(enc, &mwl2, Bt_lo, 2, dims, strides, box));
  }
  static std::atomic<uint64_t> attr_done{0};  // per device (the attribute is per device)
  NSW_TRY(ensure_dynamic_smem((const void*)cond_proj_tc_kernel, (int)CP_SMEM_BYTES, attr_done));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CondProj g;
  g.nclips = nclips;
  g.mclip = mclip;
  g.a_off = a_off;
  g.N = N;
  g.tiles_per_clip = mclip / TBM;
  g.n_tiles = (N + TBN - 1) / TBN;
  const long long items = (long long)nclips * g.tiles_per_clip * g.n_tiles;
  g.bias = bias;
  g.out = out_tiled;
  g.exp = getenv("NSW_COND_EXP") ? atoi(getenv("NSW_COND_EXP")) : 0;
  g.dbg = nullptr;
  static long long* dbg_buf = nullptr;  // NSW_COND_DEBUG only (single device)
  const bool want_dbg = getenv("NSW_COND_DEBUG") != nullptr;
  if (want_dbg) {
    if (!dbg_buf) NSW_CUDA(cudaMalloc(&dbg_buf, 512 * sizeof(long long)));
    NSW_CUDA(cudaMemsetAsync(dbg_buf, 0, 512 * sizeof(long long), stream));
    g.dbg = dbg_buf;
  }
  const bool pair = getenv("NSW_COND_NOCLUSTER") == nullptr && (nclips * g.tiles_per_clip) % 2 == 0 && sms >= 2;
  g.cluster = pair ? 2 : 1;
  // cta_group::2 MMAs (cond_proj_tc2_kernel) by default; NSW_COND_1CTA=1 keeps the pair kernel with 1-CTA MMAs
  const bool two = pair && getenv("NSW_COND_1CTA") == nullptr && g.exp == 0;
  if (two) {
    // 256-column work items unless NSW_COND_TN=128 (the earlier shape, kept for the A/B runs under scripts/)
    static const bool wide = !(getenv("NSW_COND_TN") && atoi(getenv("NSW_COND_TN")) == 128);
    auto kern = wide ? cond_proj_tc2_kernel<256> : cond_proj_tc2_kernel<128>;
    static std::atomic<uint64_t> attr2_done{0}, attr2w_done{0};
    NSW_TRY(ensure_dynamic_smem((const void*)kern, (int)CP2_SMEM_BYTES, wide ? attr2w_done : attr2_done));
    if (wide) g.n_tiles = (N + 255) / 256;
    const long long pair_items = (long long)nclips * g.tiles_per_clip * g.n_tiles / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1);
    cfg.blockDim = dim3(CP_THREADS);
    cfg.dynamicSmemBytes = CP2_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs2[64][2] = {{0}};  // per device and kernel variant: the query costs tens of microseconds
    int& n = max_pairs2[dev & 63][wide ? 1 : 0];
    if (n == 0 && (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1)) n = sms / 2;
    cfg.gridDim = dim3(2 * (unsigned)std::min<long long>(pair_items, std::min(n, sms / 2)));
    // each CTA loads TN/2 weight rows per stage: the full-height box for TN = 256, the half-height one for 128
    if (wide) NSW_CUDA(cudaLaunchKernelEx(&cfg, kern, mxh, mxl, mwh, mwl, g));
    else NSW_CUDA(cudaLaunchKernelEx(&cfg, kern, mxh, mxl, mwh2, mwl2, g));
  } else if (pair) {
    const long long pair_items = items / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1);
    cfg.blockDim = dim3(CP_THREADS);
    cfg.dynamicSmemBytes = CP_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // the shares are static, so every pair must be resident at once: GPCs with an odd SM count cannot
    // pair their last SM
    static int max_pairs = -1;
    if (max_pairs < 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, cond_proj_tc_kernel, &cfg) != cudaSuccess || n < 1) n = sms / 2;
      max_pairs = n;
    }
    cfg.gridDim = dim3(2 * (unsigned)std::min<long long>(pair_items, std::min(max_pairs, sms / 2)));
    NSW_CUDA(cudaLaunchKernelEx(&cfg, cond_proj_tc_kernel, mxh, mxl, mwh, mwl, mwh2, mwl2, g));
  } else {
    const int grid = (int)std::min<long long>(items, sms);
    cond_proj_tc_kernel<<<grid, CP_THREADS, CP_SMEM_BYTES, stream>>>(mxh, mxl, mwh, mwl, mwh2, mwl2, g);
  }
  count_launch();
  NSW_CUDA(cudaGetLastError());
  if (want_dbg) {
    long long hb[512];
    NSW_CUDA(cudaStreamSynchronize(stream));
    NSW_CUDA(cudaMemcpy(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost));
    long long t0 = 0, t1 = 0, s_last = 0, e_first = 0;
    int n_cta = 0;
    for (int b = 0; b < 250 && hb[8 + 2 * b] != 0; ++b, ++n_cta) {
      const long long a = hb[8 + 2 * b], z = hb[8 + 2 * b + 1];
      if (n_cta == 0 || a < t0) t0 = a;
      if (n_cta == 0 || a > s_last) s_last = a;
      if (n_cta == 0 || z > t1) t1 = z;
      if (n_cta == 0 || z < e_first) e_first = z;
    }
    fprintf(stderr, "[cond_proj dbg] %d CTAs: first entry -> last exit %lld ns; last entry +%lld ns, first exit +%lld ns\n", n_cta,
            t1 - t0, s_last - t0, e_first - t0);
    fprintf(stderr, "[cond_proj dbg N=%d] pair 0: %lld items in %lld cycles = %lld ns (%.0f MHz); issuer waited: accumulator %lld, "
                    "weights %lld, activations %lld; one epilogue warp busy %lld\n", N, hb[4], hb[3], hb[5],
            hb[5] > 0 ? 1e3 * (double)hb[3] / (double)hb[5] : 0.0, hb[0], hb[1], hb[2], hb[6]);
  }
  return NSW_OK;
}

}  // namespace nsw


// ---------------------------------------------------------------------------------------------
// nsw_conv_gemm_device (include/nsw.h): kernel-level parity hook.  fp32 operands are split / transposed on the
// device into scratch buffers that live for the call only.
// ---------------------------------------------------------------------------------------------
namespace nsw {
namespace {
__global__ void hook_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}
// w [K rows: taps*cin then cin2][N]  ->  Bt [N][K columns: cin2 first, then taps*cin], split
__global__ void hook_weight_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int Kmain,
                                   int K2, int N) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Kmain + K2;
  if (i >= (size_t)N * K) return;
  const int n = (int)(i / K), kc = (int)(i - (size_t)n * K);
  const int krow = kc < K2 ? Kmain + kc : kc - K2;
  const float v = w[(size_t)krow * N + n];
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}
}  // namespace
}  // namespace nsw

extern "C" int nsw_conv_gemm_device(const float* d_x, int32_t nclips, int32_t L, int32_t cin, int32_t ntaps, int32_t a_off,
                                    int32_t tap_stride, int32_t mclip, const float* d_w, int32_t N, const float* d_bias,
                                    const float* d_x2, int32_t L2, int32_t cin2, int32_t a_off2, const float* d_y,
                                    int32_t flags, float* d_out, void* stream) {
  using namespace nsw;
  NSW_CHECK(d_x && d_w && d_out && nclips >= 1 && L >= 1 && mclip >= 1 && ntaps >= 1, NSW_EINVAL,
            "nsw_conv_gemm_device: bad argument");
  NSW_CHECK(cin % 64 == 0 && N % 64 == 0 && (d_x2 == nullptr || (cin2 % 64 == 0 && cin2 > 0 && L2 >= 1)), NSW_EINVAL,
            "nsw_conv_gemm_device: cin, cin2 and N must be multiples of 64");
  cudaStream_t st = (cudaStream_t)stream;
  const int K2 = d_x2 ? cin2 : 0, Kmain = ntaps * cin, K = Kmain + K2;
  const size_t nx = (size_t)nclips * L * cin, nx2 = d_x2 ? (size_t)nclips * L2 * cin2 : 0, nw = (size_t)N * K;
  const size_t rows = (size_t)nclips * mclip, ny = d_y ? rows * N : 0;
  DevBuf xs, x2s, ws, ys, zero_bias;
  NSW_TRY(xs.ensure(nx * 2 * sizeof(__half)));
  NSW_TRY(ws.ensure(nw * 2 * sizeof(__half)));
  if (nx2) NSW_TRY(x2s.ensure(nx2 * 2 * sizeof(__half)));
  if (ny) NSW_TRY(ys.ensure(ny * 2 * sizeof(__half)));
  auto split = [&](const float* src, DevBuf& dst, size_t n) {
    hook_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst.as<__half>(), dst.as<__half>() + n, n);
  };
  split(d_x, xs, nx);
  if (nx2) split(d_x2, x2s, nx2);
  if (ny) split(d_y, ys, ny);
  hook_weight_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(d_w, ws.as<__half>(), ws.as<__half>() + nw, Kmain, K2, N);
  NSW_CUDA(cudaGetLastError());
  if (!d_bias) {
    NSW_TRY(zero_bias.ensure((size_t)N * sizeof(float)));
    NSW_CUDA(cudaMemsetAsync(zero_bias.p, 0, (size_t)N * sizeof(float), st));
  }
  ConvGemm g;
  g.nclips = nclips; g.L = L; g.cin = cin; g.ntaps = ntaps; g.a_off = a_off; g.tap_stride = tap_stride; g.mclip = mclip;
  g.N = N;
  if (d_x2) { g.cin2 = cin2; g.a_off2 = a_off2; g.L2 = L2; }
  if (d_y) { g.acc3 = 1; g.ld3 = N; }
  g.split_acc = (flags & 1) ? 1 : 0;
  g.one_cta = (flags & 2) ? 1 : 0;
  EpiParams e{};
  e.mode = EPI_ROWS;
  e.bias = d_bias ? d_bias : zero_bias.as<float>();
  e.out_f32 = d_out;
  e.ld_out = N;
  int rc = conv_gemm_tc(g, xs.as<__half>(), xs.as<__half>() + nx, ws.as<__half>(), ws.as<__half>() + nw, e, st,
                        nx2 ? x2s.as<__half>() : nullptr, nx2 ? x2s.as<__half>() + nx2 : nullptr,
                        ny ? ys.as<__half>() : nullptr, ny ? ys.as<__half>() + ny : nullptr);
  cudaStreamSynchronize(st);  // the scratch buffers die with this frame
  return rc;
}

NSW_RANGE_GUARD_TU(gemm_tc)
