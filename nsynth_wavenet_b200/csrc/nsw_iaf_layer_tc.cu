// tcgen05 version of the fused IAF residual layer (parallel_wavenet.py:227-254):
//   D1[128 x 64] = sum_tap L[t-(2-tap)d][128 x 64] . Wd_tap[64 x 64]      (split fp16, fp32 acc)
//   g = sigmoid(D1[:, even] + cond[:, even]) * tanh(D1[:, odd] + cond[:, odd])   (gate-interleaved)
//   D2[128 x 64] = g[128 x 32] . Wr[32 x 64]                                (A operand from TMEM)
//   l_new = l + br + D2  ->  fp32 stream + fp16 hi/lo planes for the next layer's MMA operand
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer: weights once, then per 128-row tile the six [128 x 64] fp16
//               operand tiles (3 taps x {lo, hi}) through an 8-stage ring (SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; MMA2 of tile i is queued
//               behind MMA1 of tile i+1 so the tensor pipe never waits for the gate epilogue
//   warps 2-5   epilogue warpgroup 0 (even local tiles), warps 6-9 warpgroup 1 (odd tiles):
//               tcgen05.ld D1 -> gate -> tcgen05.st g (fp16 hi|lo) -> ... -> tcgen05.ld D2 ->
//               residual add -> global stores
// TMEM (512 columns): D1[2] 0..127, D2[2] 128..255, G[2] 256..319 (16 cols hi + 16 cols lo each).
#include "nsw_gemm.cuh"

#include <cuda.h>

#include <cstdlib>

namespace nsw {

namespace {

constexpr int C = 64;
constexpr int LT_BM = 128;
constexpr int LT_STAGES = 4;   // ring for the two PAST taps (t-2d, t-d): 2 planes x 2 taps per tile
constexpr int LT_THREADS = 320;
constexpr uint32_t A_TILE_BYTES = LT_BM * 64 * 2;   // 16 KB
constexpr uint32_t WD_TILE_BYTES = 64 * 64 * 2;     // 8 KB per tap per plane
constexpr uint32_t WR_TILE_BYTES = 64 * 32 * 2;     // 4 KB per plane
constexpr uint32_t W_BYTES = 6 * WD_TILE_BYTES + 2 * WR_TILE_BYTES;  // 56 KB
constexpr long long LT_WATCHDOG = 4000000000ll;

struct LtBars {
  uint64_t wfull, wfree;
  uint64_t full[LT_STAGES];
  uint64_t empty[LT_STAGES];
  uint64_t cur_full[2], cur_free[2];
  uint64_t d1_full[2], d1_empty[2], g_full[2], d2_full[2], d2_empty[2];
  uint32_t tmem_base;
};
// smem: [Wd_hi 3x8K][Wd_lo 3x8K][Wr_hi 4K][Wr_lo 4K][past-tap ring 4x16K][current-tap tiles 2x2x16K][bars][scratch]
// The current-tap (t) operand tiles live in their own double buffer because they are read twice:
// by the MMA and, as the residual input l[t] = hi + lo, by the epilogue.
constexpr uint32_t OFF_WDH = 0, OFF_WDL = 3 * WD_TILE_BYTES, OFF_WRH = 6 * WD_TILE_BYTES,
                   OFF_WRL = OFF_WRH + WR_TILE_BYTES, OFF_A = OFF_WRL + WR_TILE_BYTES,
                   OFF_CUR = OFF_A + LT_STAGES * A_TILE_BYTES,
                   OFF_BARS = OFF_CUR + 4 * A_TILE_BYTES;
constexpr uint32_t OFF_SCR = OFF_BARS + 512;                 // 8 warps x 4 KB transposition scratch
constexpr uint32_t SCR_BYTES = 8 * 4096;
static_assert(sizeof(LtBars) <= 512, "barrier block grew");
constexpr size_t LT_SMEM_BYTES = OFF_SCR + SCR_BYTES + 1024;

__device__ __forceinline__ void lt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lt_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void lt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lt_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins == 2048) {
      spins = 0;
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > LT_WATCHDOG) {
        printf("nsw iaf_layer_tc: mbarrier watchdog (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void lt_tma_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                          int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void lt_tma_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                          int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major smem matrix descriptors (sm_100 UMMA): 128 B rows / SWIZZLE_128B, 64 B rows / SWIZZLE_64B
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// D=f32, A=B=f16, K-major, M=128, N=64
__device__ __forceinline__ uint32_t lt_idesc() {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc,
                                       uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void lt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// sigmoid(s) * tanh(t) = (1 - b) / ((1 + a)(1 + b)), a = e^-s, b = e^-2t : 3 MUFU ops.
// Exponents are clamped at 40 so the product stays finite; both functions are saturated to
// fp32 precision long before that.
__device__ __forceinline__ float gate_fast(float s, float t) {
  const float a = __expf(fminf(-s, 40.0f));
  const float b = __expf(fminf(-2.0f * t, 40.0f));
  return __fdividef(1.0f - b, (1.0f + a) * (1.0f + b));
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);  // .x (low 16 bits) = a
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- coalesced global access for row-per-thread data --------------------------------------
// A warp owns 32 rows x 128 B (its half of the 256 B fp32 rows).  Global memory is touched with
// fully used 128 B lines (lane -> row 4i + lane/8, 16 B chunk lane%8); a 4 KB per-warp smem
// scratch with an XOR swizzle (chunk ^ row%8) turns that into one full half-row per thread and
// back, conflict-free in both directions.  (Row-per-thread LDG/STG costs 32 L1TEX sectors per
// instruction and made L1TEX the limiter: profiles/r01.)
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
// scr: 32-bit shared-space address of this warp's 4 KB scratch
__device__ __forceinline__ void rows_issue128(const float4* __restrict__ g /*row 0 of this warp, this half*/,
                                              int lane, float4 (&t)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = __ldg(g + (size_t)(4 * i + (lane >> 3)) * 16 + (lane & 7));
}
__device__ __forceinline__ void rows_transpose128(uint32_t scr, int lane, const float4 (&t)[8],
                                                  float4 (&out)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    sts128(scr + (uint32_t)(r * 8 + ((lane & 7) ^ (r & 7))) * 16u, t[i]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = lds128(scr + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u);
  __syncwarp();
}
__device__ __forceinline__ void rows_store128(float4* __restrict__ g, uint32_t scr, int lane,
                                              const float4 (&in)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) sts128(scr + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u, in[j]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    g[(size_t)r * 16 + (lane & 7)] = lds128(scr + (uint32_t)(r * 8 + ((lane & 7) ^ (r & 7))) * 16u);
  }
  __syncwarp();
}
// 32 rows x 64 B (half of a 128 B fp16 row): lane -> row 8i + lane/4, chunk lane%4
__device__ __forceinline__ void rows_store64(uint4* __restrict__ g /*row pitch 8 uint4*/, uint32_t scr,
                                             int lane, const uint4 (&in)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = in[j];
    sts128(scr + (uint32_t)(lane * 4 + (j ^ ((lane >> 1) & 3))) * 16u,
           make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)));
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    const float4 v = lds128(scr + (uint32_t)(r * 4 + ((lane & 3) ^ ((r >> 1) & 3))) * 16u);
    g[(size_t)r * 8 + (lane & 3)] =
        make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
  }
  __syncwarp();
}

struct LayerTcParams {
  const float* cond;       // plane of layer 0 of this flow: [rows][64] gate-interleaved, biases folded
  size_t cond_plane;       // floats between consecutive layers' planes
  __half* hi[2];           // ping-pong residual stream (fp16 hi / lo planes, [rows][64] each);
  __half* lo[2];           //   layer l0 reads buffer `buf0`, writes the other, and so on
  const float* br;         // [L][64] natural order
  int buf0, l0, l1, num_stages;
  int tiles_per_clip, n_tiles;
  unsigned int* grid_counter;  // arrival counter of the inter-layer grid barrier (zeroed per launch)
  unsigned int grid_base;
  long long* dbg;  // optional timeline of CTA 0 (NULL = off): [0..31] MMA warp, [32..95] epilogue warp 2
};

// Persistent over the layers [l0, l1) of one flow: a layer needs rows t-d, t-2d written by OTHER
// CTAs in the previous layer, so layers are separated by a grid-wide barrier (one arrival per CTA
// on a global counter, polled by each CTA's TMA producer).  Barrier state, TMEM and the smem
// pipelines simply keep running across layers; the next layer's weights are fetched as soon as
// the MMA warp has retired the current layer, i.e. while the last epilogues and the barrier run.
__global__ void __launch_bounds__(LT_THREADS, 1)
iaf_layer_tc_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_l0,
                    const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_l1,
                    const __grid_constant__ CUtensorMap map_wdh, const __grid_constant__ CUtensorMap map_wdl,
                    const __grid_constant__ CUtensorMap map_wrh, const __grid_constant__ CUtensorMap map_wrl,
                    LayerTcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  LtBars* B = reinterpret_cast<LtBars*>(smem + OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);

  if (threadIdx.x == 0) {
    lt_mbar_init(&B->wfull, 1);
    lt_mbar_init(&B->wfree, 1);
    for (int s = 0; s < LT_STAGES; ++s) { lt_mbar_init(&B->full[s], 1); lt_mbar_init(&B->empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      lt_mbar_init(&B->cur_full[b], 1);
      lt_mbar_init(&B->cur_free[b], 9);  // MMA commit + 8 epilogue warps
      lt_mbar_init(&B->d1_full[b], 1);
      lt_mbar_init(&B->d1_empty[b], 8);
      lt_mbar_init(&B->g_full[b], 8);
      lt_mbar_init(&B->d2_full[b], 1);
      lt_mbar_init(&B->d2_empty[b], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&B->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = B->tmem_base;
  const long long tk0 = clock64();
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_layers = p.l1 - p.l0;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int itg = 0;  // running tile counter across layers (buffer parity / barrier phases)
      for (int li = 0; li < n_layers; ++li) {
        const int layer = p.l0 + li;
        const int dil = 1 << (layer % p.num_stages);
        const int rb = (p.buf0 + li) & 1;  // buffer this layer reads
        const CUtensorMap* mh = rb ? &map_h1 : &map_h0;
        const CUtensorMap* ml = rb ? &map_l1 : &map_l0;
        // weights of this layer (after the MMA warp retired the previous layer's)
        if (li > 0) lt_wait(&B->wfree, (uint32_t)((li - 1) & 1));
        lt_expect_tx(&B->wfull, W_BYTES);
        for (int tap = 0; tap < 3; ++tap) {
          lt_tma_2d(sbase + OFF_WDH + tap * WD_TILE_BYTES, &map_wdh, &B->wfull, 0, (layer * 3 + tap) * 64);
          lt_tma_2d(sbase + OFF_WDL + tap * WD_TILE_BYTES, &map_wdl, &B->wfull, 0, (layer * 3 + tap) * 64);
        }
        lt_tma_2d(sbase + OFF_WRH, &map_wrh, &B->wfull, 0, layer * 64);
        lt_tma_2d(sbase + OFF_WRL, &map_wrl, &B->wfull, 0, layer * 64);
        const float* cond = p.cond + (size_t)li * p.cond_plane;
        for (int it = 0; it < n_my; ++it) {  // conditioning rows stream from HBM: warm L2 early
          const float* ct = cond + (size_t)(blockIdx.x + it * gridDim.x) * LT_BM * C;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ct), "r"(LT_BM * C * 4) : "memory");
        }
        if (li > 0) {
          // grid barrier: every CTA has stored (and fenced) its rows of the previous layer
          const unsigned int target = p.grid_base + (unsigned int)li * gridDim.x;
          long long w0 = 0;
          int spins = 0;
          for (;;) {
            unsigned int seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.grid_counter) : "memory");
            if ((int)(seen - target) >= 0) break;
            if (++spins == 1024) {
              spins = 0;
              if (w0 == 0) w0 = clock64();
              else if (clock64() - w0 > LT_WATCHDOG) { printf("nsw iaf_layer_tc: grid barrier watchdog\n"); __trap(); }
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> TMA reads
        }
        for (int it = 0; it < n_my; ++it, ++itg) {
          const int tile = blockIdx.x + it * gridDim.x;
          const int clip = tile / p.tiles_per_clip;
          const int t0 = (tile - clip * p.tiles_per_clip) * LT_BM;
          for (int tap = 0; tap < 2; ++tap) {
            const int trow = t0 - (2 - tap) * dil;  // negative rows: TMA zero fill = causal padding
            for (int pl = 0; pl < 2; ++pl) {        // lo plane first, then hi
              lt_wait(&B->empty[stage], phase ^ 1);
              lt_expect_tx(&B->full[stage], A_TILE_BYTES);
              lt_tma_3d(sbase + OFF_A + stage * A_TILE_BYTES, pl == 0 ? ml : mh, &B->full[stage], 0, trow, clip);
              if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          {  // current tap: both planes into the double buffer shared with the epilogue
            const int b = itg & 1;
            lt_wait(&B->cur_free[b], (uint32_t)(((itg >> 1) & 1) ^ 1));
            lt_expect_tx(&B->cur_full[b], 2 * A_TILE_BYTES);
            lt_tma_3d(sbase + OFF_CUR + (b * 2 + 0) * A_TILE_BYTES, ml, &B->cur_full[b], 0, t0, clip);
            lt_tma_3d(sbase + OFF_CUR + (b * 2 + 1) * A_TILE_BYTES, mh, &B->cur_full[b], 0, t0, clip);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = lt_idesc();
    int stage = 0;
    uint32_t phase = 0;
    auto mma2 = [&](int j) {  // j: running tile index
      const int b = j & 1;
      const uint32_t u = (uint32_t)(j >> 1);
      lt_wait(&B->g_full[b], u & 1);
      lt_wait(&B->d2_empty[b], (u & 1) ^ 1);
      fence_after();
      if (lane == 0) {
        const uint32_t d2 = tmem + 128 + b * 64;
        const uint32_t g_hi = tmem + 256 + b * 32, g_lo = g_hi + 16;
        const uint64_t wrh = desc_sw64(sbase + OFF_WRH), wrl = desc_sw64(sbase + OFF_WRL);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          mma_ts(d2, g_lo + 8 * k, wrh + 2 * k, idesc, k != 0);
          mma_ts(d2, g_hi + 8 * k, wrl + 2 * k, idesc, 1);
          mma_ts(d2, g_hi + 8 * k, wrh + 2 * k, idesc, 1);
        }
        lt_commit(&B->d2_full[b]);
        if (dbg && j < 6) p.dbg[8 + j] = clock64() - tk0;  // MMA2(j) issued
      }
      __syncwarp();
    };
    int itg = 0;
    for (int li = 0; li < n_layers; ++li) {
      lt_wait(&B->wfull, (uint32_t)(li & 1));
      fence_after();
      if (dbg && lane == 0 && li == 0) p.dbg[0] = clock64() - tk0;  // weights landed
      for (int it = 0; it < n_my; ++it, ++itg) {
        const int b = itg & 1;
        const uint32_t u = (uint32_t)(itg >> 1);
        lt_wait(&B->d1_empty[b], (u & 1) ^ 1);
        fence_after();
        const uint32_t d1 = tmem + b * 64;
        for (int tap = 0; tap < 2; ++tap) {
          const uint64_t wh = desc_sw128(sbase + OFF_WDH + tap * WD_TILE_BYTES);
          const uint64_t wl = desc_sw128(sbase + OFF_WDL + tap * WD_TILE_BYTES);
          // lo plane of the activations: A_lo . W_hi
          lt_wait(&B->full[stage], phase);
          fence_after();
          if (lane == 0) {
            const uint64_t a = desc_sw128(sbase + OFF_A + stage * A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(d1, a + 2 * k, wh + 2 * k, idesc, (tap | k) != 0);
            lt_commit(&B->empty[stage]);
          }
          __syncwarp();
          if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
          // hi plane: A_hi . W_lo + A_hi . W_hi
          lt_wait(&B->full[stage], phase);
          fence_after();
          if (lane == 0) {
            const uint64_t a = desc_sw128(sbase + OFF_A + stage * A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_ss(d1, a + 2 * k, wl + 2 * k, idesc, 1);
              mma_ss(d1, a + 2 * k, wh + 2 * k, idesc, 1);
            }
            lt_commit(&B->empty[stage]);
          }
          __syncwarp();
          if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
        }
        {  // current tap from the shared double buffer
          const uint64_t wh = desc_sw128(sbase + OFF_WDH + 2 * WD_TILE_BYTES);
          const uint64_t wl = desc_sw128(sbase + OFF_WDL + 2 * WD_TILE_BYTES);
          lt_wait(&B->cur_full[b], u & 1);
          fence_after();
          if (lane == 0) {
            const uint64_t alo = desc_sw128(sbase + OFF_CUR + (b * 2 + 0) * A_TILE_BYTES);
            const uint64_t ahi = desc_sw128(sbase + OFF_CUR + (b * 2 + 1) * A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(d1, alo + 2 * k, wh + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_ss(d1, ahi + 2 * k, wl + 2 * k, idesc, 1);
              mma_ss(d1, ahi + 2 * k, wh + 2 * k, idesc, 1);
            }
            lt_commit(&B->cur_free[b]);
            lt_commit(&B->d1_full[b]);
            if (dbg && itg < 6) p.dbg[1 + itg] = clock64() - tk0;  // MMA1 issued
          }
          __syncwarp();
        }
        if (it >= 1) mma2(itg - 1);
      }
      if (n_my >= 1) mma2(itg - 1);
      // every MMA of this layer is issued: when they retire, the weight tiles may be replaced
      if (lane == 0) lt_commit(&B->wfree);
      __syncwarp();
    }
  } else {
    // ================================ epilogue (8 warps) ===========================
    // every tile is handled by all 8 warps: warp w and w+4 share a TMEM lane quarter and
    // split the 64 columns, so each thread owns half a row (32 conv outputs = 16 gates,
    // then 32 residual channels)
    const int half = (warp - 2) >> 2;
    const int qd = warp & 3;  // TMEM lane quarter accessible to this warp
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    const uint32_t scr = sbase + OFF_SCR + (uint32_t)(warp - 2) * 4096u;
    int itg = 0;
    float rmx = 0.f;  // fp16-range guard
    for (int li = 0; li < n_layers; ++li) {
      const int layer = p.l0 + li;
      const int wb = ((p.buf0 + li) & 1) ^ 1;  // buffer this layer writes
      __half* out_hi = p.hi[wb];
      __half* out_lo = p.lo[wb];
      const float* cond = p.cond + (size_t)li * p.cond_plane;
      const float4* bptr = reinterpret_cast<const float4*>(p.br + (size_t)layer * C) + half * 8;
      // conditioning half-rows are requested one tile ahead and sit in registers (tc) while the
      // previous tile's epilogue runs, so their HBM latency is never exposed
      float4 tc[8];
      if (n_my > 0) {
        const size_t w0 = (size_t)blockIdx.x * LT_BM + qd * 32;
        rows_issue128(reinterpret_cast<const float4*>(cond + w0 * C) + half * 8, lane, tc);
      }
      for (int it = 0; it < n_my; ++it, ++itg) {
        const int b = itg & 1;
        const uint32_t u = (uint32_t)(itg >> 1);
        const int tile = blockIdx.x + it * gridDim.x;
        const size_t wrow = (size_t)tile * LT_BM + qd * 32;  // first of this warp's 32 rows
        float4 cq[8];
        rows_transpose128(scr, lane, tc, cq);
        if (it + 1 < n_my) {
          const size_t wn = (size_t)(tile + gridDim.x) * LT_BM + qd * 32;
          rows_issue128(reinterpret_cast<const float4*>(cond + wn * C) + half * 8, lane, tc);
        }
        const bool ed = dbg && warp == 2 && lane == 0 && itg < 6;
        if (ed) p.dbg[32 + itg * 8 + 0] = clock64() - tk0;  // operands loaded
        // ---- E1: gate ----
        lt_wait(&B->d1_full[b], u & 1);
        fence_after();
        if (ed) p.dbg[32 + itg * 8 + 1] = clock64() - tk0;  // D1 ready
        uint32_t d[32];
        tmem_ld32(tmem + lane_sel + b * 64 + half * 32, d);
        tmem_ld_wait();
        fence_before();
        __syncwarp();
        if (lane == 0) lt_arrive(&B->d1_empty[b]);
        uint32_t ghi[8], glo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // columns 4i..4i+3 of this half = (sig j, tanh j, sig j+1, tanh j+1)
          const float s0 = __uint_as_float(d[4 * i]) + cq[i].x;
          const float t0 = __uint_as_float(d[4 * i + 1]) + cq[i].y;
          const float s1 = __uint_as_float(d[4 * i + 2]) + cq[i].z;
          const float t1 = __uint_as_float(d[4 * i + 3]) + cq[i].w;
          const float g0 = gate_fast(s0, t0);
          const float g1 = gate_fast(s1, t1);
          const float h0 = __half2float(__float2half_rn(g0));
          const float h1 = __half2float(__float2half_rn(g1));
          ghi[i] = pack_f16(h0, h1);
          glo[i] = pack_f16(g0 - h0, g1 - h1);
        }
        tmem_st8(tmem + lane_sel + 256 + b * 32 + half * 8, ghi);
        tmem_st8(tmem + lane_sel + 256 + b * 32 + 16 + half * 8, glo);
        tmem_st_wait();
        fence_before();
        __syncwarp();
        if (lane == 0) lt_arrive(&B->g_full[b]);
        if (ed) p.dbg[32 + itg * 8 + 2] = clock64() - tk0;  // g stored
        // ---- E2: residual ----
        lt_wait(&B->d2_full[b], u & 1);
        fence_after();
        if (ed) p.dbg[32 + itg * 8 + 3] = clock64() - tk0;  // D2 ready
        tmem_ld32(tmem + lane_sel + 128 + b * 64 + half * 32, d);
        tmem_ld_wait();
        fence_before();
        __syncwarp();
        if (lane == 0) lt_arrive(&B->d2_empty[b]);
        // l[t] = hi + lo straight from the current-tap operand tiles (SWIZZLE_128B: 16-byte chunk c of
        // row r sits at chunk c ^ (r & 7)); they were loaded for the MMA anyway
        lt_wait(&B->cur_full[b], u & 1);
        const int row = qd * 32 + lane;
        const uint32_t cur_lo = sbase + OFF_CUR + (b * 2 + 0) * A_TILE_BYTES + (uint32_t)row * 128u;
        const uint32_t cur_hi = cur_lo + A_TILE_BYTES;
        uint4 hh[4], ll2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t coff = (uint32_t)(((4 * half + j) ^ (row & 7)) * 16);
          const float4 hraw = lds128(cur_hi + coff), lraw = lds128(cur_lo + coff);
          const __half2* hp = reinterpret_cast<const __half2*>(&hraw);
          const __half2* lp = reinterpret_cast<const __half2*>(&lraw);
          const float4 b0 = __ldg(bptr + 2 * j), b1 = __ldg(bptr + 2 * j + 1);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hv = __half22float2(hp[e]), lv = __half22float2(lp[e]);
            o[2 * e] = (hv.x + lv.x) + bb[2 * e] + __uint_as_float(d[8 * j + 2 * e]);
            o[2 * e + 1] = (hv.y + lv.y) + bb[2 * e + 1] + __uint_as_float(d[8 * j + 2 * e + 1]);
          }
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            range_track_fast(rmx, o[2 * e]);
            range_track_fast(rmx, o[2 * e + 1]);
            const float a0 = __half2float(__float2half_rn(o[2 * e]));
            const float a1 = __half2float(__float2half_rn(o[2 * e + 1]));
            hw[e] = pack_f16(a0, a1);
            lw[e] = pack_f16(o[2 * e] - a0, o[2 * e + 1] - a1);
          }
          hh[j] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          ll2[j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        __syncwarp();
        if (lane == 0) lt_arrive(&B->cur_free[b]);
        rows_store64(reinterpret_cast<uint4*>(out_hi + wrow * C) + half * 4, scr, lane, hh);
        rows_store64(reinterpret_cast<uint4*>(out_lo + wrow * C) + half * 4, scr, lane, ll2);
        if (ed) p.dbg[32 + itg * 8 + 4] = clock64() - tk0;  // tile stored
      }
      if (li + 1 < n_layers) {
        // grid barrier arrival: this CTA's rows of the layer are stored and visible device-wide
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.grid_counter) : "memory");
      }
    }
    range_commit_fast(rmx);
  }

  fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) p.dbg[30] = clock64() - tk0;
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int lt_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NSW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    NSW_CHECK(ptr != nullptr && qres == cudaDriverEntryPointSuccess, NSW_ECUDA,
              "cuTensorMapEncodeTiled is not available from this driver");
    cached = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  *fn = cached;
  return NSW_OK;
}

}  // namespace
}  // namespace nsw
NSW_RANGE_GUARD_TU(layer_tc)
namespace nsw {

// ---- host API used by nsw_iaf.cu (opaque 128-byte tensor maps) ----
int layer_tc_make_act_map(void* map_out, const __half* base, int B, int T) {
  EncodeTiledFn enc;
  NSW_TRY(lt_encode_fn(&enc));
  const cuuint64_t dims[3] = {64, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[2] = {128, (cuuint64_t)T * 128};
  const cuuint32_t box[3] = {64, LT_BM, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<__half*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSW_CHECK(r == CUDA_SUCCESS, NSW_ECUDA, "cuTensorMapEncodeTiled(act) failed: %d", (int)r);
  return NSW_OK;
}

int layer_tc_make_weight_map(void* map_out, const __half* base, int rows, int k, int box_rows) {
  EncodeTiledFn enc;
  NSW_TRY(lt_encode_fn(&enc));
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {(cuuint32_t)k, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<__half*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSW_CHECK(r == CUDA_SUCCESS, NSW_ECUDA, "cuTensorMapEncodeTiled(weight) failed: %d", (int)r);
  return NSW_OK;
}

int layer_tc_launch(const void* const map_act[2][2], const void* map_wdh, const void* map_wdl,
                    const void* map_wrh, const void* map_wrl, const float* cond, size_t cond_plane,
                    __half* const hi[2], __half* const lo[2], const float* br, int T, int rows, int buf0,
                    int l0, int l1, int num_stages, unsigned int* grid_counter, int num_sms,
                    cudaStream_t stream) {
  static std::atomic<uint64_t> attr_done{0};  // per device (the attribute is per device)
  NSW_TRY(ensure_dynamic_smem((const void*)iaf_layer_tc_kernel, (int)LT_SMEM_BYTES, attr_done));
  NSW_CHECK(T % LT_BM == 0, NSW_EINVAL, "layer_tc: T=%d must be a multiple of %d", T, LT_BM);
  NSW_CHECK(l1 > l0, NSW_EINVAL, "layer_tc: empty layer range");
  LayerTcParams p;
  p.cond = cond;
  p.cond_plane = cond_plane;
  p.hi[0] = hi[0]; p.hi[1] = hi[1];
  p.lo[0] = lo[0]; p.lo[1] = lo[1];
  p.br = br;
  p.buf0 = buf0;
  p.l0 = l0;
  p.l1 = l1;
  p.num_stages = num_stages;
  p.tiles_per_clip = T / LT_BM;
  p.n_tiles = rows / LT_BM;
  p.grid_counter = grid_counter;
  p.grid_base = 0;
  NSW_CUDA(cudaMemsetAsync(grid_counter, 0, sizeof(unsigned int), stream));
  p.dbg = nullptr;
  static long long* dbg_buf = nullptr;
  const bool want_dbg = getenv("NSW_LAYER_DEBUG") != nullptr;
  if (want_dbg) {
    if (!dbg_buf) NSW_CUDA(cudaMalloc(&dbg_buf, 128 * sizeof(long long)));
    NSW_CUDA(cudaMemsetAsync(dbg_buf, 0, 128 * sizeof(long long), stream));
    p.dbg = dbg_buf;
  }
  const int grid = std::min(p.n_tiles, num_sms);
  void* args[] = {const_cast<void*>(map_act[0][0]), const_cast<void*>(map_act[0][1]),
                  const_cast<void*>(map_act[1][0]), const_cast<void*>(map_act[1][1]),
                  const_cast<void*>(map_wdh), const_cast<void*>(map_wdl), const_cast<void*>(map_wrh),
                  const_cast<void*>(map_wrl), &p};
  // cooperative launch: every CTA must be co-resident for the inter-layer grid barrier
  NSW_CUDA(cudaLaunchCooperativeKernel((void*)iaf_layer_tc_kernel, dim3(grid), dim3(LT_THREADS), args,
                                       LT_SMEM_BYTES, stream));
  count_launch();
  NSW_CUDA(cudaGetLastError());
  if (want_dbg) {
    long long hbuf[128];
    NSW_CUDA(cudaStreamSynchronize(stream));
    NSW_CUDA(cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[layer_tc dbg layers %d..%d] weights %lld | MMA1 issued:", l0, l1, hbuf[0]);
    for (int i = 0; i < 6; ++i) fprintf(stderr, " %lld", hbuf[1 + i]);
    fprintf(stderr, " | MMA2 issued:");
    for (int i = 0; i < 6; ++i) fprintf(stderr, " %lld", hbuf[8 + i]);
    fprintf(stderr, " | end %lld\n", hbuf[30]);
    for (int i = 0; i < 6; ++i)
      fprintf(stderr, "   tile %d: loaded %lld  D1 %lld  g %lld  D2 %lld  stored %lld\n", i, hbuf[32 + 8 * i],
              hbuf[33 + 8 * i], hbuf[34 + 8 * i], hbuf[35 + 8 * i], hbuf[36 + 8 * i]);
  }
  return NSW_OK;
}

}  // namespace nsw
