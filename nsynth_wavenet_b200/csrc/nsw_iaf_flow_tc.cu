// Engine "tc3": all residual layers of one IAF flow (parallel_wavenet.py:227-254) in ONE persistent
// kernel whose residual stream never leaves the SM.
//
// Work split.  Every clip's time axis is cut into 128-row tiles; consecutive tiles of ONE clip are
// owned by one CTA (<= 4 tiles, one CTA per SM, ranges never straddle clips).  The CTA keeps its
// rows resident in shared memory as the split pair (fp16 hi plane, fp16 lo plane, 22 mantissa bits)
// in the K-major SWIZZLE_128B layout tcgen05.mma reads, for all layers of the flow:
//
//   plane = [halo tile 16 KB][own tile 0][own tile 1][own tile 2][own tile 3]      (x2: hi, lo)
//
// A causal tap t-o of tile k is just a matrix descriptor whose start address is o rows earlier
// (the swizzle is a function of absolute smem address bits, so any row offset is legal - probed in
// scripts/probes/umma_rowshift_probe.cu).  For dilations d <= 64 every tap of every own tile lies in
// [halo | own]; the halo (last tile of the previous CTA, or zeros at the start of a clip) is the
// only thing fetched per layer.  For d >= 128 taps are whole tiles: own ones are read in place,
// foreign ones stream through the halo slot.
//
// A layer updates the rows IN PLACE.  Tiles are processed latest-first, so a tile's old values are
// still there for the later tiles' past taps (their MMAs were issued earlier and tcgen05.mma
// retires in order), and layer i+1 can start on the latest tile while layer i is still working on
// earlier ones: there is no grid-wide barrier.  Cross-CTA dependencies are per-tile "published
// layer" counters in global memory (written after the tile's TMA store has completed) plus one
// "consumed layer" counter per CTA that protects the ping-pong global buffers from being
// overwritten before the neighbours have read them.
//
// Per tile and layer (as in engine tc2):
//   D1[128x64] = sum_tap A_tap . Wd_tap (split fp16: lo.hi + hi.lo + hi.hi, fp32 accumulate in TMEM)
//   g = sigmoid(D1[:, even] + cond[:, even]) * tanh(D1[:, odd] + cond[:, odd])   -> TMEM (fp16 hi|lo)
//   L[128x64] += g . Wr (A operand from TMEM; L = the tile's fp32 residual rows, resident in TMEM)
//   l = L + br  ->  written back to TMEM and re-split into the fp16 hi / lo planes in shared memory
//
// Warps: 0 loader (TMA: weights, own tiles, halo / foreign tiles; polls the flags), 1 MMA1 issuer,
// 2-9 gate epilogue (E1), 10-17 residual epilogue (E2), 18 publisher (TMA stores + flags), 19 MMA2
// issuer.  Every role walks the same task sequence (layer-major, latest tile first) in order.  The conditioning planes come from the conv-GEMM in a row-interleaved layout
// ([tile][quarter][half][j][row][4 floats]) so each warp-level LDG.128 reads 512 contiguous bytes
// and needs no shared memory.
#include "nsw_gemm.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>

namespace nsw {

namespace {

constexpr int C = 64;
constexpr int BM = 128;
constexpr int KMAX = 4;
constexpr int FT_THREADS = 640;  // loader, MMA1, 8 gate warps, 8 residual warps, publisher, MMA2
constexpr uint32_t TILE_B = BM * 128;                  // one plane of one tile: 16 KB
constexpr uint32_t PLANE_B = (1 + KMAX) * TILE_B;      // halo + own tiles: 80 KB
constexpr uint32_t WD_TILE = 64 * 64 * 2;              // 8 KB per tap per plane
constexpr uint32_t WR_TILE = 64 * 32 * 2;              // 4 KB per plane
constexpr uint32_t OFF_HI = 0, OFF_LO = PLANE_B;
constexpr uint32_t OFF_WDH = 2 * PLANE_B, OFF_WDL = OFF_WDH + 3 * WD_TILE;
constexpr uint32_t OFF_WR = OFF_WDL + 3 * WD_TILE;     // 2 buffers x [hi 4 KB][lo 4 KB]
constexpr uint32_t OFF_BARS = OFF_WR + 4 * WR_TILE;
constexpr size_t FT_SMEM_BYTES = OFF_BARS + 1024 + 1024;
constexpr long long FT_WATCHDOG = 4000000000ll;
// TMEM columns: D1[2] (conv accumulators, preloaded with cond) 0..127, G[2] (gate, fp16 hi 16 | lo 16) 128..191,
// L[4] (the fp32 residual rows of the own tiles; MMA2 accumulates g.Wr straight into them) 192..447
constexpr uint32_t TM_G = 128, TM_L = 192;

struct FtBars {
  uint64_t own_loaded[KMAX];
  uint64_t tile_ready[KMAX];
  uint64_t pub_done[KMAX];
  uint64_t ring_full, ring_free;
  uint64_t wd_full, wd_free;
  uint64_t wr_full[2], wr_free[2];
  uint64_t d1_full[2], d1_empty[2], g_full[2], g_free[2];
  uint64_t d2_full[KMAX];  // per own tile: MMA2 of the current layer has retired
  uint64_t l_init;
  uint64_t halo0;  // fused start conv: the halo rows of layer 0 are in shared memory
  uint32_t tmem_base;
};
static_assert(sizeof(FtBars) <= 1024, "barrier block grew");

// ---------------------------------- PTX wrappers ----------------------------------
__device__ __forceinline__ void ft_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ft_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void ft_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool ft_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void ft_die(const char* what) {
  printf("nsw iaf_flow_tc: watchdog in %s (block %d thread %d)\n", what, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void ft_wait(uint64_t* bar, uint32_t parity, const char* what) {
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
    if (++spins == 1024) {
      spins = 0;
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > FT_WATCHDOG) ft_die(what);
    }
  }
}
__device__ __forceinline__ void ft_poll_ge(const unsigned int* p, unsigned int target, const char* what) {
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    unsigned int seen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p) : "memory");
    if ((int)(seen - target) >= 0) return;
    if (++spins == 256) {
      spins = 0;
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > FT_WATCHDOG) ft_die(what);
    }
  }
}
__device__ __forceinline__ void ft_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                               int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ft_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                               int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void ft_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ uint64_t ft_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t ft_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// D = f32, A = B = f16, K-major, M = 128, N = 64
__device__ __forceinline__ uint32_t ft_idesc() {
  return (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void ft_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ft_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ft_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool ft_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ft_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void ft_tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ft_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ft_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void ft_tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void ft_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// streaming 16-byte load that does not allocate in L1: the L1 / shared-memory data array is the scarce resource
// of this kernel (the SS MMAs alone fetch their operands at ~125 of its 128 B/clk)
__device__ __forceinline__ float4 ft_ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void ft_sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ft_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(saddr)
               : "memory");
  return v;
}

// sigmoid(s) * tanh(t) = (1 - b) / ((1 + a)(1 + b)), a = e^-s, b = e^-2t : 3 MUFU ops
__device__ __forceinline__ float ft_gate(float s, float t) {
  const float a = __expf(fminf(-s, 40.0f));
  const float b = __expf(fminf(-2.0f * t, 40.0f));
  return __fdividef(1.0f - b, (1.0f + a) * (1.0f + b));
}
__device__ __forceinline__ uint32_t ft_pack_f16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct FlowTcParams {
  const float* cond;      // tiled conditioning plane of layer l0 (see file header)
  size_t cond_plane;      // floats between consecutive layers' planes
  const float* br;        // [L][64] RUNNING SUM over layers of the residual biases, natural channel order
  unsigned int* flags;    // [clips * tiles_per_clip] layers published per tile   (zeroed per launch)
  unsigned int* cons;     // [grid] layers whose foreign loads a CTA has finished (zeroed per launch)
  int buf0, l0, l1, num_stages;
  int tiles_per_clip, clip0, nclips;
  int cta_base, cta_rem;  // clip c (0-based within the launch) is split over cta_base + (c < cta_rem) CTAs
  int reach_tiles;        // 2 * max dilation / 128 (how far ahead a published tile is read)
  long long* dbg;
  int dbg_cta;            // block whose timeline NSW_LAYER_DEBUG records
  // fused head (parallel_wavenet.py:256-277, 316-330, 348-359): a pseudo-layer after layer l1-1 whose
  // "conv" is the 1x1 out1 on relu(l) (weights = tile head_w_tile of the Wd maps) accumulated on the
  // out1 conditioning plane, and whose epilogue is out2_mean / out2_scale + the affine composition
  int fuse_head;
  int head_w_tile;        // 64-row tile index of W1^T in the Wd tensor maps
  const float* wm;        // out2_mean W [64]
  const float* ws;        // out2_scale W [64]
  float bm, bs;
  const float* x_in;      // input of this flow [B*T]
  const float* z;         // noise [B*T]
  float* x_out;
  float* mean_tot;
  float* scale_tot;
  float* log_scale_tot;
  int T;                  // samples per clip
  int first, last, quantize, use_mu_law;
  float quant_chann;
  // fused start conv (parallel_wavenet.py:222-225): l0[t] = b + W0 x[t-3] + W1 x[t-2] + W2 x[t-1] computed by
  // the residual warps from start_x instead of loading the rows a separate kernel wrote (needs l0 == 0)
  int fuse_start;
  const float* start_x;   // input of this flow [B*T]
  const float* start_w;   // [3][64]
  const float* start_b;   // [64]
};

// ---- work split and publish rules: plain integer functions shared by the kernel and by the host test hook
//      nsw_flow_plan_host (tests/test_flow_plan.py checks on the CPU that every cross-CTA read has a publish) ----
struct Range {  // the tiles a CTA owns
  int clip, tk0, K, idx, n_c;
};
__host__ __device__ __forceinline__ Range ft_range_of(int tiles_per_clip, int cta_base, int cta_rem, int b) {
  Range r;
  const int big = cta_rem * (cta_base + 1);
  int c;
  if (b < big) {
    c = b / (cta_base + 1);
    r.idx = b - c * (cta_base + 1);
    r.n_c = cta_base + 1;
  } else {
    const int bb = b - big;
    c = cta_rem + bb / cta_base;
    r.idx = bb - (c - cta_rem) * cta_base;
    r.n_c = cta_base;
  }
  r.clip = c;
  const int bt = tiles_per_clip / r.n_c, rt = tiles_per_clip - bt * r.n_c;
  r.K = bt + (r.idx < rt ? 1 : 0);
  r.tk0 = r.idx * bt + (r.idx < rt ? r.idx : rt);
  return r;
}
__device__ __forceinline__ Range ft_range(const FlowTcParams& p, int b) {
  return ft_range_of(p.tiles_per_clip, p.cta_base, p.cta_rem, b);
}
// index (within its clip's CTA group) of the CTA that owns tile tk of a clip split over n_c CTAs
__host__ __device__ __forceinline__ int ft_owner(int tiles_per_clip, int n_c, int tk) {
  const int bt = tiles_per_clip / n_c, rt = tiles_per_clip - bt * n_c;
  const int big = rt * (bt + 1);
  return tk < big ? tk / (bt + 1) : rt + (tk - big) / bt;
}
// does the output of (own tile k of K, launch-relative layer li of nl) have a reader outside its CTA?
__host__ __device__ __forceinline__ bool ft_published(int k, int K, int li, int nl, int l0, int num_stages,
                                                      int fuse_head) {
  if (li == nl - 1) return fuse_head == 0;                      // the head kernel reads every row
  if ((1 << ((l0 + li + 1) % num_stages)) >= BM) return true;   // next layer reads whole foreign tiles
  return k == K - 1;                                            // next CTA's halo
}
// where tap `tap` (0: t-2d, 1: t-d) of own tile k comes from in a layer with dilation d:
//   >= 0 : foreign tile (index inside the clip) fetched from the published global copy
//   -1   : own shared memory (including the current tile itself)      -2 : causal zeros, nothing is read
// (for d <= 64 the only foreign tile is the halo tk0-1 of own tile 0, shared by both taps)
__host__ __device__ __forceinline__ int ft_tap_source(int tk0, int k, int tap, int d) {
  if (2 * d <= BM) return (k == 0) ? (tk0 >= 1 ? tk0 - 1 : -2) : -1;
  const int src = tk0 + k - (2 - tap) * (d / BM);
  return src < 0 ? -2 : (src >= tk0 ? -1 : src);
}

__global__ void __launch_bounds__(FT_THREADS, 1)
iaf_flow_tc_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_l0,
                   const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_l1,
                   const __grid_constant__ CUtensorMap map_wdh, const __grid_constant__ CUtensorMap map_wdl,
                   const __grid_constant__ CUtensorMap map_wrh, const __grid_constant__ CUtensorMap map_wrl,
                   FlowTcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FtBars* B = reinterpret_cast<FtBars*>(smem + OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int k = 0; k < KMAX; ++k) {
      ft_mbar_init(&B->own_loaded[k], p.fuse_start ? 8 : 1);
      ft_mbar_init(&B->tile_ready[k], 8);
      ft_mbar_init(&B->pub_done[k], 1);
      ft_mbar_init(&B->d2_full[k], 1);
    }
    ft_mbar_init(&B->ring_full, 1);
    ft_mbar_init(&B->ring_free, 1);
    ft_mbar_init(&B->wd_full, 1);
    ft_mbar_init(&B->wd_free, 1);
    ft_mbar_init(&B->l_init, 8);
    ft_mbar_init(&B->halo0, 8);
    for (int b = 0; b < 2; ++b) {
      ft_mbar_init(&B->wr_full[b], 1);
      ft_mbar_init(&B->wr_free[b], 1);
      ft_mbar_init(&B->d1_full[b], 1);
      ft_mbar_init(&B->d1_empty[b], 8);
      ft_mbar_init(&B->g_full[b], 8);
      ft_mbar_init(&B->g_free[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ft_fence_before();
  __syncthreads();
  ft_fence_after();
  const uint32_t tmem = B->tmem_base;

  const Range R = ft_range(p, (int)blockIdx.x);
  const int K = R.K;
  const int nl = p.l1 - p.l0;
  const int total = nl * K;                           // gate / residual tasks
  const int total1 = (nl + (p.fuse_head ? 1 : 0)) * K;  // MMA1 / E1 tasks (+ the head pseudo-layer)
  const int gclip = p.clip0 + R.clip;                 // clip coordinate in the tensor maps
  const int gt0 = gclip * p.tiles_per_clip + R.tk0;   // global tile id of own tile 0
  const long long tk_start = clock64();
  unsigned long long gt_start = 0;
  if (p.dbg != nullptr && (int)blockIdx.x == p.dbg_cta && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_start));
  const bool dbg = p.dbg != nullptr && (int)blockIdx.x == p.dbg_cta;

  // does the output of (own tile k, layer li) have a reader outside this CTA?
  auto published = [&](int k, int li) -> bool {
    return ft_published(k, K, li, nl, p.l0, p.num_stages, p.fuse_head);
  };

  if (warp == 0) {
    // =================================== loader ===================================
    if (lane == 0) {
      auto load_wd = [&](int layer) {
        ft_expect_tx(&B->wd_full, 6 * WD_TILE);
        for (int tap = 0; tap < 3; ++tap) {
          ft_tma_load_2d(sbase + OFF_WDH + tap * WD_TILE, &map_wdh, &B->wd_full, 0, (layer * 3 + tap) * 64);
          ft_tma_load_2d(sbase + OFF_WDL + tap * WD_TILE, &map_wdl, &B->wd_full, 0, (layer * 3 + tap) * 64);
        }
      };
      auto load_wr = [&](int layer, int b) {
        ft_expect_tx(&B->wr_full[b], 2 * WR_TILE);
        ft_tma_load_2d(sbase + OFF_WR + (2 * b) * WR_TILE, &map_wrh, &B->wr_full[b], 0, layer * 64);
        ft_tma_load_2d(sbase + OFF_WR + (2 * b + 1) * WR_TILE, &map_wrl, &B->wr_full[b], 0, layer * 64);
      };
      auto prefetch_cond = [&](int li) {
        if (li >= nl) return;
        const float* c = p.cond + (size_t)li * p.cond_plane + (size_t)gt0 * (BM * C);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c), "r"((uint32_t)(K * BM * C * 4)) : "memory");
      };
      load_wd(p.l0);
      load_wr(p.l0, 0);
      if (nl > 1) load_wr(p.l0 + 1, 1);
      if (!p.fuse_start) {
        const CUtensorMap* mh = (p.buf0 & 1) ? &map_h1 : &map_h0;
        const CUtensorMap* ml = (p.buf0 & 1) ? &map_l1 : &map_l0;
        for (int k = K - 1; k >= 0; --k) {
          ft_expect_tx(&B->own_loaded[k], 2 * TILE_B);
          ft_tma_load_3d(sbase + OFF_LO + (1 + k) * TILE_B, ml, &B->own_loaded[k], 0, (R.tk0 + k) * BM, gclip);
          ft_tma_load_3d(sbase + OFF_HI + (1 + k) * TILE_B, mh, &B->own_loaded[k], 0, (R.tk0 + k) * BM, gclip);
        }
      }
      prefetch_cond(0);
      prefetch_cond(1);
      int n_ring = 0;
      auto foreign = [&](int li, int src_tk, const CUtensorMap* mh, const CUtensorMap* ml) {
        // src_tk may be -1 (halo before the clip start): TMA zero-fills the out-of-range rows
        if (li > 0 && src_tk >= 0)
          ft_poll_ge(p.flags + (size_t)gclip * p.tiles_per_clip + src_tk, (unsigned int)li, "flag poll");
        if (n_ring > 0) ft_wait(&B->ring_free, (uint32_t)((n_ring - 1) & 1), "ring_free");
        ft_fence_async();
        ft_expect_tx(&B->ring_full, 2 * TILE_B);
        ft_tma_load_3d(sbase + OFF_LO, ml, &B->ring_full, 0, src_tk * BM, gclip);
        ft_tma_load_3d(sbase + OFF_HI, mh, &B->ring_full, 0, src_tk * BM, gclip);
        ++n_ring;
      };
      for (int li = 0; li < nl; ++li) {
        const int layer = p.l0 + li;
        const int d = 1 << (layer % p.num_stages);
        if (li > 0) {
          ft_wait(&B->wd_free, (uint32_t)((li - 1) & 1), "wd_free");
          load_wd(layer);
          if (li >= 2) {
            ft_wait(&B->wr_free[li & 1], (uint32_t)(((li >> 1) - 1) & 1), "wr_free");
            load_wr(layer, li & 1);
          }
          prefetch_cond(li + 1);
        }
        const int rb = (p.buf0 + li) & 1;  // global buffer holding the previous layer's output
        const CUtensorMap* mh = rb ? &map_h1 : &map_h0;
        const CUtensorMap* ml = rb ? &map_l1 : &map_l0;
        if (2 * d <= BM) {
          // halo for own tile 0 (the last task of the layer); with the fused start conv the residual
          // warps compute layer 0's halo themselves
          if (!(p.fuse_start && li == 0)) foreign(li, R.tk0 - 1, mh, ml);
        } else {
          const int dt = d / BM;
          for (int k = K - 1; k >= 0; --k)
            for (int tap = 1; tap >= 0; --tap) {  // same order as the MMA issuer: t-d, then t-2d
              const int src = R.tk0 + k - (2 - tap) * dt;
              if (src < 0 || src >= R.tk0) continue;  // causal zeros / own tile
              foreign(li, src, mh, ml);
            }
        }
      }
      if (p.fuse_head) {  // W1^T (hi, lo) into the current-tap weight slot once the last layer's MMA1s retired
        ft_wait(&B->wd_free, (uint32_t)((nl - 1) & 1), "wd_free (head)");
        ft_expect_tx(&B->wd_full, 2 * WD_TILE);
        ft_tma_load_2d(sbase + OFF_WDH + 2 * WD_TILE, &map_wdh, &B->wd_full, 0, p.head_w_tile * 64);
        ft_tma_load_2d(sbase + OFF_WDL + 2 * WD_TILE, &map_wdl, &B->wd_full, 0, p.head_w_tile * 64);
      }
    }
  } else if (warp == 1) {
    // =================================== MMA1 issuer ===================================
    // D1 = cond (preloaded by the gate warps) + sum_tap A_tap . Wd_tap.  All 32 lanes run the loop
    // with warp-uniform control flow so ptxas keeps the descriptors in uniform registers; one elected
    // lane issues.  (A scheduler run by a single divergent lane cost ~16 instructions per
    // tcgen05.mma, and a combined MMA1/MMA2 polling scheduler ~500 instructions per task: the scalar
    // instruction stream of this warp, not the tensor pipe, was the limiter.)
    {
      const uint32_t idesc = ft_idesc();
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
      auto wait_ready = [&](int x, int li) {  // own tile x holds layer li-1's output, visible to the async proxy
        if (li == 0) ft_wait(&B->own_loaded[x], 0, "own_loaded");
        else ft_wait(&B->tile_ready[x], (uint32_t)((li - 1) & 1), "tile_ready (mma)");
      };
      auto issue_tap = [&](uint32_t d1, uint32_t a_row_bytes, int tap) {
        const uint64_t alo = ft_desc_sw128(sbase_u + OFF_LO + a_row_bytes);
        const uint64_t ahi = ft_desc_sw128(sbase_u + OFF_HI + a_row_bytes);
        const uint64_t wh = ft_desc_sw128(sbase_u + OFF_WDH + tap * WD_TILE);
        const uint64_t wl = ft_desc_sw128(sbase_u + OFF_WDL + tap * WD_TILE);
        if (ft_elect()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ft_mma_ss(d1, alo + 2 * k, wh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ft_mma_ss(d1, ahi + 2 * k, wl + 2 * k, idesc, 1);
            ft_mma_ss(d1, ahi + 2 * k, wh + 2 * k, idesc, 1);
          }
        }
        __syncwarp();
      };
      // One burst per task: every wait and every descriptor is done BEFORE the first tcgen05.mma of the
      // task, because the tensor pipe only buffers a few instructions: a pause between two taps
      // (barrier test, fence, descriptor arithmetic) is a pipe bubble.  Only foreign whole-tile taps
      // (d >= 128, one shared landing slot) are issued separately, after the burst.
      auto issue_burst = [&](uint32_t d1, const uint32_t (&a_bytes)[3], const int (&taps)[3], int n) {
        uint64_t alo[3], ahi[3], wh[3], wl[3];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          alo[s2] = ft_desc_sw128(sbase_u + OFF_LO + a_bytes[s2]);
          ahi[s2] = ft_desc_sw128(sbase_u + OFF_HI + a_bytes[s2]);
          wh[s2] = ft_desc_sw128(sbase_u + OFF_WDH + taps[s2] * WD_TILE);
          wl[s2] = ft_desc_sw128(sbase_u + OFF_WDL + taps[s2] * WD_TILE);
        }
        if (ft_elect()) {
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            if (s2 < n) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ft_mma_ss(d1, alo[s2] + 2 * k, wh[s2] + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ft_mma_ss(d1, ahi[s2] + 2 * k, wl[s2] + 2 * k, idesc, 1);
                ft_mma_ss(d1, ahi[s2] + 2 * k, wh[s2] + 2 * k, idesc, 1);
              }
            }
          }
        }
        __syncwarp();
      };
      int stage_mod = p.l0 % p.num_stages;
      int n_ring = 0, j = 0;
      for (int li = 0; li < nl; ++li) {
        const int d = 1 << stage_mod;
        if (++stage_mod == p.num_stages) stage_mod = 0;
        const bool small = 2 * d <= BM;
        const int dt = d / BM;
        ft_wait(&B->wd_full, (uint32_t)(li & 1), "wd_full");
        if (dbg && lane == 0 && li < 32) p.dbg[80 + li] = clock64() - tk_start;
        long long w_d1 = 0, w_rdy = 0, w_ring = 0;  // NSW_LAYER_DEBUG: where this layer's issuer time went
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int k = K - 1 - kk;
          const int b = j & 1;
          const uint32_t d1 = tmem_u + b * 64;
          uint32_t a_bytes[3] = {(uint32_t)(1 + k) * TILE_B, 0u, 0u};
          int taps[3] = {2, 1, 0};
          int n = 1;
          bool foreign[2] = {false, false};  // [tap] of the d >= 128 layers
          bool halo_used = false;
          long long tw0 = dbg ? clock64() : 0;
          // D1[b] holds the conditioning rows of this task (written by the gate warps)
          ft_wait(&B->d1_empty[b], (uint32_t)((j >> 1) & 1), "d1_empty");
          if (dbg) { const long long now = clock64(); w_d1 += now - tw0; tw0 = now; }
          if (dbg && j < 8 && lane == 0) p.dbg[64 + 2 * j] = clock64() - tk_start;
          wait_ready(k, li);
          if (small) {
            // windows [128k - o, 128k - o + 128) of [halo | own], o = d, 2d
            const bool own_halo = p.fuse_start && li == 0;
            if (k >= 1) wait_ready(k - 1, li);
            else if (own_halo) ft_wait(&B->halo0, 0, "halo0");
            else { ft_wait(&B->ring_full, (uint32_t)(n_ring & 1), "ring_full (halo)"); halo_used = true; }
            a_bytes[1] = (uint32_t)((1 + k) * BM - d) * 128u;
            a_bytes[2] = (uint32_t)((1 + k) * BM - 2 * d) * 128u;
            n = 3;
          } else {
            for (int tap = 1; tap >= 0; --tap) {
              const int src = R.tk0 + k - (2 - tap) * dt;
              if (src < 0) continue;  // causal zeros
              if (src >= R.tk0) {
                wait_ready(src - R.tk0, li);
                a_bytes[n] = (uint32_t)(1 + src - R.tk0) * TILE_B;
                taps[n] = tap;
                ++n;
              } else {
                foreign[tap] = true;
              }
            }
          }
          if (dbg) { const long long now = clock64(); w_rdy += now - tw0; tw0 = now; }
          if (dbg && j < 8 && lane == 0) p.dbg[2 * j] = clock64() - tk_start;
          ft_fence_after();
          issue_burst(d1, a_bytes, taps, n);
          if (halo_used) {
            if (ft_elect()) ft_commit(&B->ring_free);
            __syncwarp();
            ++n_ring;
          }
          for (int tap = 1; tap >= 0; --tap) {  // same order as the loader: t-d, then t-2d
            if (!foreign[tap]) continue;
            tw0 = dbg ? clock64() : 0;
            ft_wait(&B->ring_full, (uint32_t)(n_ring & 1), "ring_full");
            if (dbg) w_ring += clock64() - tw0;
            ft_fence_after();
            issue_tap(d1, 0u, tap);
            if (ft_elect()) ft_commit(&B->ring_free);
            __syncwarp();
            ++n_ring;
          }
          if (ft_elect()) {
            ft_commit(&B->d1_full[b]);
            if (kk == K - 1) {
              ft_commit(&B->wd_free);
              // every foreign tile of this layer has landed in shared memory: the global buffer it
              // came from may be overwritten (two layers from now) as far as this CTA is concerned
              // relaxed is enough: the foreign reads completed before the mbarrier flip this thread observed
              asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.cons + blockIdx.x), "r"((unsigned int)(li + 1))
                           : "memory");
            }
            if (dbg && j < 8) p.dbg[2 * j + 1] = clock64() - tk_start;
          }
          __syncwarp();
        }
        if (dbg && lane == 0 && li < 32) {
          p.dbg[128 + li] = w_d1;
          p.dbg[160 + li] = w_rdy;
          p.dbg[192 + li] = w_ring;
        }
      }
      if (p.fuse_head) {
        // head pseudo-layer: D1 = cond_out1 (preloaded) + relu(l) . W1, one "tap" on the tile itself
        ft_wait(&B->wd_full, (uint32_t)(nl & 1), "wd_full (head)");
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int k = K - 1 - kk;
          const int b = j & 1;
          ft_wait(&B->d1_empty[b], (uint32_t)((j >> 1) & 1), "d1_empty (head)");
          wait_ready(k, nl);
          ft_fence_after();
          issue_tap(tmem_u + b * 64, (uint32_t)(1 + k) * TILE_B, 2);
          if (ft_elect()) ft_commit(&B->d1_full[b]);
          __syncwarp();
        }
      }
    }
  } else if (warp == 19) {
    // =================================== MMA2 issuer ===================================
    // D2 = g . Wr with the A operand (g, fp16 hi | lo) read from TMEM
    {
      const uint32_t idesc = ft_idesc();
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
      int j = 0;
      ft_wait(&B->l_init, 0, "l_init");  // the residual warps have written l into TMEM
      for (int li = 0; li < nl; ++li) {
        ft_wait(&B->wr_full[li & 1], (uint32_t)((li >> 1) & 1), "wr_full");
        const uint64_t wrh = ft_desc_sw64(sbase_u + OFF_WR + (2 * (li & 1)) * WR_TILE);
        const uint64_t wrl = ft_desc_sw64(sbase_u + OFF_WR + (2 * (li & 1) + 1) * WR_TILE);
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int b = j & 1;
          const uint32_t u = (uint32_t)(j >> 1);
          ft_wait(&B->g_full[b], u & 1, "g_full");
          if (dbg && j < 8 && lane == 0) p.dbg[16 + 2 * j] = clock64() - tk_start;
          ft_fence_after();
          const uint32_t d2 = tmem_u + TM_L + (K - 1 - kk) * 64;  // the tile's fp32 residual rows
          const uint32_t g_hi = tmem_u + TM_G + b * 32, g_lo = g_hi + 16;
          if (ft_elect()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              ft_mma_ts(d2, g_lo + 8 * k, wrh + 2 * k, idesc, 1);
              ft_mma_ts(d2, g_hi + 8 * k, wrl + 2 * k, idesc, 1);
              ft_mma_ts(d2, g_hi + 8 * k, wrh + 2 * k, idesc, 1);
            }
            ft_commit(&B->d2_full[K - 1 - kk]);
            ft_commit(&B->g_free[b]);
            if (kk == K - 1) ft_commit(&B->wr_free[li & 1]);
            if (dbg && j < 8) p.dbg[16 + 2 * j + 1] = clock64() - tk_start;
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 10) {
    // =================================== E1: gate (8 warps) ===================================
    // warps w and w+4 share a TMEM lane quarter and split the 64 columns: a thread owns half a row
    // (32 conv outputs = 16 gates).  The conditioning half-row of task j+2 is written into the D1
    // accumulator as soon as task j's values have been read out of it, so the MMAs accumulate on top
    // of it and the gate needs no separate add.
    const int half = (warp - 2) >> 2;
    const int qd = warp & 3;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    auto cond_ptr = [&](int j) -> const float4* {
      const int li = j / K, k = K - 1 - (j - li * K);
      return reinterpret_cast<const float4*>(p.cond + (size_t)li * p.cond_plane +
                                             ((size_t)(gt0 + k) * 8 + qd * 2 + half) * 1024) + lane;
    };
    auto preload = [&](int j, const float4 (&cn)[8]) {  // cond(j) -> D1[j & 1], then release the buffer
      const int b = j & 1;
      ft_tmem_st32(tmem + lane_sel + b * 64 + half * 32, reinterpret_cast<const uint32_t*>(cn));
      ft_tmem_st_wait();
      ft_fence_before();
      __syncwarp();
      if (lane == 0) ft_arrive(&B->d1_empty[b]);
    };
    for (int j = 0; j < 2 && j < total1; ++j) {
      float4 cn[8];
      const float4* src = cond_ptr(j);
#pragma unroll
      for (int i = 0; i < 8; ++i) cn[i] = ft_ldg_stream(src + i * 32);
      preload(j, cn);
    }
    for (int j = 0; j < total1; ++j) {
      const int b = j & 1;
      float4 cn[8];
      if (j + 2 < total1) {
        const float4* src = cond_ptr(j + 2);
#pragma unroll
        for (int i = 0; i < 8; ++i) cn[i] = ft_ldg_stream(src + i * 32);
      }
      ft_wait(&B->d1_full[b], (uint32_t)((j >> 1) & 1), "d1_full");
      if (dbg && warp == 2 && lane == 0 && j < 8) p.dbg[32 + 2 * j] = clock64() - tk_start;
      ft_fence_after();
      uint32_t d[32];
      ft_tmem_ld32(tmem + lane_sel + b * 64 + half * 32, d);
      ft_tmem_ld_wait();
      if (j + 2 < total1) preload(j + 2, cn);
      if (j >= total) {
        // ---------------- head epilogue (pseudo-layer task) ----------------
        // d = out1(relu(l)) + cond_out1 for this thread's 32 channels; out2_mean / out2_scale are 64-wide
        // dot products: the two warps sharing the rows exchange their partial sums through the (now idle)
        // tap-0 weight slot
        const int k = K - 1 - (j - total);
        const float4* wm4 = reinterpret_cast<const float4*>(p.wm) + half * 8;
        const float4* ws4 = reinterpret_cast<const float4*>(p.ws) + half * 8;
        float pm = 0.f, ps = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 wmv = __ldg(wm4 + i), wsv = __ldg(ws4 + i);
          const float h0 = fmaxf(__uint_as_float(d[4 * i]), 0.f), h1 = fmaxf(__uint_as_float(d[4 * i + 1]), 0.f);
          const float h2 = fmaxf(__uint_as_float(d[4 * i + 2]), 0.f), h3 = fmaxf(__uint_as_float(d[4 * i + 3]), 0.f);
          pm = fmaf(h0, wmv.x, fmaf(h1, wmv.y, fmaf(h2, wmv.z, fmaf(h3, wmv.w, pm))));
          ps = fmaf(h0, wsv.x, fmaf(h1, wsv.y, fmaf(h2, wsv.z, fmaf(h3, wsv.w, ps))));
        }
        const int row = qd * 32 + lane;
        const uint32_t exch = sbase + OFF_WDH + (uint32_t)(j & 1) * 1024u + (uint32_t)row * 8u;
        if (half == 1) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(exch), "f"(pm), "f"(ps) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        if (half == 0) {
          float qm, qs;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(qm), "=f"(qs) : "r"(exch) : "memory");
          const size_t grow = (size_t)gclip * p.T + (size_t)(R.tk0 + k) * BM + row;
          const float mean = pm + qm + p.bm;
          const float sp = softplusf_acc(ps + qs + p.bs);
          const float scale = fminf(fmaxf(sp, 1.2340980408667956e-4f /*e^-9*/), 1096.6331584284585f /*e^7*/);
          const float log_scale = logf(scale);
          float mt, st, lt;
          if (p.first) {
            mt = mean; st = scale; lt = log_scale;
          } else {
            mt = fmaf(p.mean_tot[grow], scale, mean);
            st = p.scale_tot[grow] * scale;
            lt = p.log_scale_tot[grow] + log_scale;
          }
          float xo = fmaf(p.x_in[grow], scale, mean);
          if (p.last) {
            st = fminf(st, 1096.6331584284585f);
            lt = fminf(lt, 7.0f);
            xo = fmaf(p.z[grow], st, mt);  // new_x = x * scale_tot + mean_tot (:330)
            if (p.quantize) xo = clip_quant_scale_dev(xo, p.quant_chann, p.use_mu_law);
          }
          p.mean_tot[grow] = mt;
          p.scale_tot[grow] = st;
          p.log_scale_tot[grow] = lt;
          p.x_out[grow] = xo;
        }
        continue;
      }
      // G[b] was last read by MMA2 of task j-2 (issued by another warp): wait until it has retired
      if (j >= 2) ft_wait(&B->g_free[b], (uint32_t)((((j - 2) >> 1)) & 1), "g_free");
      uint32_t ghi[8], glo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // columns 4i..4i+3 of this half = (sig jj, tanh jj, sig jj+1, tanh jj+1), cond already added
        const float g0 = ft_gate(__uint_as_float(d[4 * i]), __uint_as_float(d[4 * i + 1]));
        const float g1 = ft_gate(__uint_as_float(d[4 * i + 2]), __uint_as_float(d[4 * i + 3]));
        const float h0 = __half2float(__float2half_rn(g0));
        const float h1 = __half2float(__float2half_rn(g1));
        ghi[i] = ft_pack_f16(h0, h1);
        glo[i] = ft_pack_f16(g0 - h0, g1 - h1);
      }
      ft_tmem_st8(tmem + lane_sel + TM_G + b * 32 + half * 8, ghi);
      ft_tmem_st8(tmem + lane_sel + TM_G + b * 32 + 16 + half * 8, glo);
      ft_tmem_st_wait();
      ft_fence_before();
      __syncwarp();
      if (lane == 0) ft_arrive(&B->g_full[b]);
      if (dbg && warp == 2 && lane == 0 && j < 8) p.dbg[32 + 2 * j + 1] = clock64() - tk_start;
    }
  } else if (warp < 18) {
    // =================================== E2: residual (8 warps) ===================================
    // The fp32 residual rows of the own tiles live in TMEM (L[k]); MMA2 accumulates g.Wr into them.
    // E2 adds the running sum of the residual biases (kept out of TMEM so nothing is written back) and
    // refreshes the tile's fp16 hi / lo planes in shared memory (the next layer's MMA operand, and what the publisher stores).
    const int half = (warp - 10) >> 2;
    const int qd = warp & 3;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    const int row = qd * 32 + lane;  // row inside a tile
    uint32_t rmx = 0;                // fp16-range guard: largest magnitude this thread split (start conv)
    float rmxf = 0.f;                // the same for the per-layer residual rows (one FMNMX per value)
    if (p.fuse_start) {
      // init: start conv of the own tiles (-> TMEM master + smem planes) and of the halo (smem only;
      // zeros at the start of a clip: the layer input is zero-padded, masked.py:203-204)
      const float4* w4 = reinterpret_cast<const float4*>(p.start_w) + half * 8;
      const float4* b4 = reinterpret_cast<const float4*>(p.start_b) + half * 8;
      for (int k = K - 1; k >= -1; --k) {
        const int t = (R.tk0 + k) * BM + row;  // time index inside the clip (negative only for the halo of tile 0)
        const float* xr = p.start_x + (size_t)gclip * p.T + t;
        const float x1 = t >= 1 ? __ldg(xr - 1) : 0.f, x2 = t >= 2 ? __ldg(xr - 2) : 0.f, x3 = t >= 3 ? __ldg(xr - 3) : 0.f;
        const uint32_t cur_lo = sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
        const uint32_t cur_hi = sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w0 = __ldg(w4 + i), w1 = __ldg(w4 + 16 + i), w2 = __ldg(w4 + 32 + i), bb = __ldg(b4 + i);
          float o[4];
          o[0] = fmaf(w2.x, x1, fmaf(w1.x, x2, fmaf(w0.x, x3, bb.x)));
          o[1] = fmaf(w2.y, x1, fmaf(w1.y, x2, fmaf(w0.y, x3, bb.y)));
          o[2] = fmaf(w2.z, x1, fmaf(w1.z, x2, fmaf(w0.z, x3, bb.z)));
          o[3] = fmaf(w2.w, x1, fmaf(w1.w, x2, fmaf(w0.w, x3, bb.w)));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[4 * i + e] = __float_as_uint(t >= 0 ? o[e] : 0.f);
            range_track(rmx, __uint_as_float(v[4 * i + e]));
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t coff = (uint32_t)(((4 * half + jj) ^ (row & 7)) * 16);
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float o0 = __uint_as_float(v[8 * jj + 2 * e]), o1 = __uint_as_float(v[8 * jj + 2 * e + 1]);
            const float a0 = __half2float(__float2half_rn(o0));
            const float a1 = __half2float(__float2half_rn(o1));
            hw[e] = ft_pack_f16(a0, a1);
            lw[e] = ft_pack_f16(o0 - a0, o1 - a1);
          }
          ft_sts128(cur_hi + coff, make_uint4(hw[0], hw[1], hw[2], hw[3]));
          ft_sts128(cur_lo + coff, make_uint4(lw[0], lw[1], lw[2], lw[3]));
        }
        if (k >= 0) ft_tmem_st32(tmem + lane_sel + TM_L + k * 64 + half * 32, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) ft_arrive(k >= 0 ? &B->own_loaded[k] : &B->halo0);
      }
    } else {
    // init: L[k] = hi + lo of the start-conv output the loader fetched
    for (int k = K - 1; k >= 0; --k) {
      ft_wait(&B->own_loaded[k], 0, "own_loaded (E2)");
      const uint32_t cur_lo = sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      const uint32_t cur_hi = sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      uint32_t v[32];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t coff = (uint32_t)(((4 * half + jj) ^ (row & 7)) * 16);
        const uint4 hraw = ft_lds128(cur_hi + coff), lraw = ft_lds128(cur_lo + coff);
        const __half2* hp = reinterpret_cast<const __half2*>(&hraw);
        const __half2* lp = reinterpret_cast<const __half2*>(&lraw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 hv = __half22float2(hp[e]), lv = __half22float2(lp[e]);
          v[8 * jj + 2 * e] = __float_as_uint(hv.x + lv.x);
          v[8 * jj + 2 * e + 1] = __float_as_uint(hv.y + lv.y);
        }
      }
      ft_tmem_st32(tmem + lane_sel + TM_L + k * 64 + half * 32, v);
    }
    }
    ft_tmem_st_wait();
    ft_fence_before();
    __syncwarp();
    if (lane == 0) ft_arrive(&B->l_init);
    uint32_t npub = 0;  // publishes so far per own tile (8 bits each)
    int li = 0, kk = 0;
    for (int j = 0; j < total; ++j) {
      const int b = j & 1;
      const int k = K - 1 - kk;
      // TMEM holds l without the residual biases; p.br is their running sum over the layers (the sum up to
      // layer l0-1 is already inside the rows this launch started from)
      const float4* bptr = reinterpret_cast<const float4*>(p.br + (size_t)(p.l0 + li) * C) + half * 8;
      float4 bb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bb[i] = __ldg(bptr + i);
      if (p.l0 > 0) {
        const float4* b0 = reinterpret_cast<const float4*>(p.br + (size_t)(p.l0 - 1) * C) + half * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = __ldg(b0 + i);
          bb[i].x -= q.x; bb[i].y -= q.y; bb[i].z -= q.z; bb[i].w -= q.w;
        }
      }
      ft_wait(&B->d2_full[k], (uint32_t)(li & 1), "d2_full");
      if (dbg && warp == 10 && lane == 0 && j < 8) p.dbg[48 + 2 * j] = clock64() - tk_start;
      ft_fence_after();
      uint32_t d[32];
      ft_tmem_ld32(tmem + lane_sel + TM_L + k * 64 + half * 32, d);
      ft_tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d[4 * i] = __float_as_uint(__uint_as_float(d[4 * i]) + bb[i].x);
        d[4 * i + 1] = __float_as_uint(__uint_as_float(d[4 * i + 1]) + bb[i].y);
        d[4 * i + 2] = __float_as_uint(__uint_as_float(d[4 * i + 2]) + bb[i].z);
        d[4 * i + 3] = __float_as_uint(__uint_as_float(d[4 * i + 3]) + bb[i].w);
      }
      // the TMA store of this tile's previous published value must have finished reading it
      const uint32_t np = (npub >> (8 * k)) & 0xffu;
      if (np > 0) ft_wait(&B->pub_done[k], (np - 1) & 1, "pub_done");
      const uint32_t cur_lo = sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      const uint32_t cur_hi = sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      const bool relu_out = p.fuse_head && li == nl - 1;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t coff = (uint32_t)(((4 * half + jj) ^ (row & 7)) * 16);
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float o0 = __uint_as_float(d[8 * jj + 2 * e]), o1 = __uint_as_float(d[8 * jj + 2 * e + 1]);
          range_track_fast(rmxf, o0);
          range_track_fast(rmxf, o1);
          if (relu_out) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }  // the head's out1 consumes relu(l)
          const float a0 = __half2float(__float2half_rn(o0));
          const float a1 = __half2float(__float2half_rn(o1));
          hw[e] = ft_pack_f16(a0, a1);
          lw[e] = ft_pack_f16(o0 - a0, o1 - a1);
        }
        ft_sts128(cur_hi + coff, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        ft_sts128(cur_lo + coff, make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
      ft_fence_before();
      // generic-proxy writes to shared memory -> tcgen05.mma / TMA store reads (async proxy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ft_arrive(&B->tile_ready[k]);
      if (published(k, li)) npub += 1u << (8 * k);
      if (dbg && warp == 10 && lane == 0 && j < 8) p.dbg[48 + 2 * j + 1] = clock64() - tk_start;
      if (++kk == K) { kk = 0; ++li; }
    }
    range_commit(rmx);
    range_commit_fast(rmxf);
  } else if (warp == 18) {
    // =================================== publisher ===================================
    if (lane == 0) {
      // CTAs of this clip that may read what this CTA publishes: those owning the next reach_tiles tiles
      const int last_tk = R.tk0 + K - 1;
      const int far_tk = min(last_tk + p.reach_tiles, p.tiles_per_clip - 1);
      const int far_idx = ft_owner(p.tiles_per_clip, R.n_c, far_tk);
      const int cta0 = (int)blockIdx.x - R.idx;  // first CTA of this clip
      for (int li = 0; li < nl; ++li) {
        bool checked = false;
        const int wb = (p.buf0 + li + 1) & 1;
        const CUtensorMap* mh = wb ? &map_h1 : &map_h0;
        const CUtensorMap* ml = wb ? &map_l1 : &map_l0;
        for (int k = K - 1; k >= 0; --k) {
          if (!published(k, li)) continue;
          ft_wait(&B->tile_ready[k], (uint32_t)(li & 1), "tile_ready (publisher)");
          if (!checked && li >= 1) {
            // the buffer written now holds layer li-2's output, read by the neighbours during layer li-1
            for (int i = R.idx + 1; i <= far_idx; ++i) ft_poll_ge(p.cons + cta0 + i, (unsigned int)li, "cons poll");
            checked = true;
          }
          ft_tma_store_3d(mh, sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B, 0, (R.tk0 + k) * BM, gclip);
          ft_tma_store_3d(ml, sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B, 0, (R.tk0 + k) * BM, gclip);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          ft_arrive(&B->pub_done[k]);  // shared memory may be overwritten by the next layer
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          ft_fence_async();
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + gt0 + k), "r"((unsigned int)(li + 1))
                       : "memory");
        }
      }
    }
  }

  ft_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) {
    p.dbg[127] = clock64() - tk_start;
    unsigned long long gt_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_end));
    p.dbg[126] = (long long)(gt_end - gt_start);  // ns
  }
  if (warp == 1) {
    ft_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace
}  // namespace nsw
NSW_RANGE_GUARD_TU(flow_tc)
namespace nsw {

// how many clips one launch can take: every clip needs ceil(tiles / KMAX) CTAs of its own
int flow_tc_clips_per_launch(int T, int num_sms) {
  if (T % BM != 0 || T <= 0) return 0;
  const int tiles = T / BM;
  const int need = (tiles + KMAX - 1) / KMAX;
  return num_sms / need;
}

// CTAs per clip for a launch: clip c gets cta_base + (c < cta_rem)
static void flow_tc_split(int tiles, int nclips, int num_sms, int* cta_base, int* cta_rem) {
  const int per_clip = std::min(num_sms / nclips, tiles);
  *cta_base = per_clip;
  *cta_rem = per_clip < tiles ? std::min(num_sms - per_clip * nclips, nclips) : 0;
}

int flow_tc_launch(const void* const map_act[2][2], const void* map_wdh, const void* map_wdl,
                   const void* map_wrh, const void* map_wrl, const float* cond_tiled, size_t cond_plane,
                   const float* br, int T, int clip0, int nclips, int buf0, int l0, int l1, int num_stages,
                   unsigned int* sync_words /* >= nclips*T/128 + num_sms */, int num_sms, const FlowHead* head,
                   const FlowStart* start, cudaStream_t stream) {
  static std::atomic<uint64_t> attr_done{0};  // per device (the attribute is per device)
  NSW_TRY(ensure_dynamic_smem((const void*)iaf_flow_tc_kernel, (int)FT_SMEM_BYTES, attr_done));
  NSW_CHECK(T % BM == 0, NSW_EINVAL, "flow_tc: T=%d must be a multiple of %d", T, BM);
  NSW_CHECK(l1 > l0 && nclips >= 1, NSW_EINVAL, "flow_tc: empty layer or clip range");
  const int tiles = T / BM;
  NSW_CHECK(nclips <= flow_tc_clips_per_launch(T, num_sms), NSW_EINVAL,
            "flow_tc: %d clips of %d tiles do not fit %d SMs", nclips, tiles, num_sms);
  FlowTcParams p;
  p.cond = cond_tiled;
  p.cond_plane = cond_plane;
  p.br = br;
  p.buf0 = buf0;
  p.l0 = l0;
  p.l1 = l1;
  p.num_stages = num_stages;
  p.tiles_per_clip = tiles;
  p.clip0 = clip0;
  p.nclips = nclips;
  flow_tc_split(tiles, nclips, num_sms, &p.cta_base, &p.cta_rem);
  // the remainder CTAs only help if they lower some clip's tiles per CTA; harmless otherwise
  const int grid = p.cta_base * nclips + p.cta_rem;
  p.reach_tiles = std::max(1, 2 * (1 << (num_stages - 1)) / BM);
  // flags are indexed by global tile id (clip coordinate of the tensor map), cons by CTA
  const size_t n_flags = (size_t)(clip0 + nclips) * tiles;
  p.flags = sync_words;
  p.cons = sync_words + n_flags;
  NSW_CUDA(cudaMemsetAsync(sync_words, 0, (n_flags + grid) * sizeof(unsigned int), stream));
  p.fuse_head = head != nullptr;
  if (head) {
    p.head_w_tile = head->w_tile;
    p.wm = head->wm; p.ws = head->ws; p.bm = head->bm; p.bs = head->bs;
    p.x_in = head->x_in; p.z = head->z; p.x_out = head->x_out;
    p.mean_tot = head->mean_tot; p.scale_tot = head->scale_tot; p.log_scale_tot = head->log_scale_tot;
    p.first = head->first; p.last = head->last; p.quantize = head->quantize; p.use_mu_law = head->use_mu_law;
    p.quant_chann = head->quant_chann;
  } else {
    p.head_w_tile = 0;
    p.wm = p.ws = p.x_in = p.z = nullptr;
    p.x_out = p.mean_tot = p.scale_tot = p.log_scale_tot = nullptr;
    p.bm = p.bs = p.quant_chann = 0.f;
    p.first = p.last = p.quantize = p.use_mu_law = 0;
  }
  p.T = T;
  p.fuse_start = (start != nullptr && l0 == 0) ? 1 : 0;
  p.start_x = start ? start->x : nullptr;
  p.start_w = start ? start->w : nullptr;
  p.start_b = start ? start->b : nullptr;
  p.dbg = nullptr;
  p.dbg_cta = getenv("NSW_LAYER_DEBUG_CTA") ? atoi(getenv("NSW_LAYER_DEBUG_CTA")) : 0;
  static long long* dbg_bufs[64] = {nullptr};  // NSW_LAYER_DEBUG only: one buffer per device
  const bool want_dbg = getenv("NSW_LAYER_DEBUG") != nullptr;
  long long* dbg_buf = nullptr;
  if (want_dbg) {
    int dbg_dev = 0;
    NSW_CUDA(cudaGetDevice(&dbg_dev));
    if (!dbg_bufs[dbg_dev & 63]) NSW_CUDA(cudaMalloc(&dbg_bufs[dbg_dev & 63], 256 * sizeof(long long)));
    dbg_buf = dbg_bufs[dbg_dev & 63];
    NSW_CUDA(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), stream));
    p.dbg = dbg_buf;
  }
  void* args[] = {const_cast<void*>(map_act[0][0]), const_cast<void*>(map_act[0][1]),
                  const_cast<void*>(map_act[1][0]), const_cast<void*>(map_act[1][1]),
                  const_cast<void*>(map_wdh), const_cast<void*>(map_wdl), const_cast<void*>(map_wrh),
                  const_cast<void*>(map_wrl), &p};
  // cooperative launch: the per-tile flags need every CTA to be co-resident
  static const bool no_coop = getenv("NSW_FLOW_NOCOOP") != nullptr;  // timing experiment
  if (no_coop)
    NSW_CUDA(cudaLaunchKernel((void*)iaf_flow_tc_kernel, dim3(grid), dim3(FT_THREADS), args, FT_SMEM_BYTES, stream));
  else
    NSW_CUDA(cudaLaunchCooperativeKernel((void*)iaf_flow_tc_kernel, dim3(grid), dim3(FT_THREADS), args,
                                         FT_SMEM_BYTES, stream));
  count_launch();
  NSW_CUDA(cudaGetLastError());
  if (want_dbg) {
    long long h[256];
    NSW_CUDA(cudaStreamSynchronize(stream));
    NSW_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[flow_tc dbg layers %d..%d grid %d] end %lld cycles in %lld ns = %.0f MHz\n", l0, l1, grid, h[127], h[126],
            h[126] > 0 ? 1e3 * (double)h[127] / (double)h[126] : 0.0);
    const char* names[4] = {"MMA1 (operands ready, issued)", "MMA2 (g ready, issued)", "E1   (D1 ready, g stored)",
                            "E2   (D2 ready, tile stored)"};
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, "  %-30s:", names[r]);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld-%lld", h[16 * r + 2 * i], h[16 * r + 2 * i + 1]);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "  %-30s:", "MMA1 D1 buffer free at");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld", h[64 + 2 * i]);
    fprintf(stderr, "\n");
    fprintf(stderr, "  %-30s:", "layer start (MMA1 sees weights)");
    for (int i = 0; i < std::min(l1 - l0, 32); ++i) fprintf(stderr, " %lld", h[80 + i]);
    fprintf(stderr, "  [cta %d]\n", p.dbg_cta);
    const char* wn[3] = {"MMA1 waits: D1 buffer", "MMA1 waits: operand tiles/halo", "MMA1 waits: foreign ring"};
    for (int r = 0; r < 3; ++r) {
      fprintf(stderr, "  %-30s:", wn[r]);
      for (int i = 0; i < std::min(l1 - l0, 32); ++i) fprintf(stderr, " %lld", h[128 + 32 * r + i]);
      fprintf(stderr, "\n");
    }
  }
  return NSW_OK;
}

}  // namespace nsw

// TEST HOOK (host only, no CUDA): the work split and the publish / read plan of one flow-kernel launch, produced by the
// same integer functions the kernel uses.  out[] receives, per CTA, a header {clip, tk0, K, n_c} followed by
// (l1 - l0) * K records {published, source of tap t-2d, source of tap t-d} (see ft_tap_source); *n_out = ints written
// (or needed, if cap is too small); returns the grid size, or a negative NSW_E* code.
extern "C" int nsw_flow_plan_host(int32_t T, int32_t nclips, int32_t num_sms, int32_t l0, int32_t l1, int32_t num_stages,
                                  int32_t fuse_head, int32_t* out, int64_t cap, int64_t* n_out) {
  using namespace nsw;
  NSW_CHECK(T > 0 && T % BM == 0 && nclips >= 1 && l1 > l0 && num_stages >= 1 && n_out, NSW_EINVAL,
            "nsw_flow_plan_host: bad argument");
  NSW_CHECK(nclips <= flow_tc_clips_per_launch(T, num_sms), NSW_EINVAL, "nsw_flow_plan_host: clips do not fit one launch");
  const int tiles = T / BM, nl = l1 - l0;
  int cta_base, cta_rem;
  flow_tc_split(tiles, nclips, num_sms, &cta_base, &cta_rem);
  const int grid = cta_base * nclips + cta_rem;
  int64_t n = 0;
  auto put = [&](int v) { if (out && n < cap) out[n] = v; ++n; };
  for (int b = 0; b < grid; ++b) {
    const Range r = ft_range_of(tiles, cta_base, cta_rem, b);
    put(r.clip); put(r.tk0); put(r.K); put(r.n_c);
    for (int li = 0; li < nl; ++li) {
      const int d = 1 << ((l0 + li) % num_stages);
      for (int k = 0; k < r.K; ++k) {
        put(ft_published(k, r.K, li, nl, l0, num_stages, fuse_head) ? 1 : 0);
        put(ft_tap_source(r.tk0, k, 0, d));
        put(ft_tap_source(r.tk0, k, 1, d));
      }
    }
  }
  *n_out = n;
  return grid;
}

