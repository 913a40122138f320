// Engine "tc3", CTA-PAIR variant of the persistent IAF flow kernel (nsw_iaf_flow_tc.cu: read that header first).
//
// Two CTAs of a cluster (rank 0 = leader) own 2K consecutive 128-row tiles of one clip -- the leader the first K, the
// peer the next K -- and walk the same task sequence (layer-major, own tile K-1 .. 0) IN LOCK STEP: every tcgen05.mma is
// issued once, by the leader, with cta_group::2 (M = 256: 128 rows of each CTA), so each SM holds only HALF of every
// weight tile (32 of the 64 output channels: the tensor cores of the pair exchange the B operand).  That
//   * takes 1 KB of the 6 KB of shared-memory operand fetch out of every MMA, and
//   * halves the weight bytes per SM, which makes room to DOUBLE-BUFFER the dilated-conv weights (2 x 24 KB instead of
//     1 x 48 KB): the single-CTA kernel stalls ~3 000-4 000 of ~17 500 cycles per layer between the last burst of a
//     layer and the first of the next, waiting for the reload of its only weight set (profiles/r02).
// Everything that is per-SM in the single-CTA kernel stays per-SM (loader, gate warps, residual warps, publisher,
// resident planes, TMEM layout); what changes is who waits for whom:
//   - barriers the issuers wait on live in the LEADER and count both CTAs (the peer's warps arrive remotely, the
//     peer's TMA loads complete on the leader's barrier);
//   - tcgen05.commit is multicast, so barriers the epilogue / loader warps wait on are per CTA as before.
// Product path only: start conv and head fused (l0 == 0, whole flow in one launch); debug taps, odd tile counts and
// the long-clip fallback use the single-CTA kernel.  Outputs are bit-identical to the single-CTA kernel's (same products
// in the same order; tests/test_iaf_tc_gpu.py).  Measured at 8 x 7680 (profiles/r02): issue of a 36-MMA burst 1 300
// instead of 2 700 cycles, layers 0.752 vs 0.776 ms, step 1.142 vs 1.170 ms.  Remote arrives use CTA-scope release: with
// mbarrier.arrive.release.cluster the peer's gate warps lost ~0.5 us per arrive and the pair ran at the peer's pace.
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
constexpr uint32_t WD_TILE = 32 * 64 * 2;              // this CTA's 32 of the 64 weight rows: 4 KB per tap per plane
constexpr uint32_t WR_TILE = 32 * 32 * 2;              // 2 KB per plane
constexpr uint32_t OFF_HI = 0, OFF_LO = PLANE_B;
constexpr uint32_t OFF_WD = 2 * PLANE_B;               // 2 buffers x [3 taps hi | 3 taps lo]
__host__ __device__ constexpr uint32_t wdh_off(int buf, int tap) { return OFF_WD + (uint32_t)(buf * 6 + tap) * WD_TILE; }
__host__ __device__ constexpr uint32_t wdl_off(int buf, int tap) { return OFF_WD + (uint32_t)(buf * 6 + 3 + tap) * WD_TILE; }
constexpr uint32_t OFF_WR = OFF_WD + 12 * WD_TILE;     // 2 buffers x [hi 2 KB][lo 2 KB]
constexpr uint32_t OFF_BARS = OFF_WR + 4 * WR_TILE;
constexpr size_t FP_SMEM_BYTES = OFF_BARS + 1024 + 1024;
constexpr long long FT_WATCHDOG = 4000000000ll;
// TMEM columns (per CTA, as in the single-CTA kernel): D1[2] 0..127, G[2] 128..191, L[4] 192..447
constexpr uint32_t TM_G = 128, TM_L = 192;

struct FpBars {
  // leader-only (waited on by the issuers, count both CTAs)
  uint64_t own_loaded[KMAX];   // start conv of own tile k done (16 residual warps)
  uint64_t tile_ready_pair[KMAX];
  uint64_t halo0, l_init;
  uint64_t ring_full;          // tx bytes of both CTAs' halo / foreign loads
  uint64_t wd_full[2], wr_full[2];
  uint64_t d1_empty[2], g_full[2];
  // per CTA
  uint64_t tile_ready[KMAX];   // local residual warps -> local publisher
  uint64_t pub_done[KMAX];
  uint64_t ring_free, wd_free[2], wr_free[2];
  uint64_t d1_full[2], g_free[2];
  uint64_t d2_full[KMAX];
  uint32_t tmem_base;
};
static_assert(sizeof(FpBars) <= 1024, "barrier block grew");

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
  printf("nsw iaf_flow_pair: watchdog in %s (block %d thread %d)\n", what, blockIdx.x, threadIdx.x);
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


// ---- cluster-pair variants ----
__device__ __forceinline__ uint32_t fp_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void fp_arrive_leader(uint64_t* bar, bool leader) {  // arrive on the LEADER's barrier at this offset
  if (leader) ft_arrive(bar);
  // (default semantics = release at CTA scope: what travels between the SMs is TMEM / shared memory that only the
  //  owning SM's tensor core reads, ordered by tcgen05.fence / fence.proxy.async; a cluster-scope release costs the
  //  peer's gate warps ~0.5 us per arrive -- measured: the peer ran 2 us behind the leader on every task)
  else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(fp_mapa(smem_u32(bar), 0)) : "memory");
}
__device__ __forceinline__ void fp_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                               int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fp_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// D = f32, A = B = f16, K-major, M = 256 (128 per CTA), N = 64
__device__ __forceinline__ uint32_t fp_idesc() {
  return (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void fp_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void fp_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void fp_commit(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void fp_commit_leader(uint64_t* bar) {  // leader's barrier only
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)1)
      : "memory");
}

struct FlowPairParams {
  const float* cond;      // tiled conditioning plane of layer 0 of this flow
  size_t cond_plane;      // floats between consecutive layers' planes
  const float* br;        // [L][64] RUNNING SUM over layers of the residual biases
  unsigned int* flags;    // [clips * tiles_per_clip] layers published per tile   (zeroed per launch)
  unsigned int* cons;     // [grid] layers whose foreign loads a CTA has finished (zeroed per launch)
  int buf0, nl, num_stages;
  int tiles_per_clip, clip0, nclips;
  int pairs_per_clip;     // every clip is split over this many CTA pairs
  int reach_tiles;
  int head_w_tile;        // 64-row tile index of W1^T in the Wd tensor maps
  const float* wm;
  const float* ws;
  float bm, bs;
  const float* x_in;
  const float* z;
  float* x_out;
  float* mean_tot;
  float* scale_tot;
  float* log_scale_tot;
  int T;
  int first, last, quantize, use_mu_law;
  float quant_chann;
  const float* start_w;   // [3][64]
  const float* start_b;   // [64]
  long long* dbg;         // NSW_FLOW_PAIR_DEBUG: where the MMA1 issuer of pair dbg_pair waited (cycles, summed over the flow)
  int dbg_pair;
};

// ---- work split (plain integer functions) ----
// A clip's tiles come in "double tiles" (two consecutive tiles); pair idx of P owns K double tiles starting at h0, i.e.
// tiles [2 h0, 2 h0 + 2K): the leader the first K, the peer the next K.
struct PairRange {
  int clip, idx, K, h0;
};
__host__ __device__ __forceinline__ PairRange fp_range_of(int tiles_per_clip, int P, int pair) {
  PairRange r;
  r.clip = pair / P;
  r.idx = pair - r.clip * P;
  const int half = tiles_per_clip / 2;
  const int bt = half / P, rt = half - bt * P;
  r.K = bt + (r.idx < rt ? 1 : 0);
  r.h0 = r.idx * bt + (r.idx < rt ? r.idx : rt);
  return r;
}
// CTA (index within its clip: 2 * pair + rank) that owns tile tk
__host__ __device__ __forceinline__ int fp_owner(int tiles_per_clip, int P, int tk) {
  const int half = tiles_per_clip / 2;
  const int bt = half / P, rt = half - bt * P;
  const int hk = tk >> 1;
  const int big = rt * (bt + 1);
  const int pidx = hk < big ? hk / (bt + 1) : rt + (hk - big) / bt;
  const int K = bt + (pidx < rt ? 1 : 0);
  const int h0 = pidx * bt + (pidx < rt ? pidx : rt);
  return 2 * pidx + ((tk - 2 * h0) >= K ? 1 : 0);
}
__host__ __device__ __forceinline__ bool fp_published(int k, int K, int li, int nl, int num_stages) {
  if (li == nl - 1) return false;                              // the fused head consumes the last layer in place
  if ((1 << ((li + 1) % num_stages)) >= BM) return true;       // next layer reads whole foreign tiles
  return k == K - 1;                                           // halo of the next CTA
}

__global__ void __launch_bounds__(FT_THREADS, 1)
iaf_flow_pair_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_l0,
                     const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_l1,
                     const __grid_constant__ CUtensorMap map_wdh, const __grid_constant__ CUtensorMap map_wdl,
                     const __grid_constant__ CUtensorMap map_wrh, const __grid_constant__ CUtensorMap map_wrl,
                     FlowPairParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FpBars* B = reinterpret_cast<FpBars*>(smem + OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const int crank = (int)(blockIdx.x & 1);
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int k = 0; k < KMAX; ++k) {
      ft_mbar_init(&B->own_loaded[k], 16);
      ft_mbar_init(&B->tile_ready_pair[k], 16);
      ft_mbar_init(&B->tile_ready[k], 8);
      ft_mbar_init(&B->pub_done[k], 1);
      ft_mbar_init(&B->d2_full[k], 1);
    }
    ft_mbar_init(&B->halo0, 16);
    ft_mbar_init(&B->l_init, 16);
    ft_mbar_init(&B->ring_full, 1);
    ft_mbar_init(&B->ring_free, 1);
    for (int b = 0; b < 2; ++b) {
      ft_mbar_init(&B->wd_full[b], 1);
      ft_mbar_init(&B->wd_free[b], 1);
      ft_mbar_init(&B->wr_full[b], 1);
      ft_mbar_init(&B->wr_free[b], 1);
      ft_mbar_init(&B->d1_full[b], 1);
      ft_mbar_init(&B->d1_empty[b], 16);
      ft_mbar_init(&B->g_full[b], 16);
      ft_mbar_init(&B->g_free[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ft_fence_before();
  __syncthreads();
  // barriers of both CTAs must exist before the peer's TMA completes on / arrives at them
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  ft_fence_after();
  const uint32_t tmem = B->tmem_base;

  const PairRange R = fp_range_of(p.tiles_per_clip, p.pairs_per_clip, (int)blockIdx.x >> 1);
  const int K = R.K;
  const int tk0 = 2 * R.h0 + crank * K;  // first own tile of THIS CTA (index inside the clip)
  const int nl = p.nl;
  const int total = nl * K;              // gate / residual tasks
  const int total1 = (nl + 1) * K;       // MMA1 / gate tasks incl. the head pseudo-layer
  const int gclip = p.clip0 + R.clip;
  const int gt0 = gclip * p.tiles_per_clip + tk0;

  const long long tk_start = clock64();
  const bool dbgt = p.dbg != nullptr && (int)(blockIdx.x >> 1) == p.dbg_pair && leader;  // timeline of tasks J0 .. J0+7
  constexpr int J0 = 8;
  auto published = [&](int k, int li) -> bool { return fp_published(k, K, li, nl, p.num_stages); };
  // foreign whole-tile tap (d >= 128) of own tile k: skipped by BOTH CTAs only if even the peer's source is before the clip
  auto tap_skipped = [&](int k, int tap, int dt) -> bool { return 2 * R.h0 + K + k - (2 - tap) * dt < 0; };

  if (warp == 0) {
    // =================================== loader (each CTA its own halves / tiles) ===================================
    if (lane == 0) {
      auto load_wd = [&](int layer, int buf) {
        const uint32_t bar = fp_mapa(smem_u32(&B->wd_full[buf]), 0);
        if (leader) ft_expect_tx(&B->wd_full[buf], 2 * 6 * WD_TILE);
        for (int tap = 0; tap < 3; ++tap) {
          fp_tma_load_2d(sbase + wdh_off(buf, tap), &map_wdh, bar, 0, (layer * 3 + tap) * 64 + crank * 32);
          fp_tma_load_2d(sbase + wdl_off(buf, tap), &map_wdl, bar, 0, (layer * 3 + tap) * 64 + crank * 32);
        }
      };
      auto load_head_w = [&](int buf) {  // W1^T into the tap-2 slot of `buf`
        const uint32_t bar = fp_mapa(smem_u32(&B->wd_full[buf]), 0);
        if (leader) ft_expect_tx(&B->wd_full[buf], 2 * 2 * WD_TILE);
        fp_tma_load_2d(sbase + wdh_off(buf, 2), &map_wdh, bar, 0, p.head_w_tile * 64 + crank * 32);
        fp_tma_load_2d(sbase + wdl_off(buf, 2), &map_wdl, bar, 0, p.head_w_tile * 64 + crank * 32);
      };
      auto load_wr = [&](int layer, int b) {
        const uint32_t bar = fp_mapa(smem_u32(&B->wr_full[b]), 0);
        if (leader) ft_expect_tx(&B->wr_full[b], 2 * 2 * WR_TILE);
        fp_tma_load_2d(sbase + OFF_WR + (2 * b) * WR_TILE, &map_wrh, bar, 0, layer * 64 + crank * 32);
        fp_tma_load_2d(sbase + OFF_WR + (2 * b + 1) * WR_TILE, &map_wrl, bar, 0, layer * 64 + crank * 32);
      };
      auto prefetch_cond = [&](int li) {
        if (li > nl) return;
        const float* c = p.cond + (size_t)li * p.cond_plane + (size_t)gt0 * (BM * C);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c), "r"((uint32_t)(K * BM * C * 4)) : "memory");
      };
      load_wd(0, 0);
      if (nl > 1) load_wd(1, 1); else load_head_w(1);
      load_wr(0, 0);
      if (nl > 1) load_wr(1, 1);
      prefetch_cond(0);
      prefetch_cond(1);
      int n_ring = 0;
      const uint32_t ring_bar = fp_mapa(smem_u32(&B->ring_full), 0);
      auto foreign = [&](int li, int src_tk, const CUtensorMap* mh, const CUtensorMap* ml) {
        // src_tk may be negative (before the clip start): TMA zero-fills the out-of-range rows
        if (li > 0 && src_tk >= 0)
          ft_poll_ge(p.flags + (size_t)gclip * p.tiles_per_clip + src_tk, (unsigned int)li, "flag poll");
        if (n_ring > 0) ft_wait(&B->ring_free, (uint32_t)((n_ring - 1) & 1), "ring_free");
        ft_fence_async();
        if (leader) ft_expect_tx(&B->ring_full, 2 * 2 * TILE_B);
        fp_tma_load_3d(sbase + OFF_LO, ml, ring_bar, 0, src_tk * BM, gclip);
        fp_tma_load_3d(sbase + OFF_HI, mh, ring_bar, 0, src_tk * BM, gclip);
        ++n_ring;
      };
      for (int li = 0; li < nl; ++li) {
        const int d = 1 << (li % p.num_stages);
        if (li > 0) {
          // weights two layers ahead into the buffer layer li-1 has just released (a whole layer of lead time)
          ft_wait(&B->wd_free[(li + 1) & 1], (uint32_t)(((li - 1) >> 1) & 1), "wd_free");
          if (li + 1 < nl) load_wd(li + 1, (li + 1) & 1);
          else load_head_w((li + 1) & 1);
          if (li >= 2) {  // (the residual weights follow one layer later: their buffer was released by layer li-2)
            ft_wait(&B->wr_free[li & 1], (uint32_t)(((li >> 1) - 1) & 1), "wr_free");
            load_wr(li, li & 1);
          }
          prefetch_cond(li + 1);
        }
        const int rb = (p.buf0 + li) & 1;  // global buffer holding the previous layer's output
        const CUtensorMap* mh = rb ? &map_h1 : &map_h0;
        const CUtensorMap* ml = rb ? &map_l1 : &map_l0;
        if (2 * d <= BM) {
          if (li > 0) foreign(li, tk0 - 1, mh, ml);  // (layer 0: the residual warps compute the halo themselves)
        } else {
          const int dt = d / BM;
          for (int k = K - 1; k >= 0; --k)
            for (int tap = 1; tap >= 0; --tap) {
              if (k - (2 - tap) * dt >= 0) continue;  // own tile
              if (tap_skipped(k, tap, dt)) continue;  // causal zeros for both CTAs
              foreign(li, tk0 + k - (2 - tap) * dt, mh, ml);
            }
        }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA1 issuer (leader) ===================================
    if (leader) {
      const uint32_t idesc = fp_idesc();
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
      auto wait_ready = [&](int x, int li) {  // own tile x of BOTH CTAs holds layer li-1's output
        if (li == 0) ft_wait(&B->own_loaded[x], 0, "own_loaded");
        else ft_wait(&B->tile_ready_pair[x], (uint32_t)((li - 1) & 1), "tile_ready (mma)");
      };
      auto issue_tap = [&](uint32_t d1, uint32_t a_row_bytes, int buf, int tap) {
        const uint64_t alo = ft_desc_sw128(sbase_u + OFF_LO + a_row_bytes);
        const uint64_t ahi = ft_desc_sw128(sbase_u + OFF_HI + a_row_bytes);
        const uint64_t wh = ft_desc_sw128(sbase_u + wdh_off(buf, tap));
        const uint64_t wl = ft_desc_sw128(sbase_u + wdl_off(buf, tap));
        if (ft_elect()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) fp_mma_ss(d1, alo + 2 * k, wh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            fp_mma_ss(d1, ahi + 2 * k, wl + 2 * k, idesc, 1);
            fp_mma_ss(d1, ahi + 2 * k, wh + 2 * k, idesc, 1);
          }
        }
        __syncwarp();
      };
      auto issue_burst = [&](uint32_t d1, const uint32_t (&a_bytes)[3], const int (&taps)[3], int n, int buf) {
        uint64_t alo[3], ahi[3], wh[3], wl[3];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          alo[s2] = ft_desc_sw128(sbase_u + OFF_LO + a_bytes[s2]);
          ahi[s2] = ft_desc_sw128(sbase_u + OFF_HI + a_bytes[s2]);
          wh[s2] = ft_desc_sw128(sbase_u + wdh_off(buf, taps[s2]));
          wl[s2] = ft_desc_sw128(sbase_u + wdl_off(buf, taps[s2]));
        }
        if (ft_elect()) {
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            if (s2 < n) {
#pragma unroll
              for (int k = 0; k < 4; ++k) fp_mma_ss(d1, alo[s2] + 2 * k, wh[s2] + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                fp_mma_ss(d1, ahi[s2] + 2 * k, wl[s2] + 2 * k, idesc, 1);
                fp_mma_ss(d1, ahi[s2] + 2 * k, wh[s2] + 2 * k, idesc, 1);
              }
            }
          }
        }
        __syncwarp();
      };
      int n_ring = 0, j = 0;
      const bool dbg = p.dbg != nullptr && (int)(blockIdx.x >> 1) == p.dbg_pair;
      long long w_wd = 0, w_d1 = 0, w_rdy = 0, w_ring = 0, t_issue = 0;
      const long long t_begin = clock64();
      for (int li = 0; li < nl; ++li) {
        const int d = 1 << (li % p.num_stages);
        const bool small = 2 * d <= BM;
        const int dt = d / BM;
        const int buf = li & 1;
        long long tw0 = dbg ? clock64() : 0;
        ft_wait(&B->wd_full[buf], (uint32_t)((li >> 1) & 1), "wd_full");
        if (dbg) w_wd += clock64() - tw0;
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int k = K - 1 - kk;
          const int b = j & 1;
          const uint32_t d1 = tmem_u + b * 64;
          uint32_t a_bytes[3] = {(uint32_t)(1 + k) * TILE_B, 0u, 0u};
          int taps[3] = {2, 1, 0};
          int n = 1;
          bool foreign[2] = {false, false};
          bool halo_used = false;
          tw0 = dbg ? clock64() : 0;
          ft_wait(&B->d1_empty[b], (uint32_t)((j >> 1) & 1), "d1_empty");
          if (dbg) { const long long now = clock64(); w_d1 += now - tw0; tw0 = now; }
          wait_ready(k, li);
          if (small) {
            if (k >= 1) wait_ready(k - 1, li);
            else if (li == 0) ft_wait(&B->halo0, 0, "halo0");
            else { ft_wait(&B->ring_full, (uint32_t)(n_ring & 1), "ring_full (halo)"); halo_used = true; }
            a_bytes[1] = (uint32_t)((1 + k) * BM - d) * 128u;
            a_bytes[2] = (uint32_t)((1 + k) * BM - 2 * d) * 128u;
            n = 3;
          } else {
            for (int tap = 1; tap >= 0; --tap) {
              const int rel = k - (2 - tap) * dt;  // source tile relative to this CTA's first tile
              if (rel >= 0) {
                wait_ready(rel, li);
                a_bytes[n] = (uint32_t)(1 + rel) * TILE_B;
                taps[n] = tap;
                ++n;
              } else if (!tap_skipped(k, tap, dt)) {
                foreign[tap] = true;
              }
            }
          }
          if (dbg) { const long long now = clock64(); w_rdy += now - tw0; tw0 = now; }
          if (dbgt && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[16 + 2 * (j - J0)] = clock64() - tk_start;
          ft_fence_after();
          issue_burst(d1, a_bytes, taps, n, buf);
          if (dbg) t_issue += clock64() - tw0;
          if (dbgt && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[16 + 2 * (j - J0) + 1] = clock64() - tk_start;
          if (halo_used) {
            if (ft_elect()) fp_commit(&B->ring_free);
            __syncwarp();
            ++n_ring;
          }
          for (int tap = 1; tap >= 0; --tap) {
            if (!foreign[tap]) continue;
            tw0 = dbg ? clock64() : 0;
            ft_wait(&B->ring_full, (uint32_t)(n_ring & 1), "ring_full");
            if (dbg) w_ring += clock64() - tw0;
            ft_fence_after();
            issue_tap(d1, 0u, buf, tap);
            if (ft_elect()) fp_commit(&B->ring_free);
            __syncwarp();
            ++n_ring;
          }
          if (ft_elect()) {
            fp_commit(&B->d1_full[b]);
            if (kk == K - 1) {
              fp_commit(&B->wd_free[buf]);
              // every foreign tile of this layer has landed in BOTH CTAs' shared memory
              asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.cons + blockIdx.x), "r"((unsigned int)(li + 1)) : "memory");
              asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.cons + blockIdx.x + 1), "r"((unsigned int)(li + 1))
                           : "memory");
            }
          }
          __syncwarp();
        }
      }
      if (dbg && lane == 0) {
        p.dbg[0] = w_wd; p.dbg[1] = w_d1; p.dbg[2] = w_rdy; p.dbg[3] = w_ring; p.dbg[4] = t_issue;
        p.dbg[5] = clock64() - t_begin; p.dbg[6] = K; p.dbg[7] = nl;
      }
      {
        // head pseudo-layer: D1 = cond_out1 (preloaded) + relu(l) . W1, one "tap" on the tile itself
        const int buf = nl & 1;
        ft_wait(&B->wd_full[buf], (uint32_t)((nl >> 1) & 1), "wd_full (head)");
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int k = K - 1 - kk;
          const int b = j & 1;
          ft_wait(&B->d1_empty[b], (uint32_t)((j >> 1) & 1), "d1_empty (head)");
          wait_ready(k, nl);
          ft_fence_after();
          issue_tap(tmem_u + b * 64, (uint32_t)(1 + k) * TILE_B, buf, 2);
          if (ft_elect()) fp_commit(&B->d1_full[b]);
          __syncwarp();
        }
      }
    }
  } else if (warp == 19) {
    // =================================== MMA2 issuer (leader) ===================================
    if (leader) {
      const uint32_t idesc = fp_idesc();
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
      const uint32_t sbase_u = __reduce_or_sync(0xffffffffu, sbase);
      int j = 0;
      ft_wait(&B->l_init, 0, "l_init");
      for (int li = 0; li < nl; ++li) {
        ft_wait(&B->wr_full[li & 1], (uint32_t)((li >> 1) & 1), "wr_full");
        const uint64_t wrh = ft_desc_sw64(sbase_u + OFF_WR + (2 * (li & 1)) * WR_TILE);
        const uint64_t wrl = ft_desc_sw64(sbase_u + OFF_WR + (2 * (li & 1) + 1) * WR_TILE);
        for (int kk = 0; kk < K; ++kk, ++j) {
          const int b = j & 1;
          const uint32_t u = (uint32_t)(j >> 1);
          ft_wait(&B->g_full[b], u & 1, "g_full");
          if (dbgt && lane == 0 && j >= J0 && j < J0 + 8) {
            p.dbg[32 + 2 * (j - J0)] = clock64() - tk_start;
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            p.dbg[192 + (j - J0)] = (long long)t;
          }
          ft_fence_after();
          const uint32_t d2 = tmem_u + TM_L + (K - 1 - kk) * 64;
          const uint32_t g_hi = tmem_u + TM_G + b * 32, g_lo = g_hi + 16;
          if (ft_elect()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              fp_mma_ts(d2, g_lo + 8 * k, wrh + 2 * k, idesc, 1);
              fp_mma_ts(d2, g_hi + 8 * k, wrl + 2 * k, idesc, 1);
              fp_mma_ts(d2, g_hi + 8 * k, wrh + 2 * k, idesc, 1);
            }
            fp_commit(&B->d2_full[K - 1 - kk]);
            fp_commit(&B->g_free[b]);
            if (kk == K - 1) fp_commit(&B->wr_free[li & 1]);
          }
          __syncwarp();
          if (dbgt && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[32 + 2 * (j - J0) + 1] = clock64() - tk_start;
        }
      }
    }
  } else if (warp < 10) {
    // =================================== E1: gate (8 warps per CTA) ===================================
    const int half = (warp - 2) >> 2;
    const int qd = warp & 3;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    auto cond_ptr = [&](int j) -> const float4* {
      const int li = j / K, k = K - 1 - (j - li * K);
      return reinterpret_cast<const float4*>(p.cond + (size_t)li * p.cond_plane +
                                             ((size_t)(gt0 + k) * 8 + qd * 2 + half) * 1024) + lane;
    };
    auto preload = [&](int j, const float4 (&cn)[8]) {  // cond(j) -> D1[j & 1], then release the buffer to the leader
      const int b = j & 1;
      ft_tmem_st32(tmem + lane_sel + b * 64 + half * 32, reinterpret_cast<const uint32_t*>(cn));
      ft_tmem_st_wait();
      ft_fence_before();
      __syncwarp();
      if (lane == 0) fp_arrive_leader(&B->d1_empty[b], leader);
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
      if (dbgt && warp == 2 && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[48 + 3 * (j - J0)] = clock64() - tk_start;
      if (p.dbg != nullptr && (int)(blockIdx.x >> 1) == p.dbg_pair && warp == 2 && lane == 0 && j >= J0 && j < J0 + 8) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.dbg[128 + crank * 32 + 2 * (j - J0)] = (long long)t;
      }
      ft_fence_after();
      uint32_t d[32];
      ft_tmem_ld32(tmem + lane_sel + b * 64 + half * 32, d);
      ft_tmem_ld_wait();
      if (j + 2 < total1) preload(j + 2, cn);
      if (dbgt && warp == 2 && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[48 + 3 * (j - J0) + 1] = clock64() - tk_start;
      if (j >= total) {
        // ---------------- head epilogue (pseudo-layer task) ----------------
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
        // exchange area: the tap-0 slot of the weight buffer the head does not use (idle from here on)
        const uint32_t exch = sbase + wdh_off((nl & 1) ^ 1, 0) + (uint32_t)(j & 1) * 1024u + (uint32_t)row * 8u;
        if (half == 1) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(exch), "f"(pm), "f"(ps) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        if (half == 0) {
          float qm, qs;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(qm), "=f"(qs) : "r"(exch) : "memory");
          const size_t grow = (size_t)gclip * p.T + (size_t)(tk0 + k) * BM + row;
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
            xo = fmaf(p.z[grow], st, mt);
            if (p.quantize) xo = clip_quant_scale_dev(xo, p.quant_chann, p.use_mu_law);
          }
          p.mean_tot[grow] = mt;
          p.scale_tot[grow] = st;
          p.log_scale_tot[grow] = lt;
          p.x_out[grow] = xo;
        }
        continue;
      }
      if (j >= 2) ft_wait(&B->g_free[b], (uint32_t)((((j - 2) >> 1)) & 1), "g_free");
      uint32_t ghi[8], glo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
      if (lane == 0) fp_arrive_leader(&B->g_full[b], leader);
      if (dbgt && warp == 2 && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[48 + 3 * (j - J0) + 2] = clock64() - tk_start;
      if (p.dbg != nullptr && (int)(blockIdx.x >> 1) == p.dbg_pair && warp == 2 && lane == 0 && j >= J0 && j < J0 + 8) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.dbg[128 + crank * 32 + 2 * (j - J0) + 1] = (long long)t;
      }
    }
  } else if (warp < 18) {
    // =================================== E2: residual (8 warps per CTA) ===================================
    const int half = (warp - 10) >> 2;
    const int qd = warp & 3;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    const int row = qd * 32 + lane;
    uint32_t rmx = 0;
    float rmxf = 0.f;
    {
      // start conv of the own tiles (-> TMEM master + smem planes) and of the halo (smem only)
      const float4* w4 = reinterpret_cast<const float4*>(p.start_w) + half * 8;
      const float4* b4 = reinterpret_cast<const float4*>(p.start_b) + half * 8;
      for (int k = K - 1; k >= -1; --k) {
        const int t = (tk0 + k) * BM + row;
        const float* xr = p.x_in + (size_t)gclip * p.T + t;
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
        if (lane == 0) fp_arrive_leader(k >= 0 ? &B->own_loaded[k] : &B->halo0, leader);
      }
    }
    ft_tmem_st_wait();
    ft_fence_before();
    __syncwarp();
    if (lane == 0) fp_arrive_leader(&B->l_init, leader);
    uint32_t npub = 0;
    int li = 0, kk = 0;
    for (int j = 0; j < total; ++j) {
      const int k = K - 1 - kk;
      const float4* bptr = reinterpret_cast<const float4*>(p.br + (size_t)li * C) + half * 8;
      float4 bb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bb[i] = __ldg(bptr + i);
      ft_wait(&B->d2_full[k], (uint32_t)(li & 1), "d2_full");
      if (dbgt && warp == 10 && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[80 + 2 * (j - J0)] = clock64() - tk_start;
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
      const uint32_t np = (npub >> (8 * k)) & 0xffu;
      if (np > 0) ft_wait(&B->pub_done[k], (np - 1) & 1, "pub_done");
      const uint32_t cur_lo = sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      const uint32_t cur_hi = sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B + (uint32_t)row * 128u;
      const bool relu_out = li == nl - 1;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t coff = (uint32_t)(((4 * half + jj) ^ (row & 7)) * 16);
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float o0 = __uint_as_float(d[8 * jj + 2 * e]), o1 = __uint_as_float(d[8 * jj + 2 * e + 1]);
          range_track_fast(rmxf, o0);
          range_track_fast(rmxf, o1);
          if (relu_out) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
          const float a0 = __half2float(__float2half_rn(o0));
          const float a1 = __half2float(__float2half_rn(o1));
          hw[e] = ft_pack_f16(a0, a1);
          lw[e] = ft_pack_f16(o0 - a0, o1 - a1);
        }
        ft_sts128(cur_hi + coff, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        ft_sts128(cur_lo + coff, make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
      ft_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        ft_arrive(&B->tile_ready[k]);                      // local publisher
        fp_arrive_leader(&B->tile_ready_pair[k], leader);  // the MMA1 issuer
      }
      if (published(k, li)) npub += 1u << (8 * k);
      if (dbgt && warp == 10 && lane == 0 && j >= J0 && j < J0 + 8) p.dbg[80 + 2 * (j - J0) + 1] = clock64() - tk_start;
      if (++kk == K) { kk = 0; ++li; }
    }
    range_commit(rmx);
    range_commit_fast(rmxf);
  } else if (warp == 18) {
    // =================================== publisher (each CTA its own tiles) ===================================
    if (lane == 0) {
      const int last_tk = tk0 + K - 1;
      const int far_tk = min(last_tk + p.reach_tiles, p.tiles_per_clip - 1);
      const int far_idx = fp_owner(p.tiles_per_clip, p.pairs_per_clip, far_tk);
      const int my_idx = 2 * R.idx + crank;
      const int cta0 = (int)blockIdx.x - my_idx;  // first CTA of this clip
      for (int li = 0; li < nl; ++li) {
        bool checked = false;
        const int wb = (p.buf0 + li + 1) & 1;
        const CUtensorMap* mh = wb ? &map_h1 : &map_h0;
        const CUtensorMap* ml = wb ? &map_l1 : &map_l0;
        for (int k = K - 1; k >= 0; --k) {
          if (!published(k, li)) continue;
          ft_wait(&B->tile_ready[k], (uint32_t)(li & 1), "tile_ready (publisher)");
          if (!checked && li >= 1) {
            for (int i = my_idx + 1; i <= far_idx; ++i) ft_poll_ge(p.cons + cta0 + i, (unsigned int)li, "cons poll");
            checked = true;
          }
          ft_tma_store_3d(mh, sbase + OFF_HI + (uint32_t)(1 + k) * TILE_B, 0, (tk0 + k) * BM, gclip);
          ft_tma_store_3d(ml, sbase + OFF_LO + (uint32_t)(1 + k) * TILE_B, 0, (tk0 + k) * BM, gclip);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          ft_arrive(&B->pub_done[k]);
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          ft_fence_async();
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + gt0 + k), "r"((unsigned int)(li + 1)) : "memory");
        }
      }
    }
  }

  ft_fence_before();
  __syncthreads();
  // do not exit (or free TMEM) while the peer's MMAs / TMA / remote arrives may still touch this CTA
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    ft_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace
}  // namespace nsw
NSW_RANGE_GUARD_TU(flow_pair)
namespace nsw {

// pairs per clip for a launch of `nclips` clips of T samples, or 0 if the pair kernel does not cover the shape
int flow_pair_pairs_per_clip(int T, int nclips, int max_pairs) {
  if (T <= 0 || T % (2 * BM) != 0 || nclips < 1) return 0;
  const int half = T / BM / 2;
  const int P = std::min(max_pairs / nclips, half);
  if (P < 1 || (half + P - 1) / P > KMAX) return 0;
  return P;
}

int flow_pair_launch(const void* const map_act[2][2], const void* map_wdh32, const void* map_wdl32, const void* map_wrh32,
                     const void* map_wrl32, const float* cond_tiled, size_t cond_plane, const float* br, int T, int nclips,
                     int buf0, int nl, int num_stages, unsigned int* sync_words, int num_sms, const FlowHead* head,
                     const FlowStart* start, cudaStream_t stream) {
  NSW_CHECK(head && start && nl >= 1 && nclips >= 1, NSW_EINVAL, "flow_pair: needs the fused start conv and head");
  static std::atomic<uint64_t> attr_done{0};
  NSW_TRY(ensure_dynamic_smem((const void*)iaf_flow_pair_kernel, (int)FP_SMEM_BYTES, attr_done));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(FT_THREADS);
  cfg.dynamicSmemBytes = FP_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // The per-tile flags need every CTA co-resident.  The grid is at most one CTA per SM and the stream has the GPU to
  // itself when the launch is reached, so a plain cluster launch is co-resident in practice; the cooperative attribute
  // (NSW_FLOW_PAIR_COOP=1) makes the driver check it, but a cooperative CLUSTER launch fails under ncu (LaunchFailed),
  // so it is not the default.  A CTA that waits more than ~2 s traps (ft_wait watchdog) instead of hanging.
  static const bool no_coop = getenv("NSW_FLOW_PAIR_COOP") == nullptr;
  const int n_attr = no_coop ? 1 : 2;
  cfg.numAttrs = n_attr;
  int dev = 0;
  NSW_CUDA(cudaGetDevice(&dev));
  static int max_pairs[64] = {0};
  if (max_pairs[dev & 63] == 0) {
    cfg.gridDim = dim3(num_sms & ~1);
    cfg.numAttrs = 1;  // (the occupancy query takes the cluster shape only)
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, iaf_flow_pair_kernel, &cfg) != cudaSuccess || n < 1) n = num_sms / 2;
    max_pairs[dev & 63] = std::min(n, num_sms / 2);
    cfg.numAttrs = n_attr;
  }
  const int P = flow_pair_pairs_per_clip(T, nclips, max_pairs[dev & 63]);
  NSW_CHECK(P > 0, NSW_EINVAL, "flow_pair: %d clips of %d samples do not fit %d CTA pairs", nclips, T, max_pairs[dev & 63]);
  const int tiles = T / BM;
  FlowPairParams p;
  p.cond = cond_tiled;
  p.cond_plane = cond_plane;
  p.br = br;
  p.buf0 = buf0;
  p.nl = nl;
  p.num_stages = num_stages;
  p.tiles_per_clip = tiles;
  p.clip0 = 0;
  p.nclips = nclips;
  p.pairs_per_clip = P;
  p.reach_tiles = std::max(1, 2 * (1 << (num_stages - 1)) / BM);
  const int grid = 2 * P * nclips;
  const size_t n_flags = (size_t)nclips * tiles;
  p.flags = sync_words;
  p.cons = sync_words + n_flags;
  NSW_CUDA(cudaMemsetAsync(sync_words, 0, (n_flags + grid) * sizeof(unsigned int), stream));
  p.head_w_tile = head->w_tile;
  p.wm = head->wm; p.ws = head->ws; p.bm = head->bm; p.bs = head->bs;
  p.x_in = head->x_in; p.z = head->z; p.x_out = head->x_out;
  p.mean_tot = head->mean_tot; p.scale_tot = head->scale_tot; p.log_scale_tot = head->log_scale_tot;
  p.first = head->first; p.last = head->last; p.quantize = head->quantize; p.use_mu_law = head->use_mu_law;
  p.quant_chann = head->quant_chann;
  p.T = T;
  p.start_w = start->w;
  p.start_b = start->b;
  NSW_CHECK(start->x == head->x_in, NSW_EINVAL, "flow_pair: the start conv reads the flow's input");
  p.dbg = nullptr;
  p.dbg_pair = getenv("NSW_FLOW_PAIR_DEBUG_PAIR") ? atoi(getenv("NSW_FLOW_PAIR_DEBUG_PAIR")) : 0;
  static long long* dbg_buf = nullptr;  // NSW_FLOW_PAIR_DEBUG only (single device)
  const bool want_dbg = getenv("NSW_FLOW_PAIR_DEBUG") != nullptr;
  if (want_dbg) {
    if (!dbg_buf) NSW_CUDA(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
    NSW_CUDA(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), stream));
    p.dbg = dbg_buf;
  }
  cfg.gridDim = dim3(grid);
  static bool said = false;
  if (!said && getenv("NSW_FLOW_PAIR_VERBOSE")) {
    said = true;
    fprintf(stderr, "[flow_pair] grid %d (%d pairs per clip, %d clips, max pairs %d), %d layers, smem %zu B, coop %d\n", grid, P,
            nclips, max_pairs[dev & 63], nl, FP_SMEM_BYTES, no_coop ? 0 : 1);
  }
  NSW_CUDA(cudaLaunchKernelEx(&cfg, iaf_flow_pair_kernel, *reinterpret_cast<const CUtensorMap*>(map_act[0][0]),
                              *reinterpret_cast<const CUtensorMap*>(map_act[0][1]),
                              *reinterpret_cast<const CUtensorMap*>(map_act[1][0]),
                              *reinterpret_cast<const CUtensorMap*>(map_act[1][1]),
                              *reinterpret_cast<const CUtensorMap*>(map_wdh32),
                              *reinterpret_cast<const CUtensorMap*>(map_wdl32),
                              *reinterpret_cast<const CUtensorMap*>(map_wrh32),
                              *reinterpret_cast<const CUtensorMap*>(map_wrl32), p));
  count_launch();
  NSW_CUDA(cudaGetLastError());
  if (want_dbg) {
    long long hb[256];
    NSW_CUDA(cudaStreamSynchronize(stream));
    NSW_CUDA(cudaMemcpy(hb, dbg_buf, sizeof(hb), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[flow_pair dbg pair %d: K=%lld, %lld layers] issuer: %lld cycles; waited: weights %lld, D1 buffer %lld, "
                    "operand tiles/halo %lld, foreign ring %lld; issuing %lld\n", p.dbg_pair, hb[6], hb[7], hb[5], hb[0], hb[1],
            hb[2], hb[3], hb[4]);
    fprintf(stderr, "  tasks 8..15  MMA1 (operands ready, issued):");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld-%lld", hb[16 + 2 * i], hb[16 + 2 * i + 1]);
    fprintf(stderr, "\n               MMA2 (g ready, issued)        :");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld-%lld", hb[32 + 2 * i], hb[32 + 2 * i + 1]);
    fprintf(stderr, "\n               E1 (D1 ready, D1 released, g stored):");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld/%lld/%lld", hb[48 + 3 * i], hb[48 + 3 * i + 1], hb[48 + 3 * i + 2]);
    fprintf(stderr, "\n               E2 (D2 ready, tile stored)    :");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld-%lld", hb[80 + 2 * i], hb[80 + 2 * i + 1]);
    fprintf(stderr, "\n               ns since leader's D1-ready of task 8: per task {leader D1 ready, g stored | peer D1 ready, g stored | MMA2 sees g}:");
    const long long t0 = hb[128];
    for (int i = 0; i < 8; ++i)
      fprintf(stderr, " {%lld,%lld|%lld,%lld|%lld}", hb[128 + 2 * i] - t0, hb[128 + 2 * i + 1] - t0, hb[160 + 2 * i] - t0,
              hb[160 + 2 * i + 1] - t0, hb[192 + i] - t0);
    fprintf(stderr, "\n");
  }
  return NSW_OK;
}

}  // namespace nsw

// TEST HOOK (host only, no CUDA): work split and publish / read plan of one launch of the CTA-pair flow kernel, from the
// same integer functions the kernel uses.  Per CTA (cluster rank fastest): {clip, first tile, tiles, CTAs of the clip,
// index within the clip of the last CTA whose "consumed" counter its publisher polls} then nl*tiles records
// {published, source of tap t-2d, source of tap t-d} as in nsw_flow_plan_host.  Returns the grid size or a negative code.
extern "C" int nsw_flow_pair_plan_host(int32_t T, int32_t nclips, int32_t max_pairs, int32_t nl, int32_t num_stages,
                                       int32_t* out, int64_t cap, int64_t* n_out) {
  using namespace nsw;
  NSW_CHECK(nl >= 1 && num_stages >= 1 && n_out, NSW_EINVAL, "nsw_flow_pair_plan_host: bad argument");
  const int P = flow_pair_pairs_per_clip(T, nclips, max_pairs);
  NSW_CHECK(P > 0, NSW_EINVAL, "nsw_flow_pair_plan_host: the pair kernel does not cover %d clips of %d samples on %d pairs",
            nclips, T, max_pairs);
  const int tiles = T / BM, reach = std::max(1, 2 * (1 << (num_stages - 1)) / BM);
  int64_t n = 0;
  auto put = [&](int v) { if (out && n < cap) out[n] = v; ++n; };
  for (int b = 0; b < 2 * P * nclips; ++b) {
    const PairRange r = fp_range_of(tiles, P, b >> 1);
    const int crank = b & 1, K = r.K, tk0 = 2 * r.h0 + crank * K;
    const int far_tk = std::min(tk0 + K - 1 + reach, tiles - 1);
    put(r.clip); put(tk0); put(K); put(2 * P); put(fp_owner(tiles, P, far_tk));
    for (int li = 0; li < nl; ++li) {
      const int d = 1 << (li % num_stages), dt = d / BM;
      for (int k = 0; k < K; ++k) {
        put(fp_published(k, K, li, nl, num_stages) ? 1 : 0);
        for (int tap = 0; tap < 2; ++tap) {
          int src;
          if (2 * d <= BM) src = k == 0 ? (tk0 >= 1 ? tk0 - 1 : -2) : -1;
          else {
            const int t = tk0 + k - (2 - tap) * dt;
            src = t < 0 ? -2 : (t >= tk0 ? -1 : t);
          }
          put(src);
        }
      }
    }
  }
  *n_out = n;
  return 2 * P * nclips;
}
