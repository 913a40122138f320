// Batched autoregressive teacher WaveNet ("gn" engine of fastgen): BT utterances advance together through ONE
// persistent cooperative kernel, so the per-step weight stream (the cost of an autoregressive step)
// is paid once for all of them.
//
// Replaces, like nsw_fastgen.cu, the per-sample Session.run loop of fastgen.synthesis (wavenet/fastgen.py:147-168)
// around Fastgen.sample (wavenet/wavenet.py:379-514) with masked.causal_linear / masked.linear
// (wavenet/masked.py:328-405), and additionally covers what the latency engine does not: double_gate_width
// (gate 1024, the default when the hparam is absent: wavenet.py:106), the 'ce' head with loss_func.ce_sample
// (loss_func.py:140-151) and batch > 1 (fastgen.py:141 batch_size = mel_encoding.shape[0]).
//
// Work split: 128 CTAs x 512 threads; CTA c owns PPC gate pairs (2 * PPC rows of every dilated conv), 4 residual
// channels and 2 skip channels of every layer.  A step has L + 3 phases, each ending in one grid barrier (an atomic
// counter; the CTAs are co-resident by cooperative launch):
//   ph = 1..L : g_ph = gate(W0 l[t-2d] + W1 l[t-d] + [W2 | M_ph] [l_{ph-2} | g_{ph-1}] + cond)   (same algebra as the
//               latency engine: M_ph = W2_ph Wr_{ph-1} folds the residual update into the next layer's contraction),
//               l_{ph-1} = l_{ph-2} + Wr g_{ph-1} + br  -> history ring of layer ph,   s += Ws g_{ph-1} + bs
//   ph = L+1  : last skip accumulation, relu          ph = L+2 : h = relu(out1 s + cond_out1)
//   ph = L+3  : out = out2 h + b (rows spread over the CTAs), then every CTA samples redundantly.
// Everything that does not depend on the barrier in flight (history taps x [W0 | W1], conditioning, the register
// preload of the next rows' weights) is done before waiting for it.  The tf.FIFOQueue pairs of causal_linear are
// rings [2d+1][BT][512] in HBM/L2; the hoisted conditioning GEMM (Fastgen.cond_vars, wavenet.py:353-377) is
// computed in time chunks so that its buffer stays bounded for long utterances.
#include "nsw_fastgen.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace nsw {

bool gn_supported(const nsw_wavenet_config& c, const char** why) {
  static const char* msg = "";
  auto no = [&](const char* m) { msg = m; if (why) *why = msg; return false; };
  if (c.width != GN_W || c.skip_width != GN_S || c.deconv_width != GN_D)
    return no("fastgen engines need width=512, skip_width=256, deconv_width=256");
  if (c.gate_width != GN_W && c.gate_width != 2 * GN_W) return no("gate_width must be width or 2*width (wavenet.py:204)");
  if (c.filter_length != 3) return no("filter_length must be 3 (masked.py:349)");
  if (c.num_layers < 2 || c.num_layers > GN_MAX_L) return no("num_layers out of range");
  if (c.num_stages < 1 || c.num_stages > 16) return no("bad num_stages");
  if (c.loss_type != NSW_LOSS_MOL && c.loss_type != NSW_LOSS_GAUSS && c.loss_type != NSW_LOSS_CE)
    return no("loss_type must be mol, gauss or ce");
  if (c.out_width < 2 || c.out_width > GN_MAX_O)
    return no("out_width > 256: a 'ce' head without mu-law would be a 65536-way softmax, which is not built");
  if (c.loss_type == NSW_LOSS_MOL && (c.out_width % 3 != 0 || c.out_width / 3 > 31)) return no("bad mol out_width");
  if (c.loss_type == NSW_LOSS_GAUSS && c.out_width != 2) return no("gauss head needs out_width 2");
  return true;
}

// ----------------------------------------------------------------------------------------------------------------
// host-side repacking of the TF-named tensors (checkpoint contract: wavenet.py:227-287) into per-(phase, CTA) blocks
// ----------------------------------------------------------------------------------------------------------------
int gn_pack(const nsw_wavenet_config& cfg, const TensorMap& tm, GnPacked& pk) {
  const char* why = nullptr;
  NSW_CHECK(gn_supported(cfg, &why), NSW_EINVAL, "fastgen: %s", why);
  const int L = cfg.num_layers, O = cfg.out_width, G = cfg.gate_width, MH = G / 2;
  const int W = GN_W, S = GN_S, D = GN_D;
  const GnLayout lay = GnLayout::make(MH);
  pk.lay = lay;
  pk.L = L;
  pk.O = O;
  pk.N = L * G + S;
  std::vector<const float*> Wd(L + 1), bd(L + 1), Wc(L + 1), bc(L + 1), Wr(L + 1), br(L + 1), Ws(L + 1), bs(L + 1);
  for (int i = 1; i <= L; ++i) {
    const std::string li = std::to_string(i);
    Wd[i] = tm.get("dilated_conv_" + li + "/W", (int64_t)3 * W * G);  // [1,3,W,G] (tap, cin, cout)
    bd[i] = tm.get("dilated_conv_" + li + "/biases", G);
    Wc[i] = tm.get("mel_cond_" + li + "/W", (int64_t)D * G);
    bc[i] = tm.get("mel_cond_" + li + "/biases", G);
    Wr[i] = tm.get("res_" + li + "/W", (int64_t)MH * W);  // [1,1,MH,W] (k, cout)
    br[i] = tm.get("res_" + li + "/biases", W);
    Ws[i] = tm.get("skip_" + li + "/W", (int64_t)MH * S);
    bs[i] = tm.get("skip_" + li + "/biases", S);
    if (!Wd[i] || !bd[i] || !Wc[i] || !bc[i] || !Wr[i] || !br[i] || !Ws[i] || !bs[i]) return NSW_EMISSING;
  }
  const float* wcs = tm.get("conv_start/W", 3 * W);
  const float* bcs = tm.get("conv_start/biases", W);
  const float* wss = tm.get("skip_start/W", (int64_t)W * S);
  const float* bss = tm.get("skip_start/biases", S);
  const float* wo1 = tm.get("out1/W", (int64_t)S * S);
  const float* bo1 = tm.get("out1/biases", S);
  const float* wco = tm.get("mel_cond_out1/W", (int64_t)D * S);
  const float* bco = tm.get("mel_cond_out1/biases", S);
  const float* wo2 = tm.get("out2/W", (int64_t)S * O);
  const float* bo2 = tm.get("out2/biases", O);
  if (!wcs || !bcs || !wss || !bss || !wo1 || !bo1 || !wco || !bco || !wo2 || !bo2) return NSW_EMISSING;
  pk.wcs.assign(wcs, wcs + 3 * W);
  pk.bcs.assign(bcs, bcs + W);
  pk.dil.assign(L + 1, 1);
  pk.hist_off.assign(L + 1, 0);
  size_t off = 0;
  for (int i = 1; i <= L; ++i) {
    pk.dil[i] = 1 << ((i - 1) % cfg.num_stages);
    pk.hist_off[i] = (int)off;
    off += (size_t)(2 * pk.dil[i] + 1);
  }
  pk.hist_entries = off;

  const int NPH = L + 3, BF = lay.block_floats, PPC = lay.PPC, nD = lay.nD, K1 = lay.K1;
  pk.blocks.assign((size_t)NPH * GN_NC * BF, 0.f);
  pk.cond_w.assign((size_t)D * pk.N, 0.f);
  pk.cond_b.assign(pk.N, 0.f);
  // dilated-conv output channel of D row j of CTA c: rows [0, PPC) are the sigmoid halves of pairs c*PPC + j,
  // rows [PPC, 2 PPC) the tanh halves (wavenet.py:264-266: d[:, :, :m] sigmoid, d[:, :, m:] tanh)
  auto drow = [&](int c, int j) { return j < PPC ? c * PPC + j : MH + c * PPC + (j - PPC); };
  std::vector<double> Mi((size_t)G * MH), fb(G), W2T((size_t)G * W);
  for (int i = 1; i <= L; ++i) {
    const float* W0 = Wd[i];
    const float* W1 = Wd[i] + (size_t)W * G;
    const float* W2 = Wd[i] + (size_t)2 * W * G;
    if (i >= 2) {
      for (int cin = 0; cin < W; ++cin)
        for (int co = 0; co < G; ++co) W2T[(size_t)co * W + cin] = W2[(size_t)cin * G + co];
      for (int co = 0; co < G; ++co) {  // M_i[co][k] = sum_cin W2[cin][co] Wr_{i-1}[k][cin], fb = W2 br_{i-1}
        const double* w2r = &W2T[(size_t)co * W];
        for (int k = 0; k < MH; ++k) {
          const float* wr = Wr[i - 1] + (size_t)k * W;
          double acc = 0.0;
          for (int cin = 0; cin < W; ++cin) acc += w2r[cin] * (double)wr[cin];
          Mi[(size_t)co * MH + k] = acc;
        }
        double accb = 0.0;
        for (int cin = 0; cin < W; ++cin) accb += w2r[cin] * (double)br[i - 1][cin];
        fb[co] = accb;
      }
    }
    for (int c = 0; c < GN_NC; ++c) {
      float* blk = &pk.blocks[((size_t)(i - 1) * GN_NC + c) * BF];
      for (int j = 0; j < nD; ++j) {
        const int co = drow(c, j);
        float* d = blk + lay.off_d + j * K1;
        for (int k = 0; k < W; ++k) d[k] = W2[(size_t)k * G + co];
        if (i >= 2)
          for (int k = 0; k < MH; ++k) d[W + k] = (float)Mi[(size_t)co * MH + k];
        float* p = blk + lay.off_p + j * 2 * W;
        for (int k = 0; k < W; ++k) {
          p[k] = W0[(size_t)k * G + co];      // tap t - 2d
          p[W + k] = W1[(size_t)k * G + co];  // tap t - d
        }
        const int n = (i - 1) * G + c * nD + j;  // hoisted conditioning column + every static bias of this row
        for (int k = 0; k < D; ++k) pk.cond_w[(size_t)k * pk.N + n] = Wc[i][(size_t)k * G + co];
        pk.cond_b[n] = (float)((double)bd[i][co] + (double)bc[i][co] + (i >= 2 ? fb[co] : 0.0));
      }
      if (i >= 2) {
        for (int j = 0; j < 4; ++j) {
          const int lc = 4 * c + j;
          float* lrow = blk + lay.off_l + j * MH;
          for (int k = 0; k < MH; ++k) lrow[k] = Wr[i - 1][(size_t)k * W + lc];
          blk[lay.off_c + j] = br[i - 1][lc];
        }
        for (int j = 0; j < 2; ++j) {
          const int sc = 2 * c + j;
          float* srow = blk + lay.off_s + j * W;
          for (int k = 0; k < MH; ++k) srow[k] = Ws[i - 1][(size_t)k * S + sc];
          blk[lay.off_c + 4 + j] = bs[i - 1][sc];
        }
      } else {
        for (int j = 0; j < 2; ++j) {  // skip_start on l_0 (wavenet.py:444)
          const int sc = 2 * c + j;
          float* srow = blk + lay.off_s + j * W;
          for (int k = 0; k < W; ++k) srow[k] = wss[(size_t)k * S + sc];
          blk[lay.off_c + 4 + j] = bss[sc];
        }
      }
    }
  }
  const int RO = (O + GN_NC - 1) / GN_NC;
  for (int c = 0; c < GN_NC; ++c) {
    float* b1 = &pk.blocks[((size_t)L * GN_NC + c) * BF];  // phase L+1: skip_L, then relu
    for (int j = 0; j < 2; ++j) {
      const int sc = 2 * c + j;
      for (int k = 0; k < MH; ++k) b1[lay.off_s + j * W + k] = Ws[L][(size_t)k * S + sc];
      b1[lay.off_c + 4 + j] = bs[L][sc];
    }
    float* b2 = &pk.blocks[((size_t)(L + 1) * GN_NC + c) * BF];  // phase L+2: out1 rows 2c, 2c+1
    for (int j = 0; j < 2; ++j) {
      const int oc = 2 * c + j;
      for (int k = 0; k < S; ++k) b2[lay.off_d + j * K1 + k] = wo1[(size_t)k * S + oc];
      const int n = L * G + oc;
      for (int k = 0; k < D; ++k) pk.cond_w[(size_t)k * pk.N + n] = wco[(size_t)k * S + oc];
      pk.cond_b[n] = bo1[oc] + bco[oc];
    }
    float* b3 = &pk.blocks[((size_t)(L + 2) * GN_NC + c) * BF];  // phase L+3: out2 rows c*RO .. c*RO+RO-1
    for (int j = 0; j < RO; ++j) {
      const int o = c * RO + j;
      if (o >= O) break;
      for (int k = 0; k < S; ++k) b3[lay.off_d + j * K1 + k] = wo2[(size_t)k * O + o];
      b3[lay.off_c + 6 + j] = bo2[o];
    }
  }
  return NSW_OK;
}

namespace {

constexpr long long GN_WATCHDOG = 6000000000ll;  // ~3 s of SM clocks

struct GnParams {
  const float* blocks;
  const float* cond;       // [BT][NPL][Tc][64], chunk-local
  size_t cond_bstride;     // floats between batch rows
  int Tc;                  // rows per plane of this chunk
  float* hist;             // rings, entry = [BT][W]
  const int* hist_off;     // [L+1] in entries
  const int* dil;          // [L+1]
  float* gbuf;             // [2][BT][MH]
  float* sbuf;             // [BT][S]
  float* hbuf;             // [BT][S]
  float* obuf;             // [BT][O]
  float* xstate;           // [BT][4]: fed sample (input-encoded), x[t-1], x[t-2] carried across chunk launches
  unsigned int* bar;       // grid barrier counter, zero at launch
  const float* wcs;        // conv_start W [3][W]
  const float* bcs;        // [W]
  const float* tf;         // teacher forcing, row b at tf + b * T, or NULL
  const float* noise;      // supplied sampler noise, row b at noise + b * T * nu, or NULL
  float* audio;            // row b at audio + b * T, or NULL
  float* out;              // row b at out + b * T * O, or NULL
  int nb;                  // live batch rows (<= BT)
  int T, t0, t1;           // utterance length, chunk [t0, t1)
  int L, O, loss_type, use_mu_law, nu;
  unsigned long long seed;  // row b uses seed + golden * (b0 + b): seed_b0 carries b0
  int b0;
  float quant;
  long long* dbg;          // NSW_FASTGEN_DEBUG: [NC][8] cycle sums of thread 0 (NULL = off)
  int flags;               // experiment switches (NSW_GN_FLAGS): 1 = red.release arrive
  int l2_last;             // weight blocks of phases < l2_last are loaded L2::evict_last, the rest evict_first (0 = no hints)
};

template <int MH, int BT>
struct GnSmem {
  // weight blocks of two consecutive phases (cp.async.bulk, one phase ahead)
  alignas(128) float wring[2][GnLayout::make(MH).block_floats];
  alignas(128) float xl[BT][GN_W];    // fresh l_{ph-2} | head phases: s / h, rows of 256 | sampler: out, rows of O
  alignas(128) float xg[BT][MH];      // fresh g_{ph-1}
  alignas(128) float hv[2][BT][GN_W];  // [0] = l[t-2d], [1] = l[t-d] of the phase's layer
  // K-split partial sums: a warp contracts ONE 128-float slice of the inputs against every row that uses it, so the
  // inputs are read from shared memory once per slice instead of once per row (the dots were shared-memory-bandwidth
  // bound: rows x K x 4 B x BT = 147 - 360 KB per phase at BT = 8)
  float pf[12][64];          // fresh stage: [slice warp][row * BT + b] (+ 4 * BT for the skip rows)
  float pp[8][64];           // past taps:   [tap * 4 + slice][row * BT + b]
  float pvc[64];             // past taps + conditioning (+ every folded bias) per (row, b), summed off the critical path
  float cnd[8][BT];
  float lst[4][BT];          // owned residual channels 4c..4c+3
  float sst[2][BT];          // owned skip channels 2c, 2c+1
  float xn[BT], x1[BT], x2[BT];
  int dil[GN_MAX_L + 4], hoff[GN_MAX_L + 4], pos[GN_MAX_L + 4];
  unsigned long long wbar[2], hbar, fbar;
};

static_assert(sizeof(GnSmem<512, GN_MAX_BT>) <= 227 * 1024, "batched fastgen shared memory");

__device__ __noinline__ void gn_die(const char* what) {
  printf("nsw fastgen(gn): watchdog in %s (block %d thread %d)\n", what, blockIdx.x, threadIdx.x);
  __trap();
}

__device__ __forceinline__ void gn_mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gn_expect(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gn_bulk(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// weight blocks: the per-step stream (156 - 330 MB) is larger than L2 and cyclic, so plain LRU never hits; the blocks of
// the first phases are kept (evict_last), the rest stream through (evict_first) -- as in the latency engine
__device__ __forceinline__ void gn_bulk_hint(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, bool keep) {
  unsigned long long pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void gn_mbar_wait(unsigned long long* bar, uint32_t parity, const char* what) {
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
      else if (clock64() - t0 > GN_WATCHDOG) gn_die(what);
    }
  }
}

// transposing butterfly over V = 2^k per-lane partials (k <= 5 halving steps with xor offsets 16, 8, ...; plain
// xor-adds after that): afterwards a[0] holds the warp total of value number lane >> (5 - k)
template <int N, int O>
__device__ __forceinline__ void red_step(float* a, int lane) {
  if constexpr (O >= 1) {
    if constexpr (N > 1) {
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        const float keep = up ? a[N / 2 + i] : a[i];
        const float send = up ? a[i] : a[N / 2 + i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      red_step<N / 2, O / 2>(a, lane);
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], O);
      red_step<1, O / 2>(a, lane);
    }
  }
}
// One 128-float input slice (x: BT rows, xs floats apart, already at the slice) against NR <= 4 weight rows (w: rows ws
// floats apart, already at the slice): out[r * BT + b] = sum over the slice.  NR * BT is a power of two <= 32.
template <int NR, int BT>
__device__ __forceinline__ void slice_rows(const float* w, int ws, const float4 (&xv)[BT], int lane, float* out) {
  constexpr int V = NR * BT;
  static_assert(V <= 32 && (V & (V - 1)) == 0, "slice_rows: NR * BT must be a power of two <= 32");
  float acc[V];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)r * ws + 4 * lane);
#pragma unroll
    for (int b = 0; b < BT; ++b)
      acc[r * BT + b] = fmaf(wv.x, xv[b].x, fmaf(wv.y, xv[b].y, fmaf(wv.z, xv[b].z, wv.w * xv[b].w)));
  }
  red_step<V, 16>(acc, lane);
  constexpr int STEP = 32 / V;   // lanes per value
  if ((lane & (STEP - 1)) == 0) out[lane / STEP] = acc[0];
}
template <int NR, int BT>
__device__ __forceinline__ void slice_job(const float* w, int ws, const float* x, int xs, int lane, float* out) {
  float4 xv[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) xv[b] = *reinterpret_cast<const float4*>(x + (size_t)b * xs + 4 * lane);
  if constexpr (NR <= 4) {
    slice_rows<NR, BT>(w, ws, xv, lane, out);
  } else {  // 8 rows as two groups of 4: at most 32 accumulators live
    slice_rows<4, BT>(w, ws, xv, lane, out);
    slice_rows<4, BT>(w + (size_t)4 * ws, ws, xv, lane, out + 4 * BT);
  }
}

__device__ __forceinline__ float gn_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gn_tanh(float x) { return 2.0f * gn_sigmoid(2.0f * x) - 1.0f; }

// Warp 15 is the "comm" warp: its lane 0 streams the weight blocks (one phase ahead) and the history taps (as soon as
// the previous past-tap dots are done) into shared memory with cp.async.bulk, polls the grid barrier of the previous
// phase WHILE the other warps do their barrier-independent work, and fetches the fresh inputs the moment it opens.
template <int MH, int BT>
__global__ void __launch_bounds__(GN_NT, 1) fastgen_gn_kernel(GnParams P) {
  constexpr int W = GN_W, S = GN_S, NT = GN_NT, NC = GN_NC;
  constexpr int PPC = MH / NC, nD = 2 * PPC, K1 = W + MH, G = 2 * MH;
  constexpr int OFF_D = 0, OFF_P = OFF_D + nD * K1, OFF_L = OFF_P + nD * 2 * W, OFF_S = OFF_L + 4 * MH,
                OFF_C = OFF_S + 2 * W, BF = OFF_C + 16;
  constexpr int NFM = MH / 128;
  constexpr int COMM = NT / 32 - 1;
  static_assert(4 + 2 * NFM <= COMM - 1, "slice warps, the past-sum warp and the comm warp must not overlap");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Smem = GnSmem<MH, BT>;
  Smem& Sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
  const int L = P.L, O = P.O, T = P.T, NPH = L + 3;
  const int RO = (O + NC - 1) / NC;
  const bool comm = warp == COMM && lane == 0;
  float* outv = &Sm.xl[0][0];  // sampler view of the head outputs, rows of O floats

  if (tid <= L) {
    const int d = tid >= 1 ? P.dil[tid] : 1;
    Sm.dil[tid] = d;
    Sm.hoff[tid] = tid >= 1 ? P.hist_off[tid] : 0;
    Sm.pos[tid] = P.t0 % (2 * d + 1);
  }
  if (tid < BT) {
    Sm.xn[tid] = P.xstate[4 * tid];
    Sm.x1[tid] = P.xstate[4 * tid + 1];
    Sm.x2[tid] = P.xstate[4 * tid + 2];
  }
  if (tid == 0) {
    gn_mbar_init(&Sm.wbar[0], 1);
    gn_mbar_init(&Sm.wbar[1], 1);
    gn_mbar_init(&Sm.hbar, 1);
    gn_mbar_init(&Sm.fbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t ring_entry = (size_t)BT * W;
  constexpr uint32_t ENTRY_B = BT * W * 4, G_B = BT * MH * 4, S_B = BT * S * 4, BLOCK_B = BF * 4;
  // ---- comm-lane helpers ----
  auto issue_weights = [&](int ph0 /* 0-based phase */, int slot) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot was read through the generic proxy
    gn_expect(&Sm.wbar[slot], BLOCK_B);
    if (P.l2_last > 0)
      gn_bulk_hint(Sm.wring[slot], P.blocks + ((size_t)ph0 * NC + c) * BF, BLOCK_B, &Sm.wbar[slot], ph0 < P.l2_last);
    else
      gn_bulk(Sm.wring[slot], P.blocks + ((size_t)ph0 * NC + c) * BF, BLOCK_B, &Sm.wbar[slot]);
  };
  auto issue_hist = [&](int ph, int t, int pos) {  // taps of layer ph for step t; pos = ring slot of step t
    const int d = Sm.dil[ph], R = 2 * d + 1;
    int p1 = pos - d; if (p1 < 0) p1 += R;
    int p2 = p1 - d; if (p2 < 0) p2 += R;
    const float* ring = P.hist + (size_t)Sm.hoff[ph] * ring_entry;
    const bool ok2 = t - 2 * d >= 0, ok1 = t - d >= 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    gn_expect(&Sm.hbar, (ok2 ? ENTRY_B : 0u) + (ok1 ? ENTRY_B : 0u));
    if (ok2) gn_bulk(&Sm.hv[0][0][0], ring + (size_t)p2 * ring_entry, ENTRY_B, &Sm.hbar);
    if (ok1) gn_bulk(&Sm.hv[1][0][0], ring + (size_t)p1 * ring_entry, ENTRY_B, &Sm.hbar);
  };
  unsigned int bar_target = 0;
  auto poll_grid = [&]() {  // comm lane: every CTA has published the previous phase
    long long t0 = 0;
    int spins = 0;
    for (;;) {
      unsigned int seen;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.bar) : "memory");
      if ((int)(seen - bar_target) >= 0) break;
      if (++spins == 1024) {
        spins = 0;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > GN_WATCHDOG) gn_die("grid barrier");
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy publishes -> the bulk copies issued next
  };
  auto arrive = [&]() {
    __syncthreads();
    if (tid == 0) {
      if (P.flags & 1) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.bar) : "memory");
      } else {
        __threadfence();
        atomicAdd(P.bar, 1u);
      }
    }
    bar_target += NC;
  };

  if (comm) {
    asm volatile("fence.proxy.async;" ::: "memory");
    issue_weights(0, 0);
    issue_hist(1, P.t0, Sm.pos[1]);
  }
  uint32_t nh = 0, nfr = 0;  // completed uses of hbar / fbar (parities)
  long long gq = 0;          // phases since launch
  long long dacc[6] = {0, 0, 0, 0, 0, 0};
  const bool dbg = P.dbg != nullptr && tid == 0;

  // sample step ts from obuf (all CTAs, redundantly); runs between the grid barrier of its phase L+3 and phase 1
  auto sample_step = [&](int ts) {
    for (int i = tid; i < BT * O; i += NT) outv[i] = __ldcg(P.obuf + i);
    __syncthreads();
    if (c == 0 && P.out) {
      for (int i = tid; i < P.nb * O; i += NT) {
        const int b = i / O, o = i - b * O;
        P.out[((size_t)b * T + ts) * O + o] = outv[i];
      }
    }
    if (warp < BT) {
      const int b = warp;
      const float* ov = outv + (size_t)b * O;
      const float Q = P.quant;
      const unsigned long long sd = P.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(P.b0 + b);
      const uint2 key = make_uint2((uint32_t)sd, (uint32_t)(sd >> 32));
      const float* nz = (P.noise && b < P.nb) ? P.noise + ((size_t)b * T + ts) * P.nu : nullptr;
      auto bits_of = [&](int j) -> uint32_t {  // same stream as the latency engine: counter (t, j / 4), word j % 4
        const uint4 rr = philox4x32_10(make_uint4((uint32_t)ts, (uint32_t)(j >> 2), 0x66617374u, 0u), key);
        return (j & 3) == 0 ? rr.x : (j & 3) == 1 ? rr.y : (j & 3) == 2 ? rr.z : rr.w;
      };
      float a;  // dequantised sample
      if (P.loss_type == NSW_LOSS_MOL) {
        // loss_func.mol_sample (loss_func.py:154-186)
        const int nr = O / 3;
        float v = -INFINITY;
        if (lane < nr) {
          const float u = nz ? __ldg(nz + lane) : u01_clipped(bits_of(lane));
          v = ov[lane] - logf(-logf(u));
        }
        int sel = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // argmax, lowest index on ties (np.argmax / tf.argmax)
          const float ov2 = __shfl_xor_sync(0xffffffffu, v, o);
          const int os = __shfl_xor_sync(0xffffffffu, sel, o);
          if (ov2 > v || (ov2 == v && os < sel)) { v = ov2; sel = os; }
        }
        const float u2 = nz ? __ldg(nz + nr) : u01_clipped(bits_of(nr));
        const float mu = ov[nr + sel];
        const float lsc = fminf(fmaxf(ov[2 * nr + sel], -7.0f), 7.0f);
        const float x = fmaf(expf(lsc), logf(u2) - logf(1.0f - u2), mu);
        a = clip_quant_scale_dev(x, Q, P.use_mu_law);
      } else if (P.loss_type == NSW_LOSS_GAUSS) {
        // loss_func.gauss_sample (loss_func.py:200-206); Box-Muller radius from an unclipped uniform in (0,1]
        float n;
        if (nz) n = __ldg(nz);
        else {
          const float u0 = ((float)(bits_of(0) >> 8) + 1.0f) * (1.0f / 16777216.0f);
          const float u1 = ((float)(bits_of(1) >> 8) + 0.5f) * (1.0f / 16777216.0f);
          n = sqrtf(-2.0f * logf(u0)) * cosf(6.283185307179586f * u1);
        }
        const float x = fmaf(expf(fmaxf(ov[1], -7.0f)), n, ov[0]);
        a = clip_quant_scale_dev(x, Q, P.use_mu_law);
      } else {
        // loss_func.ce_sample (loss_func.py:140-151): one categorical draw from softmax(out), by inverse CDF on a
        // single uniform: k = min{ k : sum_{i<=k} p_i > u * sum_i p_i };  s = k - Q/2
        const int NE = (O + 31) / 32;
        float m = -INFINITY;
        for (int e = 0; e < NE; ++e) { const int k = lane * NE + e; if (k < O) m = fmaxf(m, ov[k]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float mine = 0.f;
        for (int e = 0; e < NE; ++e) { const int k = lane * NE + e; if (k < O) mine += expf(ov[k] - m); }
        float incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const float total = __shfl_sync(0xffffffffu, incl, 31);
        const float u = nz ? __ldg(nz) : ((float)(bits_of(0) >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float thr = u * total;
        const unsigned hit = __ballot_sync(0xffffffffu, incl > thr);
        int k = O - 1;
        if (hit) {
          const int src = __ffs(hit) - 1;
          int kk = O - 1;
          if (lane == src) {
            float cum = incl - mine;
            for (int e = 0; e < NE; ++e) {
              const int q = lane * NE + e;
              if (q >= O) break;
              cum += expf(ov[q] - m);
              if (cum > thr || e == NE - 1 || q == O - 1) { kk = q; break; }
            }
          }
          k = __shfl_sync(0xffffffffu, kk, src);
        }
        a = inv_quant_dev((float)k - 0.5f * Q, Q, P.use_mu_law);  // fastgen.py:162-166
      }
      if (lane == 0) {
        const float fed = (P.tf && b < P.nb) ? P.tf[(size_t)b * T + ts] : a;
        Sm.xn[b] = P.use_mu_law ? mu_law_scaled_dev(fed, Q) : fed;  // wavenet.py:411-414
        if (c == 0 && P.audio && b < P.nb) P.audio[(size_t)b * T + ts] = fed;
      }
    }
    __syncthreads();
  };

  for (int t = P.t0; t < P.t1; ++t) {
    const int tl = t - P.t0;
    for (int ph = 1; ph <= NPH; ++ph, ++gq) {
      const int slot = (int)(gq & 1);
      const float* blk = Sm.wring[slot];
      long long tq = dbg ? clock64() : 0;
      auto lap = [&](int i) { if (dbg) { const long long now = clock64(); dacc[i] += now - tq; tq = now; } };
      const bool last_phase = (t + 1 == P.t1) && ph == NPH;
      // ---------------- comm lane: next weights, then the previous phase's exchange ----------------
      if (comm) {
        // (the next weight block is requested AFTER the fresh inputs: bulk copies are served in order, and 37-78 KB
        // of weights in front of them would delay the critical path by their transfer time)
        if (ph >= 2 || t > P.t0) poll_grid();
        if (ph >= 2) {
          // fresh inputs of this phase (xl / xg were last read before the previous arrive)
          if (ph <= L) {
            gn_expect(&Sm.fbar, ENTRY_B + G_B);
            gn_bulk(&Sm.xl[0][0], P.hist + ((size_t)Sm.hoff[ph - 1] + Sm.pos[ph - 1]) * ring_entry, ENTRY_B, &Sm.fbar);
            gn_bulk(&Sm.xg[0][0], P.gbuf + (size_t)((ph - 1) & 1) * BT * MH, G_B, &Sm.fbar);
          } else if (ph == L + 1) {
            gn_expect(&Sm.fbar, G_B);
            gn_bulk(&Sm.xg[0][0], P.gbuf + (size_t)(L & 1) * BT * MH, G_B, &Sm.fbar);
          } else {
            gn_expect(&Sm.fbar, S_B);
            gn_bulk(&Sm.xl[0][0], ph == L + 2 ? P.sbuf : P.hbuf, S_B, &Sm.fbar);
          }
        }
        if (!last_phase) issue_weights(ph == NPH ? 0 : ph, slot ^ 1);
      }
      // ---------------- everything that does not depend on the exchange in flight ----------------
      gn_mbar_wait(&Sm.wbar[slot], (uint32_t)((gq >> 1) & 1), "weight block");
      bool ok2 = false, ok1 = false;
      if (ph <= L) {
        const int d = Sm.dil[ph];
        ok2 = t - 2 * d >= 0;
        ok1 = t - d >= 0;
        if (tid >= 256 && tid < 256 + nD * BT) {
          const int j = (tid - 256) / BT, b = (tid - 256) - j * BT;
          const int n = (ph - 1) * G + c * nD + j;
          Sm.cnd[j][b] = __ldg(P.cond + (size_t)b * P.cond_bstride + ((size_t)(n >> 6) * P.Tc + tl) * 64 + (n & 63));
        }
        if (warp < 8) {
          // past taps, K-split: warp = tap * 4 + slice; every dilated-conv row of the CTA against that slice
          const int tap = warp >> 2, sl = warp & 3;
          if (tap == 0 ? ok2 : ok1) {
            gn_mbar_wait(&Sm.hbar, nh & 1u, "history taps");
            slice_job<nD, BT>(blk + OFF_P + tap * W + sl * 128, 2 * W, &Sm.hv[tap][0][sl * 128], W, lane, Sm.pp[warp]);
          }
        }
        ++nh;
      } else if (ph == L + 2) {
        if (tid >= 256 && tid < 256 + 2 * BT) {
          const int j = (tid - 256) / BT, b = (tid - 256) - j * BT;
          const int n = L * G + 2 * c + j;
          Sm.cnd[j][b] = __ldg(P.cond + (size_t)b * P.cond_bstride + ((size_t)(n >> 6) * P.Tc + tl) * 64 + (n & 63));
        }
      }
      __syncthreads();  // S1: past taps done (hv is free), cnd / pv staged; the comm lane has seen the grid barrier
      lap(0);
      if (warp == COMM - 1 && ph <= L) {
        // while the fresh inputs are in flight: past-tap partials of the 8 slice warps + conditioning -> one value per
        // (row, b), so that the epilogue on the critical path adds a single term
        for (int idx = lane; idx < nD * BT; idx += 32) {
          float a = Sm.cnd[idx / BT][idx % BT];
          if (ok2) a += (Sm.pp[0][idx] + Sm.pp[1][idx]) + (Sm.pp[2][idx] + Sm.pp[3][idx]);
          if (ok1) a += (Sm.pp[4][idx] + Sm.pp[5][idx]) + (Sm.pp[6][idx] + Sm.pp[7][idx]);
          Sm.pvc[idx] = a;
        }
      }
      if (comm && !last_phase) {
        // history taps of the next phase with a dilated conv (phase 1 of the next step after the head phases)
        if (ph < L) issue_hist(ph + 1, t, Sm.pos[ph + 1]);
        else if (ph == NPH) {
          const int R = 2 * Sm.dil[1] + 1;
          const int pn = Sm.pos[1] + 1;
          issue_hist(1, t + 1, pn >= R ? 0 : pn);
        }
      }
      // ---------------- fresh inputs ----------------
      if (ph == 1) {
        if (t > P.t0) sample_step(t - 1);
        // conv_start on the fed-back sample, every CTA the full vector (causal_linear rate 1, masked.py:352-376)
        for (int i = tid; i < BT * W; i += NT) {
          const int b = i / W, k = i - b * W;
          Sm.xl[b][k] = fmaf(__ldg(P.wcs + 2 * W + k), Sm.xn[b],
                             fmaf(__ldg(P.wcs + W + k), Sm.x1[b], fmaf(__ldg(P.wcs + k), Sm.x2[b], __ldg(P.bcs + k))));
        }
        for (int i = tid; i < BT * MH; i += NT) (&Sm.xg[0][0])[i] = 0.f;  // no gate output before layer 1
        __syncthreads();
        if (tid < BT) {  // conv_start's two queues (rate 1)
          Sm.x2[tid] = Sm.x1[tid];
          Sm.x1[tid] = Sm.xn[tid];
        }
      } else {
        gn_mbar_wait(&Sm.fbar, nfr & 1u, "fresh inputs");
        ++nfr;
      }
      lap(1);
      // fresh rows, K-split: one warp per 128-float slice of xl (warps 0-3) and of xg (D rows: warps 4..4+NFM-1,
      // residual + skip rows: the next NFM warps)
      if (ph <= L) {
        if (warp < 4) {
          slice_job<nD, BT>(blk + OFF_D + warp * 128, K1, &Sm.xl[0][warp * 128], W, lane, Sm.pf[warp]);
        } else if (ph == 1) {
          if (warp < 8)  // skip_start on l_0 (K = W)
            slice_job<2, BT>(blk + OFF_S + (warp - 4) * 128, W, &Sm.xl[0][(warp - 4) * 128], W, lane, Sm.pf[warp]);
        } else if (warp < 4 + NFM) {
          const int sl = warp - 4;
          slice_job<nD, BT>(blk + OFF_D + W + sl * 128, K1, &Sm.xg[0][sl * 128], MH, lane, Sm.pf[warp]);
        } else if (warp < 4 + 2 * NFM) {
          const int sl = warp - 4 - NFM;
          slice_job<4, BT>(blk + OFF_L + sl * 128, MH, &Sm.xg[0][sl * 128], MH, lane, Sm.pf[warp]);
          slice_job<2, BT>(blk + OFF_S + sl * 128, W, &Sm.xg[0][sl * 128], MH, lane, Sm.pf[warp] + 4 * BT);
        }
      } else if (ph == L + 1) {
        if (warp >= 4 + NFM && warp < 4 + 2 * NFM) {
          const int sl = warp - 4 - NFM;
          slice_job<2, BT>(blk + OFF_S + sl * 128, W, &Sm.xg[0][sl * 128], MH, lane, Sm.pf[warp] + 4 * BT);
        }
      } else {
        if (warp < S / 128)  // out1 (L+2) / out2 (L+3) rows, K = S, inputs in the first S floats of xl with stride S
          slice_job<2, BT>(blk + OFF_D + warp * 128, K1, &Sm.xl[0][0] + warp * 128, S, lane, Sm.pf[warp]);
      }
      __syncthreads();  // S2
      lap(2);
      // ---------------- epilogues: one thread per (row, batch row) ----------------
      auto sum_pf = [&](int w0, int w1, int idx) {  // two chains, fixed trip count: the loads issue together
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < 12; w += 2) {
          if (w >= w0 && w < w1) a0 += Sm.pf[w][idx];
          if (w + 1 >= w0 && w + 1 < w1) a1 += Sm.pf[w + 1][idx];
        }
        return a0 + a1;
      };
      if (ph <= L) {
        const int dw = ph == 1 ? 4 : 4 + NFM;  // slice warps holding partials of the dilated-conv rows
        if (tid < PPC * BT) {
          const int pr = tid / BT, b = tid - pr * BT;
          const float a = sum_pf(0, dw, pr * BT + b) + Sm.pvc[pr * BT + b];
          const float q = sum_pf(0, dw, (PPC + pr) * BT + b) + Sm.pvc[(PPC + pr) * BT + b];
          __stcg(P.gbuf + (size_t)(ph & 1) * BT * MH + (size_t)b * MH + c * PPC + pr, gn_sigmoid(a) * gn_tanh(q));
        } else if (tid >= 64 && tid < 64 + 4 * BT) {
          const int r = (tid - 64) / BT, b = (tid - 64) - r * BT;
          // l_{ph-1} = l_{ph-2} + Wr_{ph-1} g_{ph-1} + br_{ph-1}  (phase 1: l_0 from conv_start)
          const float l = ph == 1 ? Sm.xl[b][4 * c + r]
                                  : Sm.lst[r][b] + sum_pf(4 + NFM, 4 + 2 * NFM, r * BT + b) + blk[OFF_C + r];
          Sm.lst[r][b] = l;
          __stcg(P.hist + ((size_t)Sm.hoff[ph] + Sm.pos[ph]) * ring_entry + (size_t)b * W + 4 * c + r, l);
        } else if (tid >= 128 && tid < 128 + 2 * BT) {
          const int r = (tid - 128) / BT, b = (tid - 128) - r * BT;
          const float a = (ph == 1 ? sum_pf(4, 8, r * BT + b) : sum_pf(4 + NFM, 4 + 2 * NFM, 4 * BT + r * BT + b)) +
                          blk[OFF_C + 4 + r];
          Sm.sst[r][b] = ph == 1 ? a : Sm.sst[r][b] + a;  // skip_start, then skip_{ph-1}
        }
      } else if (ph == L + 1) {
        if (tid >= 128 && tid < 128 + 2 * BT) {
          const int r = (tid - 128) / BT, b = (tid - 128) - r * BT;
          const float s = Sm.sst[r][b] + sum_pf(4 + NFM, 4 + 2 * NFM, 4 * BT + r * BT + b) + blk[OFF_C + 4 + r];
          __stcg(P.sbuf + (size_t)b * S + 2 * c + r, fmaxf(s, 0.f));  // relu(s) (wavenet.py:494)
        }
      } else if (ph == L + 2) {
        if (tid < 2 * BT) {
          const int r = tid / BT, b = tid - r * BT;
          __stcg(P.hbuf + (size_t)b * S + 2 * c + r, fmaxf(sum_pf(0, S / 128, r * BT + b) + Sm.cnd[r][b], 0.f));
        }
      } else {
        if (tid < RO * BT) {
          const int r = tid / BT, b = tid - r * BT;
          const int o = c * RO + r;
          if (o < O) __stcg(P.obuf + (size_t)b * O + o, sum_pf(0, S / 128, r * BT + b) + blk[OFF_C + 6 + r]);
        }
      }
      lap(3);
      arrive();
      lap(4);
    }
    if (tid >= 1 && tid <= L) {  // ring slot of step t+1 (the comm lane already used pos[1] + 1 for its prefetch)
      const int R = 2 * Sm.dil[tid] + 1;
      const int pn = Sm.pos[tid] + 1;
      Sm.pos[tid] = pn >= R ? 0 : pn;
    }
    __syncthreads();
  }
  // the last step of this launch: its head outputs are behind the last grid barrier
  if (comm) poll_grid();
  __syncthreads();
  sample_step(P.t1 - 1);
  if (dbg)
    for (int i = 0; i < 5; ++i) P.dbg[8 * c + i] = dacc[i];
  if (c == 0 && tid < BT) {
    P.xstate[4 * tid] = Sm.xn[tid];
    P.xstate[4 * tid + 1] = Sm.x1[tid];
    P.xstate[4 * tid + 2] = Sm.x2[tid];
  }
}

__global__ void gn_split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                    size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  uint32_t rmx = 0;
  range_track(rmx, v);
  range_commit(rmx);
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

template <int MH, int BT>
int gn_launch(const GnParams& P, cudaStream_t st) {
  using Smem = GnSmem<MH, BT>;
  auto kern = fastgen_gn_kernel<MH, BT>;
  NSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  GnParams p = P;
  void* args[] = {&p};
  NSW_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(GN_NC), dim3(GN_NT), args, sizeof(Smem), st));
  count_launch();
  return NSW_OK;
}

}  // namespace

struct GnEngine {
  nsw_wavenet_config cfg;
  int device = 0, L = 0, O = 0, MH = 0, N = 0, NPL = 0;
  size_t hist_entries = 0;
  int l2_last = 0;
  int chunk = 2048;
  DevBuf blocks, wcs, bcs, cond_w, cond_wt_hi, cond_wt_lo, cond_b, hist_off, dil;
  DevBuf hist, gbuf, sbuf, hbuf, obuf, xstate, bar, cond, enc_split, dbg;
};

int gn_create(const nsw_wavenet_config& cfg, const TensorMap& tm, int device, GnEngine** out) {
  GnPacked pk;
  NSW_TRY(gn_pack(cfg, tm, pk));
  NSW_CUDA(cudaSetDevice(device));
  int coop = 0, sms = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  NSW_CHECK(coop && sms >= GN_NC, NSW_EINVAL, "fastgen needs cooperative launch and >= %d SMs (device has %d)", GN_NC,
            sms);
  GnEngine* g = new GnEngine();
  g->cfg = cfg;
  g->device = device;
  g->L = pk.L;
  g->O = pk.O;
  g->MH = pk.lay.MH;
  g->N = pk.N;
  g->NPL = pk.N / 64;
  g->hist_entries = pk.hist_entries;
  if (const char* e = getenv("NSW_FASTGEN_CHUNK")) g->chunk = std::max(1, atoi(e));
  int l2_bytes = 0;
  cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, device);
  g->l2_last = std::min(pk.L + 3, (int)(0.6 * (double)l2_bytes / ((double)GN_NC * pk.lay.block_floats * 4)));
  int rc = NSW_OK;
  auto up = [&](DevBuf& b, const void* p, size_t bytes) { if (rc == NSW_OK) rc = upload(b, p, bytes); };
  up(g->blocks, pk.blocks.data(), pk.blocks.size() * 4);
  up(g->wcs, pk.wcs.data(), pk.wcs.size() * 4);
  up(g->bcs, pk.bcs.data(), pk.bcs.size() * 4);
  up(g->cond_b, pk.cond_b.data(), pk.cond_b.size() * 4);
  up(g->hist_off, pk.hist_off.data(), pk.hist_off.size() * 4);
  up(g->dil, pk.dil.data(), pk.dil.size() * 4);
  if (cfg.engine >= NSW_ENGINE_TC) {
    std::vector<float> bt((size_t)pk.N * GN_D);
    for (int k = 0; k < GN_D; ++k)
      for (int n = 0; n < pk.N; ++n) bt[(size_t)n * GN_D + k] = pk.cond_w[(size_t)k * pk.N + n];
    std::vector<__half> hi(bt.size()), lo(bt.size());
    split_f16(bt.data(), bt.size(), hi.data(), lo.data());
    up(g->cond_wt_hi, hi.data(), hi.size() * 2);
    up(g->cond_wt_lo, lo.data(), lo.size() * 2);
  } else {
    up(g->cond_w, pk.cond_w.data(), pk.cond_w.size() * 4);
  }
  if (rc != NSW_OK) {
    delete g;
    return rc;
  }
  *out = g;
  return NSW_OK;
}

void gn_destroy(GnEngine* g) { delete g; }

int gn_run(GnEngine* g, const float* d_encoding, int B, int T, const float* d_tf, uint64_t seed, const float* d_noise,
           int nu, float* d_audio, float* d_out, cudaStream_t st) {
  const int L = g->L, O = g->O, MH = g->MH;
  const bool tc = g->cfg.engine >= NSW_ENGINE_TC;
  for (int b0 = 0; b0 < B; b0 += GN_MAX_BT) {
    const int nb = std::min(GN_MAX_BT, B - b0);
    const int BT = nb <= 1 ? 1 : nb <= 2 ? 2 : nb <= 4 ? 4 : 8;
    const int TC = std::min(g->chunk, T);
    const size_t cond_bstride = (size_t)g->NPL * TC * 64;
    NSW_TRY(g->cond.ensure(cond_bstride * BT * sizeof(float)));
    NSW_TRY(g->hist.ensure(g->hist_entries * BT * GN_W * sizeof(float)));
    NSW_TRY(g->gbuf.ensure((size_t)2 * BT * MH * sizeof(float)));
    NSW_TRY(g->sbuf.ensure((size_t)BT * GN_S * sizeof(float)));
    NSW_TRY(g->hbuf.ensure((size_t)BT * GN_S * sizeof(float)));
    NSW_TRY(g->obuf.ensure((size_t)BT * GN_MAX_O * sizeof(float)));
    NSW_TRY(g->xstate.ensure((size_t)GN_MAX_BT * 4 * sizeof(float)));
    NSW_TRY(g->bar.ensure(256));
    if (tc) NSW_TRY(g->enc_split.ensure((size_t)TC * GN_D * 2 * sizeof(__half)));
    // queues start at zero (fastgen.py:150) and the first fed sample is 0 (:153); ring entries older than the
    // utterance are never read (causal guards), the conditioning rows of dead batch lanes must be finite
    NSW_CUDA(cudaMemsetAsync(g->xstate.p, 0, g->xstate.bytes, st));
    if (nb < BT) NSW_CUDA(cudaMemsetAsync(g->cond.p, 0, cond_bstride * BT * sizeof(float), st));
    for (int t0 = 0; t0 < T; t0 += TC) {
      const int tcn = std::min(TC, T - t0);
      for (int b = 0; b < nb; ++b) {
        // hoisted mel conditioning of rows [t0, t0 + tcn): [tcn, 256] x [256, L*G + 256], NO centre trim
        // (fastgen.py:157 feeds encoding[:, i] directly)
        const float* enc = d_encoding + ((size_t)(b0 + b) * T + t0) * GN_D;
        ConvGemm cg;
        cg.nclips = 1; cg.L = tcn; cg.cin = GN_D; cg.ntaps = 1; cg.a_off = 0; cg.mclip = tcn; cg.N = g->N;
        EpiParams e{};
        e.mode = EPI_PLANES;
        e.bias = g->cond_b.as<float>();
        e.out_f32 = g->cond.as<float>() + (size_t)b * cond_bstride;
        if (tc) {
          __half* hi = g->enc_split.as<__half>();
          __half* lo = hi + (size_t)TC * GN_D;
          const size_t ne = (size_t)tcn * GN_D;
          gn_split_f16_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(enc, hi, lo, ne);
          count_launch();
          NSW_TRY(conv_gemm_tc(cg, hi, lo, g->cond_wt_hi.as<__half>(), g->cond_wt_lo.as<__half>(), e, st));
        } else {
          NSW_TRY(conv_gemm_ffma(cg, enc, g->cond_w.as<float>(), e, st));
        }
      }
      NSW_CUDA(cudaMemsetAsync(g->bar.p, 0, 256, st));
      GnParams P;
      P.blocks = g->blocks.as<float>();
      P.cond = g->cond.as<float>();
      P.cond_bstride = cond_bstride;
      P.Tc = tcn;
      P.hist = g->hist.as<float>();
      P.hist_off = g->hist_off.as<int>();
      P.dil = g->dil.as<int>();
      P.gbuf = g->gbuf.as<float>();
      P.sbuf = g->sbuf.as<float>();
      P.hbuf = g->hbuf.as<float>();
      P.obuf = g->obuf.as<float>();
      P.xstate = g->xstate.as<float>();
      P.bar = g->bar.as<unsigned int>();
      P.wcs = g->wcs.as<float>();
      P.bcs = g->bcs.as<float>();
      P.tf = d_tf ? d_tf + (size_t)b0 * T : nullptr;
      P.noise = d_noise ? d_noise + (size_t)b0 * T * nu : nullptr;
      P.audio = d_audio ? d_audio + (size_t)b0 * T : nullptr;
      P.out = d_out ? d_out + (size_t)b0 * T * O : nullptr;
      P.nb = nb;
      P.T = T; P.t0 = t0; P.t1 = t0 + tcn;
      P.L = L; P.O = O; P.loss_type = g->cfg.loss_type; P.use_mu_law = g->cfg.use_mu_law ? 1 : 0; P.nu = nu;
      P.seed = seed; P.b0 = b0;
      P.quant = g->cfg.use_mu_law ? 256.0f : 65536.0f;
      P.dbg = nullptr;
      P.flags = getenv("NSW_GN_FLAGS") ? atoi(getenv("NSW_GN_FLAGS")) : 0;
      P.l2_last = getenv("NSW_FASTGEN_L2LAST") ? atoi(getenv("NSW_FASTGEN_L2LAST")) : g->l2_last;
      const bool want_dbg = getenv("NSW_FASTGEN_DEBUG") != nullptr;
      if (want_dbg) {
        NSW_TRY(g->dbg.ensure(GN_NC * 8 * sizeof(long long)));
        NSW_CUDA(cudaMemsetAsync(g->dbg.p, 0, GN_NC * 8 * sizeof(long long), st));
        P.dbg = g->dbg.as<long long>();
      }
      int rc;
      if (MH == 256) {
        rc = BT == 1 ? gn_launch<256, 1>(P, st) : BT == 2 ? gn_launch<256, 2>(P, st)
             : BT == 4 ? gn_launch<256, 4>(P, st) : gn_launch<256, 8>(P, st);
      } else {
        rc = BT == 1 ? gn_launch<512, 1>(P, st) : BT == 2 ? gn_launch<512, 2>(P, st)
             : BT == 4 ? gn_launch<512, 4>(P, st) : gn_launch<512, 8>(P, st);
      }
      NSW_TRY(rc);
      if (want_dbg) {
        std::vector<long long> host(GN_NC * 8);
        NSW_CUDA(cudaStreamSynchronize(st));
        NSW_CUDA(cudaMemcpy(host.data(), g->dbg.p, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const double phases = (double)tcn * (L + 3);
        const char* names[5] = {"weights + past taps (to S1)", "fresh inputs arrived", "dots", "epilogue", "arrive"};
        for (int cta : {0, 1, 64, 127}) {
          fprintf(stderr, "[nsw fastgen gn dbg BT=%d MH=%d] cta %3d cycles/phase:", BT, MH, cta);
          for (int i = 0; i < 5; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)host[8 * cta + i] / phases);
          fprintf(stderr, "\n");
        }
      }
    }
  }
  return NSW_OK;
}

}  // namespace nsw

// TEST HOOK (host only, no CUDA): the batched engine's create-time repacking.  sizes[0] = floats in `blocks`,
// [1] = floats per block, [2] = CTAs, [3] = columns of cond_w, [4] = MH, [5] = phases.
extern "C" int nsw_fastgen_gn_pack_host(const nsw_wavenet_config* cfg, const nsw_tensor* tensors, int32_t n,
                                        float* blocks, int64_t blocks_cap, float* cond_w, float* cond_b,
                                        int64_t* sizes) {
  using namespace nsw;
  NSW_CHECK(cfg && tensors && sizes, NSW_EINVAL, "null argument");
  TensorMap tm(tensors, n);
  GnPacked pk;
  NSW_TRY(gn_pack(*cfg, tm, pk));
  sizes[0] = (int64_t)pk.blocks.size();
  sizes[1] = pk.lay.block_floats;
  sizes[2] = GN_NC;
  sizes[3] = pk.N;
  sizes[4] = pk.lay.MH;
  sizes[5] = pk.L + 3;
  if (blocks) {
    NSW_CHECK(blocks_cap >= (int64_t)pk.blocks.size(), NSW_EINVAL, "blocks buffer too small");
    memcpy(blocks, pk.blocks.data(), pk.blocks.size() * sizeof(float));
  }
  if (cond_w) memcpy(cond_w, pk.cond_w.data(), pk.cond_w.size() * sizeof(float));
  if (cond_b) memcpy(cond_b, pk.cond_b.data(), pk.cond_b.size() * sizeof(float));
  return NSW_OK;
}

NSW_RANGE_GUARD_TU(fastgen_gn)
