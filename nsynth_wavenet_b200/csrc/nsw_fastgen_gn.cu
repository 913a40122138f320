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
};

template <int MH, int BT>
struct GnSmem {
  float xin[BT][GN_W + MH];  // fresh inputs of a phase: [l_{ph-2} | g_{ph-1}]; s / h in the first 256 for the head
  float hv[BT][2 * GN_W];    // [l[t-2d] | l[t-d]] of the phase's layer
  float wcs[3 * GN_W];
  float bcs[GN_W];
  float outv[BT][GN_MAX_O];
  float rsum[16][BT];
  float pv[8][BT];
  float cnd[8][BT];
  float lst[4][BT];          // owned residual channels 4c..4c+3
  float sst[2][BT];          // owned skip channels 2c, 2c+1
  float xn[BT], x1[BT], x2[BT];
  int dil[GN_MAX_L + 4], hoff[GN_MAX_L + 4], pos[GN_MAX_L + 4];
};

__device__ __noinline__ void gn_die(const char* what) {
  printf("nsw fastgen(gn): watchdog in %s (block %d)\n", what, blockIdx.x);
  __trap();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// acc[b] = sum_i w[i] . x[b][128 i + 4 lane .. +3]  (per-lane partial; x rows are xs floats apart)
template <int NF, int BT>
__device__ __forceinline__ void dot_bt(const float4* w, const float* x, int xs, int lane, float (&acc)[BT]) {
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const float4 xx = *reinterpret_cast<const float4*>(x + (size_t)b * xs + 128 * i + 4 * lane);
      a0 = fmaf(w[i].x, xx.x, a0);
      a1 = fmaf(w[i].y, xx.y, a1);
      a0 = fmaf(w[i].z, xx.z, a0);
      a1 = fmaf(w[i].w, xx.w, a1);
    }
    acc[b] = a0 + a1;
  }
}
template <int BT>
__device__ __forceinline__ void dot_n(int nf, const float4* w, const float* x, int xs, int lane, float (&acc)[BT]) {
  switch (nf) {
    case 2: dot_bt<2, BT>(w, x, xs, lane, acc); break;
    case 4: dot_bt<4, BT>(w, x, xs, lane, acc); break;
    case 6: dot_bt<6, BT>(w, x, xs, lane, acc); break;
    default: dot_bt<8, BT>(w, x, xs, lane, acc); break;
  }
}
template <int BT>
__device__ __forceinline__ void reduce_store(float (&acc)[BT], float* dst, int lane) {
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    float v = acc[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) dst[b] = v;
  }
}

__device__ __forceinline__ float gn_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gn_tanh(float x) { return 2.0f * gn_sigmoid(2.0f * x) - 1.0f; }

template <int MH, int BT>
__global__ void __launch_bounds__(GN_NT, 1) fastgen_gn_kernel(GnParams P) {
  constexpr int W = GN_W, S = GN_S, XS = GN_W + MH, NT = GN_NT, NC = GN_NC;
  constexpr int PPC = MH / NC, nD = 2 * PPC, K1 = W + MH, G = 2 * MH;
  constexpr int OFF_D = 0, OFF_P = OFF_D + nD * K1, OFF_L = OFF_P + nD * 2 * W, OFF_S = OFF_L + 4 * MH,
                OFF_C = OFF_S + 2 * W, BF = OFF_C + 16;
  constexpr int NFD = K1 / 128, NFM = MH / 128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = GnSmem<MH, BT>;
  Smem& Sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
  const int L = P.L, O = P.O, T = P.T, NPH = L + 3;
  const int RO = (O + NC - 1) / NC;

  for (int i = tid; i < 3 * W; i += NT) Sm.wcs[i] = P.wcs[i];
  for (int i = tid; i < W; i += NT) Sm.bcs[i] = P.bcs[i];
  for (int i = tid; i < BT * XS; i += NT) (&Sm.xin[0][0])[i] = 0.f;
  for (int i = tid; i < BT * 2 * W; i += NT) (&Sm.hv[0][0])[i] = 0.f;
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
  __syncthreads();

  unsigned int bar_target = 0;
  auto arrive = [&]() {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(P.bar, 1u);
    }
    bar_target += NC;
  };
  auto wait = [&]() {
    if (tid == 0) {
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
      __threadfence();
    }
    __syncthreads();
  };

  const size_t ring_entry = (size_t)BT * W;
  for (int t = P.t0; t < P.t1; ++t) {
    const int tl = t - P.t0;
    for (int ph = 1; ph <= NPH; ++ph) {
      const float* blk = P.blocks + ((size_t)(ph - 1) * NC + c) * BF;
      if (tid == 0) {  // pull the block two phases ahead into L2 (the stream is larger than L2: plain LRU never hits)
        const int pn = ph + 1 >= NPH ? ph + 1 - NPH : ph + 1;  // 0-based index of phase ph + 2
        const float* nxt = P.blocks + ((size_t)pn * NC + c) * BF;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"((uint32_t)(BF * 4)) : "memory");
      }
      // ------------------------------- before the barrier in flight -------------------------------
      float4 wp[8], wf[8];
      int nf = 0, xoff = 0;  // fresh row of this warp: nf float4 per lane against xin[b][xoff ...]
      const bool has_past = ph <= L && warp < nD;
      if (has_past) {
#pragma unroll
        for (int i = 0; i < 8; ++i) wp[i] = __ldg(reinterpret_cast<const float4*>(blk + OFF_P + warp * 2 * W + 128 * i + 4 * lane));
      }
      {
        const float* row = nullptr;
        if (ph <= L) {
          if (warp < nD) { row = blk + OFF_D + warp * K1; nf = NFD; xoff = 0; }
          else if (warp < nD + 4) { if (ph >= 2) { row = blk + OFF_L + (warp - nD) * MH; nf = NFM; xoff = W; } }
          else if (warp < nD + 6) {
            row = blk + OFF_S + (warp - nD - 4) * W;
            if (ph == 1) { nf = W / 128; xoff = 0; } else { nf = NFM; xoff = W; }
          }
        } else if (ph == L + 1) {
          if (warp >= nD + 4 && warp < nD + 6) { row = blk + OFF_S + (warp - nD - 4) * W; nf = NFM; xoff = W; }
        } else if (ph == L + 2) {
          if (warp < 2) { row = blk + OFF_D + warp * K1; nf = S / 128; xoff = 0; }
        } else {
          if (warp < RO && c * RO + warp < O) { row = blk + OFF_D + warp * K1; nf = S / 128; xoff = 0; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < nf) wf[i] = __ldg(reinterpret_cast<const float4*>(row + 128 * i + 4 * lane));
      }
      if (ph <= L) {
        const int d = Sm.dil[ph], R = 2 * d + 1;
        const int pos = Sm.pos[ph];
        int p1 = pos - d; if (p1 < 0) p1 += R;
        int p2 = p1 - d; if (p2 < 0) p2 += R;
        const float* ring = P.hist + (size_t)Sm.hoff[ph] * ring_entry;
        const bool ok2 = t - 2 * d >= 0, ok1 = t - d >= 0;
        for (int i = tid; i < BT * (W / 4); i += NT) {  // one float4 of each tap per iteration
          const int b = i / (W / 4), k4 = i - b * (W / 4);
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a2 = ok2 ? ldcg4(ring + (size_t)p2 * ring_entry + (size_t)b * W + 4 * k4) : z;
          const float4 a1 = ok1 ? ldcg4(ring + (size_t)p1 * ring_entry + (size_t)b * W + 4 * k4) : z;
          *reinterpret_cast<float4*>(&Sm.hv[b][4 * k4]) = a2;
          *reinterpret_cast<float4*>(&Sm.hv[b][W + 4 * k4]) = a1;
        }
        if (tid < nD * BT) {
          const int j = tid / BT, b = tid - j * BT;
          const int n = (ph - 1) * G + c * nD + j;
          Sm.cnd[j][b] = __ldg(P.cond + (size_t)b * P.cond_bstride + ((size_t)(n >> 6) * P.Tc + tl) * 64 + (n & 63));
        }
      } else if (ph == L + 2) {
        if (tid < 2 * BT) {
          const int j = tid / BT, b = tid - j * BT;
          const int n = L * G + 2 * c + j;
          Sm.cnd[j][b] = __ldg(P.cond + (size_t)b * P.cond_bstride + ((size_t)(n >> 6) * P.Tc + tl) * 64 + (n & 63));
        }
      }
      __syncthreads();  // hv / cnd staged
      if (has_past) {
        float acc[BT];
        dot_bt<8, BT>(wp, &Sm.hv[0][0], 2 * W, lane, acc);
        reduce_store<BT>(acc, Sm.pv[warp], lane);
      }
      // ------------------------------- the exchange of the previous phase -------------------------------
      wait();
      if (ph == 1) {
        // conv_start on the fed-back sample, every CTA the full vector (causal_linear rate 1, masked.py:352-376)
        for (int i = tid; i < BT * W; i += NT) {
          const int b = i / W, k = i - b * W;
          Sm.xin[b][k] = fmaf(Sm.wcs[2 * W + k], Sm.xn[b], fmaf(Sm.wcs[W + k], Sm.x1[b], fmaf(Sm.wcs[k], Sm.x2[b], Sm.bcs[k])));
        }
        for (int i = tid; i < BT * MH; i += NT) Sm.xin[i / MH][W + i % MH] = 0.f;  // no gate output before layer 1
      } else if (ph <= L) {
        const float* lsrc = P.hist + ((size_t)Sm.hoff[ph - 1] + Sm.pos[ph - 1]) * ring_entry;  // l_{ph-2} of this step
        const float* gsrc = P.gbuf + (size_t)((ph - 1) & 1) * BT * MH;
        for (int i = tid; i < BT * (W / 4); i += NT) {
          const int b = i / (W / 4), k4 = i - b * (W / 4);
          *reinterpret_cast<float4*>(&Sm.xin[b][4 * k4]) = ldcg4(lsrc + (size_t)b * W + 4 * k4);
        }
        for (int i = tid; i < BT * (MH / 4); i += NT) {
          const int b = i / (MH / 4), k4 = i - b * (MH / 4);
          *reinterpret_cast<float4*>(&Sm.xin[b][W + 4 * k4]) = ldcg4(gsrc + (size_t)b * MH + 4 * k4);
        }
      } else if (ph == L + 1) {
        const float* gsrc = P.gbuf + (size_t)(L & 1) * BT * MH;
        for (int i = tid; i < BT * (MH / 4); i += NT) {
          const int b = i / (MH / 4), k4 = i - b * (MH / 4);
          *reinterpret_cast<float4*>(&Sm.xin[b][W + 4 * k4]) = ldcg4(gsrc + (size_t)b * MH + 4 * k4);
        }
      } else {
        const float* src = ph == L + 2 ? P.sbuf : P.hbuf;
        for (int i = tid; i < BT * (S / 4); i += NT) {
          const int b = i / (S / 4), k4 = i - b * (S / 4);
          *reinterpret_cast<float4*>(&Sm.xin[b][4 * k4]) = ldcg4(src + (size_t)b * S + 4 * k4);
        }
      }
      __syncthreads();
      if (ph == 1 && tid < BT) {  // conv_start's two queues (rate 1)
        Sm.x2[tid] = Sm.x1[tid];
        Sm.x1[tid] = Sm.xn[tid];
      }
      if (nf > 0) {
        float acc[BT];
        dot_n<BT>(nf, wf, &Sm.xin[0][0] + xoff, XS, lane, acc);
        reduce_store<BT>(acc, Sm.rsum[warp], lane);
      }
      __syncthreads();
      // ------------------------------- epilogues: one thread per (row, batch row) -------------------------------
      if (ph <= L) {
        if (tid < PPC * BT) {
          const int pr = tid / BT, b = tid - pr * BT;
          const float a = Sm.rsum[pr][b] + Sm.pv[pr][b] + Sm.cnd[pr][b];
          const float q = Sm.rsum[PPC + pr][b] + Sm.pv[PPC + pr][b] + Sm.cnd[PPC + pr][b];
          __stcg(P.gbuf + (size_t)(ph & 1) * BT * MH + (size_t)b * MH + c * PPC + pr, gn_sigmoid(a) * gn_tanh(q));
        } else if (tid >= 64 && tid < 64 + 4 * BT) {
          const int r = (tid - 64) / BT, b = (tid - 64) - r * BT;
          // l_{ph-1} = l_{ph-2} + Wr_{ph-1} g_{ph-1} + br_{ph-1}  (phase 1: l_0 from conv_start)
          const float l = ph == 1 ? Sm.xin[b][4 * c + r] : Sm.lst[r][b] + Sm.rsum[nD + r][b] + __ldg(blk + OFF_C + r);
          Sm.lst[r][b] = l;
          __stcg(P.hist + ((size_t)Sm.hoff[ph] + Sm.pos[ph]) * ring_entry + (size_t)b * W + 4 * c + r, l);
        } else if (tid >= 128 && tid < 128 + 2 * BT) {
          const int r = (tid - 128) / BT, b = (tid - 128) - r * BT;
          const float a = Sm.rsum[nD + 4 + r][b] + __ldg(blk + OFF_C + 4 + r);
          Sm.sst[r][b] = ph == 1 ? a : Sm.sst[r][b] + a;  // skip_start, then skip_{ph-1}
        }
      } else if (ph == L + 1) {
        if (tid >= 128 && tid < 128 + 2 * BT) {
          const int r = (tid - 128) / BT, b = (tid - 128) - r * BT;
          const float s = Sm.sst[r][b] + Sm.rsum[nD + 4 + r][b] + __ldg(blk + OFF_C + 4 + r);
          __stcg(P.sbuf + (size_t)b * S + 2 * c + r, fmaxf(s, 0.f));  // relu(s) (wavenet.py:494)
        }
      } else if (ph == L + 2) {
        if (tid < 2 * BT) {
          const int r = tid / BT, b = tid - r * BT;
          __stcg(P.hbuf + (size_t)b * S + 2 * c + r, fmaxf(Sm.rsum[r][b] + Sm.cnd[r][b], 0.f));
        }
      } else {
        if (tid < RO * BT) {
          const int r = tid / BT, b = tid - r * BT;
          const int o = c * RO + r;
          if (o < O) __stcg(P.obuf + (size_t)b * O + o, Sm.rsum[r][b] + __ldg(blk + OFF_C + 6 + r));
        }
      }
      arrive();
    }
    // ------------------------------- head: every CTA samples every row, redundantly -------------------------------
    wait();
    for (int i = tid; i < BT * O; i += NT) {
      const int b = i / O, o = i - b * O;
      Sm.outv[b][o] = __ldcg(P.obuf + (size_t)b * O + o);
    }
    __syncthreads();
    if (warp < BT) {
      const int b = warp;
      const float* ov = Sm.outv[b];
      const float Q = P.quant;
      const unsigned long long sd = P.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(P.b0 + b);
      const uint2 key = make_uint2((uint32_t)sd, (uint32_t)(sd >> 32));
      const float* nz = (P.noise && b < P.nb) ? P.noise + ((size_t)b * T + t) * P.nu : nullptr;
      auto bits_of = [&](int j) -> uint32_t {  // same stream as the latency engine: counter (t, j / 4), word j % 4
        const uint4 rr = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)(j >> 2), 0x66617374u, 0u), key);
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
        const float fed = (P.tf && b < P.nb) ? P.tf[(size_t)b * T + t] : a;
        Sm.xn[b] = P.use_mu_law ? mu_law_scaled_dev(fed, Q) : fed;  // wavenet.py:411-414
        if (c == 0 && P.audio && b < P.nb) P.audio[(size_t)b * T + t] = fed;
      }
    }
    if (c == 0 && P.out) {
      for (int i = tid; i < P.nb * O; i += NT) {
        const int b = i / O, o = i - b * O;
        P.out[((size_t)b * T + t) * O + o] = Sm.outv[b][o];
      }
    }
    if (tid >= 1 && tid <= L) {
      const int R = 2 * Sm.dil[tid] + 1;
      const int pn = Sm.pos[tid] + 1;
      Sm.pos[tid] = pn >= R ? 0 : pn;
    }
    __syncthreads();
  }
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
  int chunk = 2048;
  DevBuf blocks, wcs, bcs, cond_w, cond_wt_hi, cond_wt_lo, cond_b, hist_off, dil;
  DevBuf hist, gbuf, sbuf, hbuf, obuf, xstate, bar, cond, enc_split;
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
      int rc;
      if (MH == 256) {
        rc = BT == 1 ? gn_launch<256, 1>(P, st) : BT == 2 ? gn_launch<256, 2>(P, st)
             : BT == 4 ? gn_launch<256, 4>(P, st) : gn_launch<256, 8>(P, st);
      } else {
        rc = BT == 1 ? gn_launch<512, 1>(P, st) : BT == 2 ? gn_launch<512, 2>(P, st)
             : BT == 4 ? gn_launch<512, 4>(P, st) : gn_launch<512, 8>(P, st);
      }
      NSW_TRY(rc);
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
