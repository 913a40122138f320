// Internal interface between the two autoregressive fastgen engines.
//
//   nsw_fastgen.cu     latency engine: batch 1, width 512 / gate 512 / skip 256, MoL or Gauss head; one exchange per
//                      layer through tagged words, no grid barrier (41 us per sample).
//   nsw_fastgen_gn.cu  batched engine ("gn"): gate 512 or 1024 (double_gate_width, wavenet.py:106), MoL / Gauss / CE
//                      heads, mu-law input, BT batch rows per weight pass; one grid barrier per layer.  Weights are the
//                      cost of an autoregressive step, so B rows cost about as much as one.
#pragma once
#include "nsw_gemm.cuh"

namespace nsw {

constexpr int GN_W = 512;    // width (residual channels)
constexpr int GN_S = 256;    // skip width
constexpr int GN_D = 256;    // deconv width (conditioning channels)
constexpr int GN_NC = 128;   // CTAs
constexpr int GN_NT = 512;   // threads per CTA
constexpr int GN_MAX_O = 256;
constexpr int GN_MAX_BT = 8;  // batch rows per weight pass
constexpr int GN_MAX_L = 60;

// per-(phase, CTA) weight block of the batched engine, in floats (MH = gate_width / 2)
struct GnLayout {
  int MH, PPC, nD, K1;
  int off_d, off_p, off_l, off_s, off_c, block_floats;
  __host__ __device__ static constexpr GnLayout make(int mh) {
    GnLayout g{};
    g.MH = mh;
    g.PPC = mh / GN_NC;          // gate pairs per CTA
    g.nD = 2 * g.PPC;            // dilated-conv rows per CTA: [sigmoid rows | tanh rows]
    g.K1 = GN_W + mh;            // [W2 row | M row]
    g.off_d = 0;
    g.off_p = g.off_d + g.nD * g.K1;       // nD rows x [W0 | W1]
    g.off_l = g.off_p + g.nD * 2 * GN_W;   // 4 rows x MH: Wr of the previous layer
    g.off_s = g.off_l + 4 * mh;            // 2 rows x W: Ws of the previous layer (first MH) / skip_start (W)
    g.off_c = g.off_s + 2 * GN_W;          // 16 consts: br[4], bs[2], out2 bias[2], pad
    g.block_floats = g.off_c + 16;
    return g;
  }
};

struct GnPacked {
  GnLayout lay;
  int L = 0, O = 0, N = 0;  // N: columns of the hoisted conditioning GEMM (L * G + S)
  std::vector<float> blocks;  // [L + 3][GN_NC][block_floats]
  std::vector<float> wcs, bcs, cond_w, cond_b;
  std::vector<int> dil, hist_off;  // [L + 1]; hist_off in ring ENTRIES (an entry is [B][W] floats)
  size_t hist_entries = 0;
};

bool gn_supported(const nsw_wavenet_config& c, const char** why);
int gn_pack(const nsw_wavenet_config& cfg, const TensorMap& tm, GnPacked& pk);

struct GnEngine;
int gn_create(const nsw_wavenet_config& cfg, const TensorMap& tm, int device, GnEngine** out);
void gn_destroy(GnEngine* g);
// d_encoding [B][T][256]; d_tf [B][T] or NULL; d_noise [B][T][nu] or NULL; outputs may be NULL
int gn_run(GnEngine* g, const float* d_encoding, int B, int T, const float* d_tf, uint64_t seed, const float* d_noise,
           int nu, float* d_audio, float* d_out, cudaStream_t stream);

}  // namespace nsw
