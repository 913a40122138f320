// "Conv-GEMM": the one dense contraction shape the generation path needs.
//
//   C[(clip, m), n] = sum_{tap < ntaps} sum_{c < cin}
//                        X[clip, m + a_off + tap, c] * Bw[tap*cin + c, n]
//   with X[clip, f, :] = 0 for f outside [0, L).
//
// * mel-conditioning 1x1 projections (masked.conv1d with filter_length 1 on mel_en,
//   parallel_wavenet.py:237-243,259-261; wavenet.py:255-261,283-284): ntaps = 1,
//   a_off = centre-trim offset of wavenet._condition (wavenet.py:76-85), all
//   layers of a flow batched along n.
// * transposed-conv upsampling (masked.trans_conv1d, masked.py:235-291): output
//   phase r and input frame m give o = m*s + r - p; ntaps = k/s, a_off = -(ntaps-1),
//   n = r*cout + co.
//
// Two engines implement it: fp32 CUDA cores (nsw_gemm_ffma.cu) and tcgen05
// split-fp16 tensor cores (nsw_gemm_tc.cu).
#pragma once
#include "nsw_common.cuh"

namespace nsw {

struct ConvGemm {
  int nclips;  // batch
  int L;       // valid frames per clip in X
  int cin;     // channels of X
  int ntaps;
  int a_off;
  int mclip;  // GEMM rows per clip
  int N;      // GEMM columns (multiple of 64)
  int tap_stride = 1;  // frame = m + a_off + tap * tap_stride (dilated causal conv: stride = d)
  int x_pitch = 0;     // elements between consecutive frames of X (0 = cin)
  // optional second source in front of K (tensor-core pair kernel only): cin2 more channels read from
  // X2[clip, m + a_off2, :] (X2 is [nclips, L2, cin2]); Bt has cin2 + ntaps*cin columns, the second source's first.  The
  // teacher's dilated conv takes its mel conditioning this way (K = 256 + 3*512) instead of adding a separately
  // projected plane.
  int cin2 = 0, a_off2 = 0, L2 = 0;
  // pair kernel: the two small products of the split (lo*hi, hi*lo) accumulate in their own TMEM columns and are added
  // to the hi*hi sum by the epilogue.  tcgen05 accumulates in fp32 with truncation, an error proportional to the
  // accumulator's magnitude per instruction; with the small products kept out, the large accumulator takes a third of
  // the instructions (teacher dilated conv, K = 1792: 112 instead of 336) and the small one is 2^-11 of its size.
  // Costs the second accumulator stage (the epilogue no longer overlaps the next item's MMAs).
  int split_acc = 0;
  // pair kernel: C += Y for a matrix Y held as a split pair [nclips*mclip, ld3] (fp16 hi + lo): its 256-column block of
  // the work item rides through the tensor core as four more k-blocks against a 64 x 64 identity (N = 64 MMAs into the
  // matching column slice of the accumulator; hi*1 + lo*1, exact), fetched by TMA like every other operand.  Replaces the epilogue's read-back of the fp32 rows it accumulates onto --
  // a chain of global loads per 32-column chunk with a few KB in flight per warp, which is what the teacher's
  // residual/skip GEMM (K = 256: 6 k cycles of MMAs per item) spent 40 k cycles per item on.  In-place (Y = the split
  // output of the same call) is fine: an item reads exactly the block it later writes.
  int acc3 = 0, ld3 = 0;
  int one_cta = 0;  // force the 128 x 128 one-CTA kernel (parity hook / A-B runs)
};

// EPI_ROWS : out[row * ld_out + n] (+)= D + bias[n] (+ addend[row * ld_add + n]); optional relu;
//            optional fp16 hi/lo copy of the result (relu'd for n >= relu_split_from)
// EPI_GATE : columns are gate-interleaved (2j = sigmoid half, 2j+1 = tanh half of gate j);
//            g[row * ld_split + n/2 + j] = sigmoid(D[2j] + c[2j]) * tanh(D[2j+1] + c[2j+1]),
//            c = addend[row * ld_add + n + ...] (if given) + bias[n + ...] (if given); g is written as fp16 hi/lo only
enum EpiMode { EPI_PLANES = 0, EPI_DECONV = 1, EPI_ROWS = 2, EPI_GATE = 3 };

struct EpiParams {
  int mode;
  const float* bias;  // EPI_PLANES: [N]; EPI_DECONV: [cout]
  // EPI_PLANES: out_f32[n/64][clip*mclip + m][n%64]
  // EPI_DECONV: out_*[clip][o][co], o = m*s + n/cout - p in [0, Lout)
  float* out_f32;
  __half* out_hi;  // optional fp16 split of the same values (hi + lo ~ fp32)
  __half* out_lo;
  int s, p, cout, act, Lout;
  // EPI_ROWS / EPI_GATE
  const float* addend = nullptr;
  int ld_out = 0, ld_add = 0, ld_split = 0;
  int accumulate = 0, relu_out = 0, relu_split_from = 1 << 30;
  // EPI_PLANES: planes [0, tiled_planes) are written row-interleaved for the tc3 layer kernel:
  //   plane[((row/128)*8 + (row%128)/32*2 + (n%64)/32) * 1024 + ((n%32)/4)*128 + (row%32)*4 + n%4]
  // (a warp-level 16-byte access of one column group is 512 contiguous bytes); needs mclip % 128 == 0
  int tiled_planes = 0;
  // EPI_DECONV: rows per clip of the OUTPUT buffers (0 = Lout): an intermediate of the upsampling stack is written with
  // zero rows after every clip so that the next layer can run over all clips as ONE flattened clip (no 128-row tile
  // padding per clip: 8 x 393 rows are 25 tiles instead of 32), its taps reading the zero rows at the clip boundaries.
  int out_clip_rows = 0;
  // EPI_DECONV on such a flattened input: GEMM row m' = clip * flat_rows + m (0: rows are (clip, m) as usual)
  int flat_rows = 0;
};

// fp32 X [nclips, L, cin], fp32 Bw [ntaps*cin, N]
int conv_gemm_ffma(const ConvGemm& g, const float* X, const float* Bw, const EpiParams& e,
                   cudaStream_t stream);

// fp16 split operands: X_hi/X_lo [nclips, L, cin], Bt_hi/Bt_lo [N, ntaps*cin] (K-major)
int conv_gemm_tc(const ConvGemm& g, const __half* X_hi, const __half* X_lo,
                 const __half* Bt_hi, const __half* Bt_lo, const EpiParams& e,
                 cudaStream_t stream, const __half* X2_hi = nullptr, const __half* X2_lo = nullptr,
                 const __half* Y_hi = nullptr, const __half* Y_lo = nullptr);
bool conv_gemm_tc_supported(const ConvGemm& g);
// X-resident projection of [nclips, L, 256] activations; ALL output planes row-interleaved (tc3 engine)
int cond_proj_tc(int nclips, int L, int mclip, int a_off, int N, const __half* X_hi, const __half* X_lo,
                 const __half* Bt_hi, const __half* Bt_lo, const float* bias, float* out_tiled,
                 cudaStream_t stream);

// host-side split of fp32 into fp16 hi + fp16 lo (round-to-nearest each)
void split_f16(const float* src, size_t n, __half* hi, __half* lo);

// ---- tcgen05 fused IAF residual layer (nsw_iaf_layer_tc.cu); maps are 128-byte CUtensorMaps ----
int layer_tc_make_act_map(void* map_out, const __half* base, int B, int T);
int layer_tc_make_weight_map(void* map_out, const __half* base, int rows, int k, int box_rows = 64);
int layer_tc_launch(const void* const map_act[2][2], const void* map_wdh, const void* map_wdl,
                    const void* map_wrh, const void* map_wrl, const float* cond, size_t cond_plane,
                    __half* const hi[2], __half* const lo[2], const float* br, int T, int rows, int buf0,
                    int l0, int l1, int num_stages, unsigned int* grid_counter, int num_sms,
                    cudaStream_t stream);

// ---- tcgen05 fused IAF flow, residual stream resident in shared memory (nsw_iaf_flow_tc.cu) ----
int flow_tc_clips_per_launch(int T, int num_sms);  // 0 = shape not supported
// head fused behind the last layer (NULL = publish the final rows for iaf_head_kernel instead)
struct FlowHead {
  int w_tile;  // 64-row tile index of W1^T (hi / lo) in the Wd tensor maps
  const float* wm;
  const float* ws;
  float bm, bs;
  const float* x_in;
  const float* z;
  float* x_out;
  float* mean_tot;
  float* scale_tot;
  float* log_scale_tot;
  int first, last, quantize, use_mu_law;
  float quant_chann;
};
// start conv fused in front of layer 0 (NULL = the rows were written by iaf_start_conv_kernel)
struct FlowStart {
  const float* x;  // input of the flow [B*T]; must not alias FlowHead::x_out
  const float* w;  // start_conv W [3][64]
  const float* b;  // [64]
};
int flow_tc_launch(const void* const map_act[2][2], const void* map_wdh, const void* map_wdl,
                   const void* map_wrh, const void* map_wrl, const float* cond_tiled, size_t cond_plane,
                   const float* br, int T, int clip0, int nclips, int buf0, int l0, int l1, int num_stages,
                   unsigned int* sync_words, int num_sms, const FlowHead* head, const FlowStart* start,
                   cudaStream_t stream);

// ---- the same flow on CTA pairs (nsw_iaf_flow_pair.cu): cta_group::2 MMAs, half of every weight tile per SM, the
//      dilated-conv weights double-buffered.  Weight maps with 32-row boxes.  Whole flow, fused start conv and head only.
int flow_pair_pairs_per_clip(int T, int nclips, int max_pairs);  // 0 = shape not covered
int flow_pair_launch(const void* const map_act[2][2], const void* map_wdh32, const void* map_wdl32, const void* map_wrh32,
                     const void* map_wrl32, const float* cond_tiled, size_t cond_plane, const float* br, int T, int nclips,
                     int buf0, int nl, int num_stages, unsigned int* sync_words, int num_sms, const FlowHead* head,
                     const FlowStart* start, cudaStream_t stream);

// ---- transposed-conv upsampling stack (wavenet._deconv_stack, wavenet.py:46-73) ----
struct DeconvLayer {
  DeconvGeom g;
  DevBuf Bw;              // fp32 [ntaps*cin][s*cout]
  DevBuf Bt_hi, Bt_lo;    // fp16 [s*cout][ntaps*cin_pad] (tensor-core engine; cin_pad = cin rounded up to 64, zero columns)
  int cin_pad = 0;
  DevBuf bias;            // [cout]
};

struct DeconvStack {
  std::vector<DeconvLayer> layers;
  int act = NSW_ACT_TANH;
  int total_stride = 1;
  // prefix: "" (teacher), "iaf_share/", "iaf_3/"
  int init(const TensorMap& tm, const std::string& prefix, int num_mel, int width, int n,
           const int32_t* filt, const int32_t* stride, int act_kind, bool want_tc);
  // mel [B,F,num_mel] -> out [B, F*total_stride, width]; intermediates live in `scratch`.
  // engine FFMA: writes out_f32 (required).  engine TC: the last layer runs on tensor
  // cores when supported and writes out_hi/out_lo (and out_f32 if non-NULL).
  int forward(const float* d_mel, int B, int F, float* out_f32, __half* out_hi,
              __half* out_lo, int engine, DevBuf& scratch, cudaStream_t stream) const;
};

}  // namespace nsw
