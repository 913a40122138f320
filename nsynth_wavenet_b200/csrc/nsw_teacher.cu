// Teacher WaveNet full-sequence forward (Wavenet.feed_forward, wavenet/wavenet.py:180-291) and
// the student->teacher cross-entropy of the distillation loss (loss_func.mol_log_probs
// loss_func.py:22-63 inside ParallelWavenet.kl_loss_logistic parallel_wavenet.py:361-402).
//
// Every contraction of the teacher is a real GEMM (K = 1536 / 256 / 512, N = 512 / 768 / 256), so
// the whole forward is a sequence of tcgen05 conv-GEMM launches (nsw_gemm_tc.cu, split fp16):
//   cond_all  = mel_en . [Wc_1 .. Wc_L | Wc_out1]                 (centre trim = row offset)
//   per layer : g  = gate( dilated3tap(l) + cond_i )               EPI_GATE, tap stride = dilation
//               [l | s] += g . [Wr_i | Ws_i] + [br_i | bs_i]       EPI_ROWS accumulate, one GEMM
//   head      : h = relu(out1 . relu(s) + cond_out1);  out = out2 . h
// Activations live as fp32 master rows [B*T, 768] = [l (512) | s (256)] plus a fp16 hi/lo copy that
// feeds the next GEMM's A operand through TMA.
#include "nsw_gemm.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace nsw {
namespace {

constexpr int TW = 512, TM = 256, TS = 256, TD = 256, TLS = TW + TS;

// l0 = conv_start(shift_right(x)) : l0[t,c] = b[c] + W0[c] x[t-3] + W1[c] x[t-2] + W2[c] x[t-1]
__global__ void __launch_bounds__(256)
teacher_start_kernel(const float* __restrict__ x, const float* __restrict__ w /*[3][512]*/,
                     const float* __restrict__ b, float* __restrict__ ls, __half* __restrict__ hi,
                     __half* __restrict__ lo, int T, size_t rows) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (row, c4)
  const size_t row = idx >> 7;
  if (row >= rows) return;
  const int c = (int)(idx & 127) * 4;
  const int t = (int)(row % T);
  const float x1 = t >= 1 ? x[row - 1] : 0.f, x2 = t >= 2 ? x[row - 2] : 0.f, x3 = t >= 3 ? x[row - 3] : 0.f;
  const float4 w0 = *reinterpret_cast<const float4*>(w + c);
  const float4 w1 = *reinterpret_cast<const float4*>(w + TW + c);
  const float4 w2 = *reinterpret_cast<const float4*>(w + 2 * TW + c);
  const float4 bb = *reinterpret_cast<const float4*>(b + c);
  float f[4];
  f[0] = fmaf(w2.x, x1, fmaf(w1.x, x2, fmaf(w0.x, x3, bb.x)));
  f[1] = fmaf(w2.y, x1, fmaf(w1.y, x2, fmaf(w0.y, x3, bb.y)));
  f[2] = fmaf(w2.z, x1, fmaf(w1.z, x2, fmaf(w0.z, x3, bb.z)));
  f[3] = fmaf(w2.w, x1, fmaf(w1.w, x2, fmaf(w0.w, x3, bb.w)));
  *reinterpret_cast<float4*>(ls + row * TLS + c) = make_float4(f[0], f[1], f[2], f[3]);
  __align__(8) __half h[4], l[4];
  uint32_t rmx = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    range_track(rmx, f[i]);
    h[i] = __float2half_rn(f[i]);
    l[i] = __float2half_rn(f[i] - __half2float(h[i]));
  }
  range_commit(rmx);
  *reinterpret_cast<uint2*>(hi + row * TLS + c) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(lo + row * TLS + c) = *reinterpret_cast<uint2*>(l);
}

// ---- student -> teacher cross entropy (one thread per (b,t), S logistic draws each) ----
// log p_T(x) for a discretised mixture of logistics, exactly as loss_func.mol_log_probs.
__device__ __forceinline__ float softplus_d(float v) { return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v))); }

__global__ void __launch_bounds__(256)
mol_score_kernel(const float* __restrict__ te /*[BT][3*nr]*/, const float* __restrict__ mean,
                 const float* __restrict__ scale, const float* __restrict__ log_scale,
                 const float* __restrict__ eps /*[S][BT] or NULL*/, unsigned long long seed, int S,
                 size_t BT, int nr, float Q, double* __restrict__ acc /*[2]: sum log p, sum log_scale*/) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lp_sum = 0.0, ls_sum = 0.0;
  if (i < BT) {
    float lw[16], mu[16], inv[16];
    const float* p = te + i * (size_t)(3 * nr);
    float mx = -INFINITY;
    for (int k = 0; k < nr; ++k) { lw[k] = p[k]; mx = fmaxf(mx, lw[k]); }
    float se = 0.f;
    for (int k = 0; k < nr; ++k) se += expf(lw[k] - mx);
    const float lse = mx + logf(se);
    for (int k = 0; k < nr; ++k) {
      lw[k] -= lse;                                    // _log_prob_from_logits (loss_func.py:7-11)
      mu[k] = p[nr + k];
      inv[k] = expf(-fmaxf(p[2 * nr + k], -7.0f));    // log_scales = max(., -7) (:31-32)
    }
    const float m = mean[i], s = scale[i];
    const float max_thres = (Q - 1.0f - 0.5f) / (Q * 0.5f) - 1.0f, min_thres = 0.5f / (Q * 0.5f) - 1.0f;
    for (int j = 0; j < S; ++j) {
      float e;
      if (eps) {
        e = eps[(size_t)j * BT + i];
      } else {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)j, 0x6d6f6cu),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float u = u01_clipped(r.x);
        e = logf(u) - logf(1.0f - u);
      }
      const float x = fmaf(e, s, m);                   // x_xp = rl * scale + mean (:373-377)
      float best = -INFINITY, terms[16];
      for (int k = 0; k < nr; ++k) {
        const float cx = x - mu[k];
        const float pin = inv[k] * (cx + 1.0f / Q), nin = inv[k] * (cx - 1.0f / Q);
        float lpk;
        if (x < min_thres) lpk = pin - softplus_d(pin);                     // log cdf_plus
        else if (x > max_thres) lpk = -softplus_d(nin);                     // log(1 - cdf_min)
        else lpk = logf(fmaxf(1.0f / (1.0f + expf(-pin)) - 1.0f / (1.0f + expf(-nin)), 1e-12f));
        terms[k] = lpk + lw[k];
        best = fmaxf(best, terms[k]);
      }
      float sum = 0.f;
      for (int k = 0; k < nr; ++k) sum += expf(terms[k] - best);
      lp_sum += (double)(best + logf(sum));            // _log_sum_exp (:14-19)
    }
    ls_sum = (double)log_scale[i];
  }
  // block reduction -> one atomic per block
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
    ls_sum += __shfl_xor_sync(0xffffffffu, ls_sum, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = lp_sum; red[1][threadIdx.x >> 5] = ls_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(acc, a);
    atomicAdd(acc + 1, b);
  }
}

// ---- closed-form Gaussian KL (ClariNet distillation), HBM-bound: 20 bytes read per sample ----
// kl_loss_gauss (parallel_wavenet.py:404-428) on the teacher's [BT][2] output: per sample
//   log s_p - log s_q + (s_q^2 - s_p^2 + (m_p - m_q)^2) / (2 s_p^2),   reg = (log s_p - log s_q)^2
// with s_p = exp(max(param, -7)) (loss_func.py:66-75).  Grid-stride, fp64 block sums, one atomic pair per block.
__global__ void __launch_bounds__(256)
gauss_kl_kernel(const float2* __restrict__ te /*[BT] (mean, log-scale param)*/, const float* __restrict__ mean,
                const float* __restrict__ scale, const float* __restrict__ log_scale, size_t BT,
                double* __restrict__ acc /*[2]: sum kl, sum reg*/) {
  double kl_sum = 0.0, reg_sum = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < BT; i += (size_t)gridDim.x * blockDim.x) {
    const float2 p = __ldg(te + i);
    const float mq = __ldg(mean + i), sq = __ldg(scale + i), lq = __ldg(log_scale + i);
    const float sp = expf(fmaxf(p.y, -7.0f));
    const float lp = logf(sp);                         // the reference takes tf.log of the std (:418)
    const float vq = sq * sq, vp = sp * sp;
    const float dm = p.x - mq;
    const float kl = lp - lq + (vq - vp + dm * dm) / (2.0f * vp);
    const float dl = lp - lq;
    kl_sum += (double)kl;
    reg_sum += (double)(dl * dl);
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    kl_sum += __shfl_xor_sync(0xffffffffu, kl_sum, o);
    reg_sum += __shfl_xor_sync(0xffffffffu, reg_sum, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = kl_sum; red[1][threadIdx.x >> 5] = reg_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(acc, a);
    atomicAdd(acc + 1, b);
  }
}

}  // namespace
}  // namespace nsw

using namespace nsw;

struct TeacherLayer {
  DevBuf wd_hi, wd_lo;   // [512 pos][1536], or [512 pos][1536 + 256] with the mel-cond projection appended along K
  DevBuf bg;             // [512 pos] dilated-conv + mel-cond biases (fused conditioning only)
  DevBuf wrs_hi, wrs_lo; // [768][256]
  DevBuf brs;            // [768]
};

struct nsw_teacher {
  nsw_wavenet_config cfg;
  int device = 0, L = 0, O = 0, NC = 0;
  // the layers' mel conditioning rides inside the dilated-conv GEMM as 256 extra K columns read straight from mel_en
  // (no [rows x 30*512] fp32 conditioning plane: 3.4 GB written and read back at the distillation shape, 1.6 of the
  // 14.7 ms, and a latency-bound addend read in every gate epilogue); NSW_TEACHER_COND_SEPARATE=1 keeps the old scheme
  bool fuse_cond = true;
  // residual + skip update: l|s live ONLY as the split pair (22 mantissa bits; no fp32 master to read, add to and
  // write back) and the update is one GEMM whose accumulate source goes through the tensor core (ConvGemm::acc3);
  // NSW_TEACHER_FP32_MASTER=1 keeps the fp32 rows and the read-modify-write epilogue
  bool acc_mma = true;
  // ConvGemm::split_acc for the dilated-conv GEMMs.  Max-abs error of out_params against the fp64 reference of tests/test_teacher_gpu.py / forward time at
  // 7 x 7680 (profiles/r02, run33), conditioning fused:  split 5.6e-5 / 11.45 ms;  no split 1.01e-4 / 10.05 ms (over the
  // 1e-4 bar: 17 % more instructions into the large accumulator than the separate-conditioning scheme, 8.3e-5 / 11.5 ms).
  // On unless NSW_TEACHER_SPLIT_ACC=0.
  bool split_acc = true;
  DeconvStack deconv;
  std::vector<TeacherLayer> layers;
  DevBuf wc_hi, wc_lo, bc;          // cond_all: [NC][256], bias [NC]
  DevBuf wcs, bcs;                  // conv_start
  DevBuf wss_hi, wss_lo, bss;       // skip_start [256][512]
  DevBuf wo1_hi, wo1_lo, zeros;     // out1 [256][256]
  DevBuf wo2_hi, wo2_lo, bo2;       // out2 [64][256] (rows >= O are zero)
  // workspace
  DevBuf mel, wav, mel_en, cond_all, ls, ls_split, g_split, h_split, out_pad, scratch, acc, out_dev;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace nsw_teacher_host {
static inline int gate_ch(int pos) { return (pos & 1) ? (pos >> 1) + TM : (pos >> 1); }

static int upload_split(DevBuf& hi, DevBuf& lo, const std::vector<float>& w) {
  std::vector<__half> h(w.size()), l(w.size());
  split_f16(w.data(), w.size(), h.data(), l.data());
  NSW_TRY(upload(hi, h.data(), h.size() * 2));
  NSW_TRY(upload(lo, l.data(), l.size() * 2));
  return NSW_OK;
}
}  // namespace nsw_teacher_host
using namespace nsw_teacher_host;

extern "C" void nsw_teacher_destroy(nsw_teacher* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

extern "C" int nsw_teacher_create(const nsw_wavenet_config* cfg, const nsw_tensor* tensors, int32_t n,
                                  int32_t device, nsw_teacher** out) {
  NSW_CHECK(cfg && tensors && out, NSW_EINVAL, "nsw_teacher_create: null argument");
  NSW_CHECK(cfg->width == TW && cfg->gate_width == 2 * TM && cfg->skip_width == TS &&
                cfg->deconv_width == TD && cfg->filter_length == 3,
            NSW_EINVAL, "teacher forward is specialised for width=512, gate=512, skip=256, deconv=256, k=3");
  NSW_CHECK(cfg->out_width >= 1 && cfg->out_width <= 64, NSW_EINVAL, "teacher out_width %d > 64 (ce head is a later row)",
            cfg->out_width);
  NSW_CHECK(cfg->num_layers >= 1 && cfg->num_layers <= 64, NSW_EINVAL, "bad num_layers");
  NSW_CUDA(cudaSetDevice(device));
  NSW_TRY(range_guard_init(device));
  TensorMap tm(tensors, n);
  nsw_teacher* h = new nsw_teacher();
  h->cfg = *cfg;
  h->device = device;
  h->fuse_cond = !(getenv("NSW_TEACHER_COND_SEPARATE") && atoi(getenv("NSW_TEACHER_COND_SEPARATE")) != 0) &&
                 getenv("NSW_GEMM_1CTA") == nullptr;
  h->acc_mma = !(getenv("NSW_TEACHER_FP32_MASTER") && atoi(getenv("NSW_TEACHER_FP32_MASTER")) != 0) &&
               getenv("NSW_GEMM_1CTA") == nullptr;
  h->split_acc = !(getenv("NSW_TEACHER_SPLIT_ACC") && atoi(getenv("NSW_TEACHER_SPLIT_ACC")) == 0) &&
                 getenv("NSW_GEMM_1CTA") == nullptr;
  const int L = h->L = cfg->num_layers, O = h->O = cfg->out_width;
  const int G = 2 * TM;
  h->NC = L * G + TS;
  int rc = h->deconv.init(tm, "", cfg->num_mel, TD, cfg->num_deconv, cfg->deconv_filter,
                          cfg->deconv_stride, cfg->upsample_act, true);
  auto fail = [&](int code) { nsw_teacher_destroy(h); return code; };
  if (rc != NSW_OK) return fail(rc);
  std::vector<float> wc((size_t)h->NC * TD), bc(h->NC);
  h->layers.resize(L);
  for (int i = 0; i < L; ++i) {
    const std::string li = std::to_string(i + 1);
    const float* wd = tm.get("dilated_conv_" + li + "/W", 3 * TW * G);
    const float* bd = tm.get("dilated_conv_" + li + "/biases", G);
    const float* wcd = tm.get("mel_cond_" + li + "/W", TD * G);
    const float* bcd = tm.get("mel_cond_" + li + "/biases", G);
    const float* wr = tm.get("res_" + li + "/W", TM * TW);
    const float* br = tm.get("res_" + li + "/biases", TW);
    const float* ws = tm.get("skip_" + li + "/W", TM * TS);
    const float* bs = tm.get("skip_" + li + "/biases", TS);
    if (!wd || !bd || !wcd || !bcd || !wr || !br || !ws || !bs) return fail(NSW_EMISSING);
    const int K0 = h->fuse_cond ? TD : 0;  // fused conditioning: its 256 K columns come first
    const int KD = K0 + 3 * TW;
    std::vector<float> wdt((size_t)G * KD), wrs((size_t)TLS * TM), brs(TLS), bg(G);
    for (int pos = 0; pos < G; ++pos) {
      const int ch = gate_ch(pos);
      for (int tap = 0; tap < 3; ++tap)
        for (int c = 0; c < TW; ++c) wdt[(size_t)pos * KD + K0 + tap * TW + c] = wd[((size_t)tap * TW + c) * G + ch];
      for (int k = 0; k < TD; ++k) {
        wc[((size_t)i * G + pos) * TD + k] = wcd[(size_t)k * G + ch];
        if (h->fuse_cond) wdt[(size_t)pos * KD + k] = wcd[(size_t)k * G + ch];
      }
      bc[i * G + pos] = bg[pos] = bd[ch] + bcd[ch];
    }
    for (int c = 0; c < TW; ++c) {
      for (int j = 0; j < TM; ++j) wrs[(size_t)c * TM + j] = wr[(size_t)j * TW + c];
      brs[c] = br[c];
    }
    for (int c = 0; c < TS; ++c) {
      for (int j = 0; j < TM; ++j) wrs[(size_t)(TW + c) * TM + j] = ws[(size_t)j * TS + c];
      brs[TW + c] = bs[c];
    }
    TeacherLayer& ly = h->layers[i];
    if ((rc = upload_split(ly.wd_hi, ly.wd_lo, wdt)) != NSW_OK) return fail(rc);
    if ((rc = upload(ly.bg, bg.data(), bg.size() * 4)) != NSW_OK) return fail(rc);
    if ((rc = upload_split(ly.wrs_hi, ly.wrs_lo, wrs)) != NSW_OK) return fail(rc);
    if ((rc = upload(ly.brs, brs.data(), brs.size() * 4)) != NSW_OK) return fail(rc);
  }
  const float* wcs = tm.get("conv_start/W", 3 * TW);
  const float* bcs = tm.get("conv_start/biases", TW);
  const float* wss = tm.get("skip_start/W", TW * TS);
  const float* bss = tm.get("skip_start/biases", TS);
  const float* wo1 = tm.get("out1/W", TS * TS);
  const float* bo1 = tm.get("out1/biases", TS);
  const float* wco = tm.get("mel_cond_out1/W", TD * TS);
  const float* bco = tm.get("mel_cond_out1/biases", TS);
  const float* wo2 = tm.get("out2/W", TS * O);
  const float* bo2 = tm.get("out2/biases", O);
  if (!wcs || !bcs || !wss || !bss || !wo1 || !bo1 || !wco || !bco || !wo2 || !bo2) return fail(NSW_EMISSING);
  for (int c = 0; c < TS; ++c) {
    for (int k = 0; k < TD; ++k) wc[((size_t)L * G + c) * TD + k] = wco[(size_t)k * TS + c];
    bc[L * G + c] = bo1[c] + bco[c];
  }
  std::vector<float> wsst((size_t)TS * TW), wo1t((size_t)TS * TS), wo2t((size_t)64 * TS, 0.f), bo2p(64, 0.f),
      zeros(TLS, 0.f);
  for (int c = 0; c < TS; ++c)
    for (int k = 0; k < TW; ++k) wsst[(size_t)c * TW + k] = wss[(size_t)k * TS + c];
  for (int c = 0; c < TS; ++c)
    for (int k = 0; k < TS; ++k) wo1t[(size_t)c * TS + k] = wo1[(size_t)k * TS + c];
  for (int o = 0; o < O; ++o) {
    for (int k = 0; k < TS; ++k) wo2t[(size_t)o * TS + k] = wo2[(size_t)k * O + o];
    bo2p[o] = bo2[o];
  }
  rc = upload_split(h->wc_hi, h->wc_lo, wc);
  if (rc == NSW_OK) rc = upload(h->bc, bc.data(), bc.size() * 4);
  if (rc == NSW_OK) rc = upload(h->wcs, wcs, 3 * TW * 4);
  if (rc == NSW_OK) rc = upload(h->bcs, bcs, TW * 4);
  if (rc == NSW_OK) rc = upload_split(h->wss_hi, h->wss_lo, wsst);
  if (rc == NSW_OK) rc = upload(h->bss, bss, TS * 4);
  if (rc == NSW_OK) rc = upload_split(h->wo1_hi, h->wo1_lo, wo1t);
  if (rc == NSW_OK) rc = upload(h->zeros, zeros.data(), zeros.size() * 4);
  if (rc == NSW_OK) rc = upload_split(h->wo2_hi, h->wo2_lo, wo2t);
  if (rc == NSW_OK) rc = upload(h->bo2, bo2p.data(), bo2p.size() * 4);
  if (rc == NSW_OK) rc = h->acc.ensure(2 * sizeof(double));
  if (rc == NSW_OK) {
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) { set_error("nsw_teacher_create: %s", cudaGetErrorString(e)); rc = NSW_ECUDA; }
  }
  if (rc != NSW_OK) return fail(rc);
  *out = h;
  return NSW_OK;
}

extern "C" int nsw_teacher_forward_device(nsw_teacher* h, const float* d_wav, const float* d_mel, int32_t B,
                                          int32_t T, int32_t F, float* d_out, void* stream) {
  NSW_CHECK(h && d_wav && d_mel && d_out, NSW_EINVAL, "null argument");
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Lc = F * h->deconv.total_stride;
  NSW_CHECK(B >= 1 && T >= 128 && T % 128 == 0, NSW_EINVAL, "teacher forward needs T %% 128 == 0 (got %d)", T);
  NSW_CHECK(Lc >= T, NSW_EINVAL, "conditioning (%d) shorter than the waveform (%d)", Lc, T);
  const int left = (Lc - T) / 2;  // wavenet._condition (wavenet.py:76-85)
  const size_t rows = (size_t)B * T;
  const int L = h->L, NC = h->NC, G = 2 * TM;
  NSW_TRY(h->mel_en.ensure((size_t)B * Lc * TD * 2 * sizeof(__half)));
  const bool fuse = h->fuse_cond;
  const int NCb = fuse ? TS : NC;  // columns of the conditioning plane that is still materialised (out1's only)
  NSW_TRY(h->cond_all.ensure(rows * NCb * sizeof(float)));
  NSW_TRY(h->ls.ensure(rows * TLS * sizeof(float)));
  NSW_TRY(h->ls_split.ensure(rows * TLS * 2 * sizeof(__half)));
  NSW_TRY(h->g_split.ensure(rows * TM * 2 * sizeof(__half)));
  NSW_TRY(h->h_split.ensure(rows * TS * 2 * sizeof(__half)));
  NSW_TRY(h->out_pad.ensure(rows * 64 * sizeof(float)));
  __half* me_hi = h->mel_en.as<__half>();
  __half* me_lo = me_hi + (size_t)B * Lc * TD;
  __half* ls_hi = h->ls_split.as<__half>();
  __half* ls_lo = ls_hi + rows * TLS;
  __half* g_hi = h->g_split.as<__half>();
  __half* g_lo = g_hi + rows * TM;
  __half* h_hi = h->h_split.as<__half>();
  __half* h_lo = h_hi + rows * TS;
  float* ls = h->ls.as<float>();
  float* cond = h->cond_all.as<float>();
  NSW_CUDA(cudaEventRecord(h->ev0, st));
  NSW_TRY(h->deconv.forward(d_mel, B, F, nullptr, me_hi, me_lo, NSW_ENGINE_TC, h->scratch, st));
  {  // conditioning projections hoisted out of the layers (all of them, or only out1's); the centre trim is the row offset
    const size_t w0 = fuse ? (size_t)L * G * TD : 0;
    ConvGemm g; g.nclips = B; g.L = Lc; g.cin = TD; g.ntaps = 1; g.a_off = left; g.mclip = T; g.N = NCb;
    EpiParams e{}; e.mode = EPI_ROWS; e.bias = h->bc.as<float>() + (fuse ? L * G : 0); e.out_f32 = cond; e.ld_out = NCb;
    NSW_TRY(conv_gemm_tc(g, me_hi, me_lo, h->wc_hi.as<__half>() + w0, h->wc_lo.as<__half>() + w0, e, st));
  }
  teacher_start_kernel<<<(unsigned)((rows * 128 + 255) / 256), 256, 0, st>>>(
      d_wav, h->wcs.as<float>(), h->bcs.as<float>(), ls, ls_hi, ls_lo, T, rows);
  count_launch();
  {  // skip_start: s = Wss . l0 + b   (wavenet.py:233-235)
    ConvGemm g; g.nclips = B; g.L = T; g.cin = TW; g.x_pitch = TLS; g.ntaps = 1; g.a_off = 0; g.mclip = T; g.N = TS;
    EpiParams e{}; e.mode = EPI_ROWS; e.bias = h->bss.as<float>(); e.out_f32 = ls + TW; e.ld_out = TLS;
    e.out_hi = ls_hi + TW; e.out_lo = ls_lo + TW; e.ld_split = TLS;
    NSW_TRY(conv_gemm_tc(g, ls_hi, ls_lo, h->wss_hi.as<__half>(), h->wss_lo.as<__half>(), e, st));
  }
  for (int i = 0; i < L; ++i) {
    const int d = 1 << (i % h->cfg.num_stages);
    TeacherLayer& ly = h->layers[i];
    {  // dilated causal conv + cond + gate  (wavenet.py:244-267)
      ConvGemm g; g.nclips = B; g.L = T; g.cin = TW; g.x_pitch = TLS; g.ntaps = 3; g.tap_stride = d;
      g.a_off = -2 * d; g.mclip = T; g.N = G;
      g.split_acc = h->split_acc ? 1 : 0;
      EpiParams e{}; e.mode = EPI_GATE;
      e.out_hi = g_hi; e.out_lo = g_lo; e.ld_split = TM;
      if (fuse) {
        g.cin2 = TD; g.a_off2 = left; g.L2 = Lc;
        e.bias = ly.bg.as<float>();
        NSW_TRY(conv_gemm_tc(g, ls_hi, ls_lo, ly.wd_hi.as<__half>(), ly.wd_lo.as<__half>(), e, st, me_hi, me_lo));
      } else {
        e.addend = cond + (size_t)i * G; e.ld_add = NC;
        NSW_TRY(conv_gemm_tc(g, ls_hi, ls_lo, ly.wd_hi.as<__half>(), ly.wd_lo.as<__half>(), e, st));
      }
    }
    {  // l += res(g), s += skip(g) in one GEMM  (wavenet.py:269-274)
      ConvGemm g; g.nclips = B; g.L = T; g.cin = TM; g.ntaps = 1; g.a_off = 0; g.mclip = T; g.N = TLS;
      EpiParams e{}; e.mode = EPI_ROWS; e.bias = ly.brs.as<float>();
      e.out_hi = ls_hi; e.out_lo = ls_lo; e.ld_split = TLS;
      if (i == L - 1) e.relu_split_from = TW;  // the head consumes relu(s) (wavenet.py:281)
      if (h->acc_mma) {
        g.acc3 = 1; g.ld3 = TLS;
        NSW_TRY(conv_gemm_tc(g, g_hi, g_lo, ly.wrs_hi.as<__half>(), ly.wrs_lo.as<__half>(), e, st, nullptr, nullptr, ls_hi,
                             ls_lo));
      } else {
        e.out_f32 = ls; e.ld_out = TLS; e.accumulate = 1;
        NSW_TRY(conv_gemm_tc(g, g_hi, g_lo, ly.wrs_hi.as<__half>(), ly.wrs_lo.as<__half>(), e, st));
      }
    }
  }
  {  // h = relu(out1 . relu(s) + cond_out1)  (wavenet.py:281-286)
    ConvGemm g; g.nclips = B; g.L = T; g.cin = TS; g.x_pitch = TLS; g.ntaps = 1; g.a_off = 0; g.mclip = T; g.N = TS;
    EpiParams e{}; e.mode = EPI_ROWS; e.bias = h->zeros.as<float>(); e.addend = cond + (fuse ? 0 : (size_t)L * G); e.ld_add = NCb;
    e.relu_out = 1; e.out_hi = h_hi; e.out_lo = h_lo; e.ld_split = TS;
    NSW_TRY(conv_gemm_tc(g, ls_hi + TW, ls_lo + TW, h->wo1_hi.as<__half>(), h->wo1_lo.as<__half>(), e, st));
  }
  {  // out = out2 . h  (wavenet.py:287-288), padded to 64 columns
    ConvGemm g; g.nclips = B; g.L = T; g.cin = TS; g.ntaps = 1; g.a_off = 0; g.mclip = T; g.N = 64;
    EpiParams e{}; e.mode = EPI_ROWS; e.bias = h->bo2.as<float>(); e.out_f32 = h->out_pad.as<float>(); e.ld_out = 64;
    NSW_TRY(conv_gemm_tc(g, h_hi, h_lo, h->wo2_hi.as<__half>(), h->wo2_lo.as<__half>(), e, st));
  }
  NSW_CUDA(cudaMemcpy2DAsync(d_out, (size_t)h->O * 4, h->out_pad.p, 64 * 4, (size_t)h->O * 4, rows,
                             cudaMemcpyDeviceToDevice, st));
  NSW_CUDA(cudaEventRecord(h->ev1, st));
  return NSW_OK;
}

extern "C" int nsw_teacher_forward_host(nsw_teacher* h, const float* wav, const float* mel, int32_t B, int32_t T,
                                        int32_t F, float* out) {
  NSW_CHECK(h && wav && mel && out, NSW_EINVAL, "null argument");
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const size_t nm = (size_t)B * F * h->cfg.num_mel * 4, nw = (size_t)B * T * 4, no = (size_t)B * T * h->O * 4;
  NSW_TRY(h->mel.ensure(nm));
  NSW_TRY(h->wav.ensure(nw));
  NSW_TRY(h->out_dev.ensure(no));
  NSW_CUDA(cudaMemcpyAsync(h->mel.p, mel, nm, cudaMemcpyHostToDevice, st));
  NSW_CUDA(cudaMemcpyAsync(h->wav.p, wav, nw, cudaMemcpyHostToDevice, st));
  NSW_TRY(nsw_teacher_forward_device(h, h->wav.as<float>(), h->mel.as<float>(), B, T, F, h->out_dev.as<float>(), st));
  NSW_CUDA(cudaMemcpyAsync(out, h->out_dev.p, no, cudaMemcpyDeviceToHost, st));
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_error("teacher forward failed: %s", cudaGetErrorString(e)); return NSW_ECUDA; }
  return range_check("nsw_teacher_forward_host");
}

extern "C" int nsw_teacher_last_timing(nsw_teacher* h, float* ms) {
  NSW_CHECK(h && ms, NSW_EINVAL, "null argument");
  NSW_CUDA(cudaEventSynchronize(h->ev1));
  NSW_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return NSW_OK;
}

// result[0] = H_Ps, result[1] = H_Ps_Pt, result[2] = kl_loss  (parallel_wavenet.py:392-398)
extern "C" int nsw_mol_score_device(nsw_teacher* h, const float* d_te_out, const float* d_mean, const float* d_scale,
                                    const float* d_log_scale, const float* d_eps, uint64_t seed, int32_t S, int32_t B,
                                    int32_t T, double* result, void* stream) {
  NSW_CHECK(h && d_te_out && d_mean && d_scale && d_log_scale && result, NSW_EINVAL, "null argument");
  NSW_CHECK(h->cfg.loss_type == NSW_LOSS_MOL && h->O % 3 == 0 && h->O / 3 <= 16, NSW_EINVAL,
            "mol scoring needs a mol teacher with <= 16 mixtures");
  NSW_CHECK(S >= 1, NSW_EINVAL, "num_samples must be >= 1");
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BT = (size_t)B * T;
  NSW_CUDA(cudaMemsetAsync(h->acc.p, 0, 2 * sizeof(double), st));
  mol_score_kernel<<<(unsigned)((BT + 255) / 256), 256, 0, st>>>(d_te_out, d_mean, d_scale, d_log_scale, d_eps, seed, S,
                                                                BT, h->O / 3, h->cfg.use_mu_law ? 256.0f : 65536.0f,
                                                                h->acc.as<double>());
  count_launch();
  double host[2];
  NSW_CUDA(cudaMemcpyAsync(host, h->acc.p, sizeof(host), cudaMemcpyDeviceToHost, st));
  NSW_CUDA(cudaStreamSynchronize(st));
  const double H_Ps_Pt = -host[0] / ((double)BT * S);
  const double H_Ps = host[1] / (double)BT + 2.0;
  result[0] = H_Ps;
  result[1] = H_Ps_Pt;
  result[2] = H_Ps_Pt - H_Ps;
  return NSW_OK;
}

// result[0] = mean KL term, result[1] = mean squared log-scale difference, result[2] = kl_loss = [0] + 4 [1]
// (parallel_wavenet.py:422-428)
extern "C" int nsw_gauss_kl_device(nsw_teacher* h, const float* d_te_out, const float* d_mean, const float* d_scale,
                                   const float* d_log_scale, int32_t B, int32_t T, double* result, void* stream) {
  NSW_CHECK(h && d_te_out && d_mean && d_scale && d_log_scale && result, NSW_EINVAL, "null argument");
  NSW_CHECK(h->cfg.loss_type == NSW_LOSS_GAUSS && h->O == 2, NSW_EINVAL,
            "gaussian KL needs a gauss teacher (out_width 2, got loss_type %d out_width %d)", h->cfg.loss_type, h->O);
  NSW_CHECK(B >= 1 && T >= 1, NSW_EINVAL, "bad batch/length %d/%d", B, T);
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BT = (size_t)B * T;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  // 8 resident CTAs of 256 threads per SM cover the latency of the 4 independent streams; never more CTAs than work
  const unsigned grid = (unsigned)std::min<size_t>((size_t)sms * 8, (BT + 255) / 256);
  NSW_CUDA(cudaMemsetAsync(h->acc.p, 0, 2 * sizeof(double), st));
  gauss_kl_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(d_te_out), d_mean, d_scale, d_log_scale, BT,
                                        h->acc.as<double>());
  count_launch();
  double host[2];
  NSW_CUDA(cudaMemcpyAsync(host, h->acc.p, sizeof(host), cudaMemcpyDeviceToHost, st));
  NSW_CUDA(cudaStreamSynchronize(st));
  result[0] = host[0] / (double)BT;
  result[1] = host[1] / (double)BT;
  result[2] = result[0] + 4.0 * result[1];
  return NSW_OK;
}

NSW_RANGE_GUARD_TU(teacher)
