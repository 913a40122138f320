/*
 * nsw.h — C ABI of libnsw_b200.so, the B200 (sm_100a) implementation of the
 * bfs18/nsynth_wavenet generation hot path.
 *
 * The reference has no FFI layer: its boundary is the Python module API used by
 * eval_parallel_wavenet.py / eval_wavenet.py (wavenet/parallelgen.py:11-51,
 * wavenet/fastgen.py:61-169).  Each entry point below names the reference
 * function whose device work it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - plain C types only; tensors are row-major fp32, activations [batch,time,chan]
 *    exactly as in the reference (wavenet/masked.py:43).
 *  - every call returns 0 on success, a negative NSW_E* code otherwise;
 *    nsw_last_error() returns a thread-local message for the last failure.
 *  - "_device" entry points take device pointers owned by the caller and enqueue
 *    on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *    without synchronising.  "_host" entry points take host pointers, do the
 *    H2D / D2H copies themselves and return after the result is in host memory.
 *  - handles are not thread-safe; distinct handles may be used concurrently.
 *  - weights are handed over as a list of named tensors using the reference's
 *    TensorFlow variable names (the checkpoint contract), e.g.
 *    "iaf_1/dilated_conv_3/W" [1,3,64,64], "iaf_share/trans_conv_2/kernel"
 *    [1,80,256,256].  The library repacks them once at create time.
 */
#ifndef NSW_H_
#define NSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSW_OK 0
#define NSW_EINVAL (-1)   /* bad argument / unsupported configuration        */
#define NSW_ECUDA (-2)    /* CUDA runtime or driver error                    */
#define NSW_EMISSING (-3) /* a required weight tensor is missing / misshaped */
#define NSW_ETIMEOUT (-4) /* persistent kernel watchdog fired                */
#define NSW_ERANGE (-5)   /* an activation left the fp16 range of the split-precision tensor-core engines */

#define NSW_MAX_FLOWS 8
#define NSW_MAX_DECONV 4

/* loss_type: which noise the student consumes / which head the teacher has */
#define NSW_LOSS_LOGISTIC 0 /* student 'logistic' (parallel_wavenet.py:303-304) */
#define NSW_LOSS_GAUSS 1    /* 'gauss'                                          */
#define NSW_LOSS_MOL 2      /* teacher 'mol' (wavenet.py:123-125)               */
#define NSW_LOSS_CE 3       /* teacher 'ce'                                     */

/* upsample_act (wavenet/masked.py:28-36) */
#define NSW_ACT_TANH 0
#define NSW_ACT_RELU 1
#define NSW_ACT_LEAKY_RELU 2 /* alpha = 0.4 */

/* engine selection for the dense contractions (cond projection, deconv) */
#define NSW_ENGINE_FFMA 0 /* fp32 CUDA-core path, bit-for-bit fp32 products   */
#define NSW_ENGINE_TC 1   /* tcgen05 split-fp16 (hi*hi + hi*lo + lo*hi)       */
#define NSW_ENGINE_TC2 2  /* NSW_ENGINE_TC + the IAF residual layers on tcgen05  */
#define NSW_ENGINE_TC3 3  /* NSW_ENGINE_TC2 with the residual stream resident in shared memory for a
                             whole flow (one persistent launch per flow, no grid barrier); shapes it
                             cannot take (a clip longer than 4 * 128 * #SMs samples) run as TC2 */

typedef struct nsw_tensor {
  const char* name;    /* TF variable name, without ":0" and without EMA suffix */
  const float* data;   /* host pointer, row-major                               */
  int32_t ndim;
  int64_t shape[4];
} nsw_tensor;

/* hparams of ParallelWavenet (config_jsons/parallel_wavenet*.json;
 * parallel_wavenet.py:118-147) */
typedef struct nsw_iaf_config {
  int32_t num_flows;                       /* len(num_iaf_layers)            */
  int32_t num_iaf_layers[NSW_MAX_FLOWS];
  int32_t num_stages;                      /* dilation = 2^(i % num_stages)  */
  int32_t filter_length;                   /* must be 3                      */
  int32_t width;                           /* must be 64 (gate = width)      */
  int32_t deconv_width;                    /* must be 256                    */
  int32_t num_mel;                         /* 80                             */
  int32_t num_deconv;                      /* len(deconv_config)             */
  int32_t deconv_filter[NSW_MAX_DECONV];
  int32_t deconv_stride[NSW_MAX_DECONV];
  int32_t share_deconv;                    /* use_share_deconv || use_teacher_deconv */
  int32_t loss_type;                       /* NSW_LOSS_LOGISTIC | NSW_LOSS_GAUSS */
  int32_t upsample_act;
  int32_t use_mu_law;
  int32_t engine;                          /* NSW_ENGINE_*                   */
} nsw_iaf_config;

/* hparams of Wavenet / Fastgen (config_jsons/wavenet_*.json; wavenet.py:97-128,332-351) */
typedef struct nsw_wavenet_config {
  int32_t num_layers;
  int32_t num_stages;
  int32_t filter_length; /* must be 3 (masked.py:349) */
  int32_t width;
  int32_t gate_width;    /* 2*width if double_gate_width else width */
  int32_t skip_width;
  int32_t out_width;     /* 3*mol_mix | 2 | quant_chann */
  int32_t deconv_width;
  int32_t num_mel;
  int32_t num_deconv;
  int32_t deconv_filter[NSW_MAX_DECONV];
  int32_t deconv_stride[NSW_MAX_DECONV];
  int32_t loss_type;     /* NSW_LOSS_MOL | NSW_LOSS_GAUSS | NSW_LOSS_CE */
  int32_t upsample_act;
  int32_t use_mu_law;
  int32_t engine;
} nsw_wavenet_config;

typedef struct nsw_iaf nsw_iaf;
typedef struct nsw_fastgen nsw_fastgen;

/* ---- library ------------------------------------------------------------ */
int nsw_version(void);
const char* nsw_last_error(void);
/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches claim is read from here, not estimated) */
uint64_t nsw_kernel_launch_count(void);

/* CRC-32C (Castagnoli) of `n` bytes continuing from `crc` (0 to start): the checksum of TensorFlow's V2 checkpoint
 * bundles (index blocks and tensors), used by the TensorFlow-free bundle reader that replaces Saver.restore
 * (fastgen.py:81-84, parallelgen.py:40-41).  Host only, little-endian byte order. */
uint32_t nsw_crc32c(const void* data, size_t n, uint32_t crc);

/* The tensor-core engines carry activations as fp16 hi + fp16 lo (fp32-grade, 22 mantissa bits) and therefore share
 * fp16's range: |v| <= 65504.  Every conversion site checks it; the "_host" entry points return NSW_ERANGE instead of
 * a silently wrong result, and after "_device" calls this function (which synchronises `device`) reports and clears
 * the condition: NSW_OK or NSW_ERANGE. */
int nsw_range_status(int32_t device);

/* ---- parallel IAF student ------------------------------------------------
 * replaces: parallelgen.load_parallelgen (parallelgen.py:11-19) = graph build of
 * ParallelWavenet.feed_forward (parallel_wavenet.py:289-345) + _clip_quant_scale
 * (:348-359), and the Saver.restore of parallelgen.py:30-41 (weights arrive as
 * named tensors instead of a checkpoint path). */
int nsw_iaf_create(const nsw_iaf_config* cfg, const nsw_tensor* tensors, int32_t n_tensors,
                   int32_t device, nsw_iaf** out);
void nsw_iaf_destroy(nsw_iaf* h);

/* T = (F * prod(strides) / 2^(num_stages-1)) * 2^(num_stages-1)  (parallel_wavenet.py:302) */
int64_t nsw_iaf_length(const nsw_iaf* h, int32_t num_frames);

/* replaces: sess.run(fg_dict[...], {mel_in: mel}) (parallelgen.py:44).
 * d_mel [B,F,num_mel]; d_z [B,T] noise or NULL (then drawn on device from `seed`
 * with Philox: logistic log u - log(1-u), u~U[1e-5,1-1e-5], or N(0,1));
 * outputs [B,T] each, any of them may be NULL:
 *   d_x            feed_forward 'x' (after _clip_quant_scale iff quantize != 0)
 *   d_mean_tot, d_scale_tot, d_log_scale_tot, d_rand_input  feed_forward dict  */
int nsw_iaf_forward_device(nsw_iaf* h, const float* d_mel, const float* d_z, uint64_t seed,
                           int32_t B, int32_t F, int32_t quantize, float* d_x, float* d_mean_tot,
                           float* d_scale_tot, float* d_log_scale_tot, float* d_rand_input,
                           void* stream);

/* same with host buffers (pinned or pageable); H2D of mel (+z) and D2H of the
 * requested outputs happen inside the call.  This is what parallelgen.synthesis
 * calls. */
int nsw_iaf_forward_host(nsw_iaf* h, const float* mel, const float* z, uint64_t seed, int32_t B,
                         int32_t F, int32_t quantize, float* x, float* mean_tot, float* scale_tot,
                         float* log_scale_tot, float* rand_input);

/* kernel-level parity hooks -------------------------------------------------
 * replaces: wavenet._deconv_stack (wavenet.py:46-73).  stack = 0 for the shared
 * stack, else flow index.  d_mel_en [B, F*prod(strides), deconv_width] fp32. */
int nsw_iaf_deconv_device(nsw_iaf* h, int32_t stack, const float* d_mel, int32_t B, int32_t F,
                          float* d_mel_en, void* stream);
/* after the next nsw_iaf_forward_*, copy the residual stream l [B,T,width] as it
 * is after `layer` residual layers of `flow` (layer 0 = start_conv output) into
 * d_l.  Pass d_l = NULL to clear. */
int nsw_iaf_set_tap(nsw_iaf* h, int32_t flow, int32_t layer, float* d_l);
/* bytes of device workspace currently held (grows with B*F) */
size_t nsw_iaf_workspace_bytes(const nsw_iaf* h);
/* per-stage device time of the last forward in ms (CUDA events on the launch
 * stream): [0]=deconv, [1]=cond projection, [2]=residual layer kernels only,
 * [3]=start conv + heads/affine, [4]=total.  Valid only after nsw_iaf_set_profiling(h,1). */
int nsw_iaf_set_profiling(nsw_iaf* h, int32_t on);
int nsw_iaf_last_timing(nsw_iaf* h, float ms[5]);

/* ---- autoregressive teacher (fastgen) ------------------------------------
 * replaces: fastgen.load_deconv_stack/encode (fastgen.py:61-88),
 * load_fastgen + synthesis loop (fastgen.py:118-169) = Fastgen.sample
 * (wavenet.py:379-514) with masked.causal_linear / linear (masked.py:328-405). */
int nsw_fastgen_create(const nsw_wavenet_config* cfg, const nsw_tensor* tensors,
                       int32_t n_tensors, int32_t device, nsw_fastgen** out);
void nsw_fastgen_destroy(nsw_fastgen* h);

/* mel [B,F,num_mel] -> encoding [B, F*prod(strides), deconv_width] (fastgen.encode) */
int nsw_fastgen_encode_device(nsw_fastgen* h, const float* d_mel, int32_t B, int32_t F,
                              float* d_encoding, void* stream);
int nsw_fastgen_encode_host(nsw_fastgen* h, const float* mel, int32_t B, int32_t F,
                            float* encoding);

/* the whole per-sample loop of fastgen.synthesis as ONE persistent kernel.  Two engines sit behind this call: the
 * latency engine (batch rows one after the other, gate_width = width, mol / gauss heads) and the batched engine
 * (up to 8 rows per weight pass; also double_gate_width, the ce head and mu-law input).  B >= 3 or a configuration the
 * latency engine does not cover selects the batched one; NSW_FASTGEN_ENGINE=gn|latency overrides.
 * d_encoding [B,T,deconv_width] (fed step by step with NO centre trim,
 * fastgen.py:157); d_teacher_force [B,T] or NULL: if given the wav fed at step i
 * is teacher_force[:, i-1] (0 at i = 0) instead of the model's own sample;
 * outputs: d_audio [B,T] fp32 dequantised samples (may be NULL),
 *          d_out [B,T,out_width] pre-sample parameters (may be NULL). */
int nsw_fastgen_run_device(nsw_fastgen* h, const float* d_encoding, int32_t B, int32_t T,
                           const float* d_teacher_force, uint64_t seed, float* d_audio,
                           float* d_out, void* stream);
int nsw_fastgen_run_host(nsw_fastgen* h, const float* encoding, int32_t B, int32_t T,
                         const float* teacher_force, uint64_t seed, float* audio, float* out);
/* replaces: fastgen.load_cond_layers / calculate_cond_vars (fastgen.py:91-115) = Fastgen.cond_vars
 * (wavenet.py:353-377): encoding [B,T,deconv_width] -> out [B,T, num_layers*gate_width + skip_width], columns
 * [i*gate_width, (i+1)*gate_width) = mel_cond_{i+1}, the last skip_width columns = mel_cond_out1 (biases added). */
int nsw_fastgen_cond_vars_device(nsw_fastgen* h, const float* d_encoding, int32_t B, int32_t T, float* d_out,
                                 void* stream);
int nsw_fastgen_cond_vars_host(nsw_fastgen* h, const float* encoding, int32_t B, int32_t T, float* out);

/* PARITY HOOK: the sampler's random draws of the following nsw_fastgen_run_* calls are read from `noise` instead of
 * the in-kernel Philox stream, so that the int32 samples can be compared with loss_func.mol_sample / gauss_sample /
 * ce_sample (loss_func.py:140-206) evaluated on the same draws.  noise [B][T][nu] (host pointer, or device pointer if
 * on_device != 0; copied):
 *   mol   nu = nr_mix + 1 : u1[0..nr_mix) then u2, uniforms in [1e-5, 1 - 1e-5] as tf.random_uniform draws them
 *                           (loss_func.py:166,181)
 *   gauss nu = 1          : n ~ N(0,1), the draw of Normal.sample() (loss_func.py:203)
 *   ce    nu = 1          : u in (0,1) for the inverse-CDF categorical draw
 * B and T must match the runs that consume it.  noise = NULL returns to Philox. */
int nsw_fastgen_set_noise(nsw_fastgen* h, const float* noise, int32_t B, int32_t T, int32_t nu, int32_t on_device);
/* device time of the last nsw_fastgen_run_* (cond GEMM + persistent kernel) in ms */
int nsw_fastgen_last_timing(nsw_fastgen* h, float* ms);
/* TEST HOOK (host only, no CUDA): the create-time repacking of the TF-named tensors into
 * the persistent kernel's per-(phase, CTA) weight blocks and the hoisted conditioning GEMM.
 * sizes[0] = floats in `blocks`, sizes[1] = floats per block, sizes[2] = CTAs,
 * sizes[3] = columns of cond_w.  blocks / cond_w / cond_b may be NULL to query sizes. */
int nsw_fastgen_pack_host(const nsw_wavenet_config* cfg, const nsw_tensor* tensors,
                          int32_t n_tensors, float* blocks, int64_t blocks_cap, float* cond_w,
                          float* cond_b, int64_t* sizes);

/* TEST HOOK (host only, no CUDA): create-time repacking of the batched fastgen engine (nsw_fastgen_gn.cu: gate 512 or
 * 1024, mol / gauss / ce heads, up to 8 utterances per weight pass).  sizes[0] = floats in `blocks`, [1] = floats per
 * block, [2] = CTAs, [3] = columns of cond_w, [4] = gate_width / 2, [5] = phases per step (num_layers + 3). */
int nsw_fastgen_gn_pack_host(const nsw_wavenet_config* cfg, const nsw_tensor* tensors, int32_t n_tensors,
                             float* blocks, int64_t blocks_cap, float* cond_w, float* cond_b, int64_t* sizes);

/* TEST HOOK (host only, no CUDA): the same plan for the CTA-pair flow kernel (nsw_iaf_flow_pair.cu; the default
 * wherever it covers the launch: whole flow, an even number of 128-row tiles per clip, all clips at once).  Per CTA
 * (cluster rank fastest): {clip, first tile, tiles, CTAs of the clip, index within the clip of the last CTA whose
 * "consumed" counter this CTA's publisher polls} then nl*tiles records as above.  max_pairs = CTA pairs the device can
 * hold.  Returns the grid size (>0) or a negative NSW_E* code (shape not covered: NSW_EINVAL). */
int nsw_flow_pair_plan_host(int32_t T, int32_t nclips, int32_t max_pairs, int32_t nl, int32_t num_stages, int32_t* out,
                            int64_t cap, int64_t* n_out);

/* KERNEL-LEVEL PARITY HOOK: the tcgen05 conv-GEMM every dense contraction of the path runs on (masked.conv1d as a
 * GEMM, masked.py:160-232; trans_conv1d phases, masked.py:235-291), with fp32 operands split on the device:
 *   out[(clip, m), n] = bias[n] + sum_{tap, c} x[clip, m + a_off + tap*tap_stride, c] * w[tap*cin + c, n]
 *                       + sum_c x2[clip, m + a_off2, c] * w[ntaps*cin + c, n]      (x2 != NULL: second source)
 *                       + y[(clip, m), n]                                          (y  != NULL: accumulate source)
 * rows outside [0, L) / [0, L2) read as zeros.  cin, cin2 multiples of 64; N a multiple of 64.  flags: 1 = small split
 * products in their own accumulator, 2 = the one-CTA 128 x 128 kernel instead of the CTA-pair kernel. */
int nsw_conv_gemm_device(const float* d_x, int32_t nclips, int32_t L, int32_t cin, int32_t ntaps, int32_t a_off,
                         int32_t tap_stride, int32_t mclip, const float* d_w, int32_t N, const float* d_bias,
                         const float* d_x2, int32_t L2, int32_t cin2, int32_t a_off2, const float* d_y, int32_t flags,
                         float* d_out, void* stream);

/* TEST HOOK (host only, no CUDA): work split and publish / read plan of one launch of the persistent IAF flow kernel
 * (engine NSW_ENGINE_TC3), computed by the same integer functions the kernel uses.  Per CTA: {clip, first tile, tiles,
 * CTAs of the clip} then (l1-l0)*tiles records {published, source of tap t-2d, source of tap t-d} with source >= 0 a
 * foreign tile of the clip read from the published global copy, -1 own shared memory, -2 causal zeros.  Returns the
 * grid size (>0) or a negative NSW_E* code; *n_out = number of ints written (or needed when cap is too small). */
int nsw_flow_plan_host(int32_t T, int32_t nclips, int32_t num_sms, int32_t l0, int32_t l1, int32_t num_stages,
                       int32_t fuse_head, int32_t* out, int64_t cap, int64_t* n_out);

/* ---- teacher full-sequence forward + distillation cross-entropy (BASELINE config 5) --------
 * replaces: Wavenet.feed_forward (wavenet.py:180-291) as called by
 * ParallelWavenet.kl_loss_logistic (parallel_wavenet.py:382) — every contraction on tcgen05. */
typedef struct nsw_teacher nsw_teacher;
int nsw_teacher_create(const nsw_wavenet_config* cfg, const nsw_tensor* tensors, int32_t n_tensors,
                       int32_t device, nsw_teacher** out);
void nsw_teacher_destroy(nsw_teacher* h);
/* wav_scaled [B,T] (T % 128 == 0), mel [B,F,num_mel] with F*prod(strides) >= T (centre-trimmed,
 * wavenet.py:76-85) -> out_params [B,T,out_width] */
int nsw_teacher_forward_device(nsw_teacher* h, const float* d_wav, const float* d_mel, int32_t B,
                               int32_t T, int32_t F, float* d_out_params, void* stream);
int nsw_teacher_forward_host(nsw_teacher* h, const float* wav, const float* mel, int32_t B, int32_t T,
                             int32_t F, float* out_params);
int nsw_teacher_last_timing(nsw_teacher* h, float* ms);
/* replaces: loss_func.mol_log_probs (loss_func.py:22-63) on num_samples logistic draws per
 * (b,t) from the student's (mean_tot, scale_tot) without materialising the xS tile
 * (parallel_wavenet.py:373-398).  d_eps [S,B,T] logistic noise or NULL (Philox from seed).
 * result[0] = H_Ps, result[1] = H_Ps_Pt, result[2] = kl_loss (host doubles; synchronises). */
int nsw_mol_score_device(nsw_teacher* h, const float* d_te_out_params, const float* d_mean_tot,
                         const float* d_scale_tot, const float* d_log_scale_tot, const float* d_eps,
                         uint64_t seed, int32_t num_samples, int32_t B, int32_t T, double* result,
                         void* stream);

/* replaces: ParallelWavenet.kl_loss_gauss (parallel_wavenet.py:404-428) downstream of the teacher forward, with
 * loss_func.mean_std_from_out_params (loss_func.py:66-75): closed-form KL between the student's
 * N(mean_tot, scale_tot) and a GAUSS teacher's out_params [B,T,2], one HBM-bound reduction.
 * result[0] = mean KL term, result[1] = mean (log s_p - log s_q)^2, result[2] = kl_loss = [0] + 4*[1]
 * (host doubles; synchronises). */
int nsw_gauss_kl_device(nsw_teacher* h, const float* d_te_out_params, const float* d_mean_tot,
                        const float* d_scale_tot, const float* d_log_scale_tot, int32_t B, int32_t T,
                        double* result, void* stream);

/* ---- mel front-end ---------------------------------------------------------
 * replaces: auxilaries/mel_extractor.py:31-90 (melspectrogram / batch_melspectrogram; librosa STFT centre=True,
 * reflect padding, hann window zero-padded to n_fft; Slaney filterbank; 20 log10(max(min_amp, .)); normalise
 * against min_level_db).  twiddle_cos / twiddle_sin [win][n_bins]: window-folded DFT tables of the non-zero
 * window taps; mel_basis [n_mel][n_bins].  frames = 1 + N / hop. */
typedef struct nsw_mel nsw_mel;
int nsw_mel_create(int32_t device, int32_t n_bins, int32_t win, int32_t hop, int32_t n_mel,
                   const float* twiddle_cos, const float* twiddle_sin, const float* mel_basis, float min_amp,
                   float min_level_db, nsw_mel** out);
void nsw_mel_destroy(nsw_mel* h);
int nsw_mel_frames(nsw_mel* h, int32_t n_samples);
/* wav [B,N] (N > n_fft/2) -> mel [B, frames, n_mel] in [0,1] */
int nsw_mel_device(nsw_mel* h, const float* d_wav, int32_t B, int32_t N, float* d_mel, void* stream);
int nsw_mel_host(nsw_mel* h, const float* wav, int32_t B, int32_t N, float* mel);

/* ---- power-loss STFT ---------------------------------------------------------
 * replaces: mel_extractor._tf_stft (mel_extractor.py:111-121) = tf.contrib.signal.stft(frame_length 800, frame_step 200,
 * fft_length 2048, pad_end=True): frame j starts at sample j*hop, zeros past the end, ceil(N / hop) frames.  A handle
 * created by nsw_mel_create with twiddle tables for tap n at position n is switched to that framing with
 * nsw_mel_set_framing(h, 0, 0) (shift = sample index of tap 0 of frame 0; reflect = 1 restores librosa's centred,
 * reflect-padded framing with shift = -win/2). */
int nsw_mel_set_framing(nsw_mel* h, int32_t shift, int32_t reflect);
/* wav [B,N] -> |STFT| [B, frames, n_bins] */
int nsw_stft_mag_device(nsw_mel* h, const float* d_wav, int32_t B, int32_t N, float* d_mag, void* stream);
/* replaces: ParallelWavenet.power_loss (parallel_wavenet.py:459-479) with the shipped switches (|STFT| features,
 * squared difference, priority-frequency average :56-70); the longer of the two waves is centre-cropped (:430-435).
 * d_orig [B,N_orig], d_pred [B,N_pred]; result[0] = power_loss, [1] = mean over all bins, [2] = mean over the bins
 * below priority_freq (host doubles; synchronises). */
int nsw_power_loss_device(nsw_mel* h, const float* d_orig, int32_t N_orig, const float* d_pred, int32_t N_pred,
                          int32_t B, int32_t priority_freq, double* result, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSW_H_ */
