// Mel front-end on the GPU: wav -> normalised log-mel frames, the input format of both
// generation paths.
//
// Replaces auxilaries/mel_extractor.py:31-90 (melspectrogram / batch_melspectrogram, computed by librosa on
// the CPU in the reference): centred STFT with reflect padding (n_fft 2048, hop 200, hann window of 800
// samples zero-padded to n_fft) -> |.| -> Slaney mel filterbank -> 20 log10(max(min_amp, .)) -> normalise
// against min_level_db.  Only the 800 non-zero window taps contribute, so a frame is a 800 x 1025 complex
// contraction against window-folded twiddle tables prepared on the host in float64.
//
// Two kernels, both small next to the generation kernels (1.6 MFLOP per frame):
//   stft_mag_kernel : 32 frames x 64 bins per CTA.  The 32 overlapping frames are one contiguous span of
//                     31*hop + win samples, staged ONCE in shared memory (reflection applied while staging);
//                     the twiddle tile is staged per 32-tap chunk; each thread owns 8 frames x 1 bin.
//   mel_log_kernel  : one CTA per frame; the frame's magnitudes in shared memory, thread m owns mel bin m and
//                     reads the transposed filterbank coalesced; log / normalise / clip in the epilogue.
#include "nsw_common.cuh"

#include <vector>

namespace nsw {
namespace {

constexpr int MF_FRAMES = 32;   // frames per CTA
constexpr int MF_BINS = 64;     // frequency bins per CTA
constexpr int MF_KC = 32;       // window taps per staged twiddle chunk
constexpr int MF_THREADS = 256;

struct MelGeom {
  int n_bins;     // 1025
  int win;        // 800 non-zero window taps
  int hop;        // 200
  int n_mel;      // 80
  int shift;      // sample index of tap 0 of frame 0: lpad - n_fft/2 = -win/2 (librosa, centred); 0 for tf.contrib.signal.stft
  int reflect;    // 1: np.pad(mode='reflect') outside [0, N) (librosa centre=True); 0: zeros (tf stft, pad_end=True)
};

__global__ void __launch_bounds__(MF_THREADS)
stft_mag_kernel(const float* __restrict__ wav /*[B][stride], the clip at +off*/, const float* __restrict__ tc /*[win][n_bins]*/,
                const float* __restrict__ ts, float* __restrict__ mag /*[B][frames][n_bins]*/, MelGeom g, int N,
                int frames, int stride, int off) {
  extern __shared__ float sm[];
  const int span = (MF_FRAMES - 1) * g.hop + g.win;
  float* xs = sm;                          // [span]
  float* cs = xs + span;                   // [MF_KC][MF_BINS]
  float* ss = cs + MF_KC * MF_BINS;        // [MF_KC][MF_BINS]
  const int tid = threadIdx.x, tx = tid & (MF_BINS - 1), ty = tid / MF_BINS;  // ty in 0..3
  const int f0 = blockIdx.x * MF_FRAMES, b0 = blockIdx.y * MF_BINS, b = blockIdx.z;
  const float* w = wav + (size_t)b * stride + off;
  // stage the span; np.pad(mode='reflect') semantics: index -m -> m, N-1+m -> N-1-m
  const int base = f0 * g.hop + g.shift;
  for (int i = tid; i < span; i += MF_THREADS) {
    int m = base + i;
    if (g.reflect) {
      if (m < 0) m = -m;
      if (m >= N) m = 2 * (N - 1) - m;
    }
    xs[i] = (m >= 0 && m < N) ? w[m] : 0.f;   // frames past the last one of a ragged tile read zeros
  }
  float re[8], im[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { re[j] = 0.f; im[j] = 0.f; }
  const int bin = b0 + tx;
  for (int k0 = 0; k0 < g.win; k0 += MF_KC) {
    __syncthreads();
    for (int i = tid; i < MF_KC * MF_BINS; i += MF_THREADS) {
      const int kk = i / MF_BINS, bb = i % MF_BINS;
      const bool ok = (k0 + kk < g.win) && (b0 + bb < g.n_bins);
      cs[i] = ok ? __ldg(tc + (size_t)(k0 + kk) * g.n_bins + b0 + bb) : 0.f;
      ss[i] = ok ? __ldg(ts + (size_t)(k0 + kk) * g.n_bins + b0 + bb) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < MF_KC; ++kk) {
      const float c = cs[kk * MF_BINS + tx], s = ss[kk * MF_BINS + tx];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = xs[(ty + 4 * j) * g.hop + k0 + kk];   // broadcast inside the warp
        re[j] = fmaf(x, c, re[j]);
        im[j] = fmaf(x, s, im[j]);
      }
    }
  }
  if (bin < g.n_bins) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = f0 + ty + 4 * j;
      if (f < frames) mag[((size_t)b * frames + f) * g.n_bins + bin] = sqrtf(re[j] * re[j] + im[j] * im[j]);
    }
  }
}

__global__ void __launch_bounds__(128)
mel_log_kernel(const float* __restrict__ mag /*[BF][n_bins]*/, const float* __restrict__ basis_t /*[n_bins][n_mel]*/,
               float* __restrict__ mel /*[BF][n_mel]*/, MelGeom g, float min_amp, float min_level_db) {
  extern __shared__ float row[];
  const size_t r = blockIdx.x;
  for (int i = threadIdx.x; i < g.n_bins; i += blockDim.x) row[i] = mag[r * g.n_bins + i];
  __syncthreads();
  for (int m = threadIdx.x; m < g.n_mel; m += blockDim.x) {
    float acc = 0.f;
    for (int f = 0; f < g.n_bins; ++f) acc = fmaf(__ldg(basis_t + (size_t)f * g.n_mel + m), row[f], acc);
    const float db = 20.0f * log10f(fmaxf(min_amp, acc));                 // mel_extractor.py:76-77
    const float v = (db - min_level_db) / -min_level_db;                  // :80-81
    mel[r * g.n_mel + m] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

// sum over all (frame, bin) of (a - b)^2, and over the bins below `priority` (PWNHelper.diff_fn / avg_loss_fn,
// parallel_wavenet.py:56-70, with USE_L1_LOSS = False): fp64 block sums, two atomics per block
__global__ void __launch_bounds__(256)
power_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t rows, int n_bins, int priority,
                  double* __restrict__ acc /*[2]*/) {
  double s_all = 0.0, s_pri = 0.0;
  const size_t n = rows * (size_t)n_bins;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    const double q = (double)d * (double)d;
    s_all += q;
    if ((int)(i % n_bins) < priority) s_pri += q;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s_all;
  sh[1][threadIdx.x] = s_pri;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(acc, sh[0][0]);
    atomicAdd(acc + 1, sh[1][0]);
  }
}

}  // namespace
}  // namespace nsw

using namespace nsw;

struct nsw_mel {
  int device = 0;
  MelGeom g{};
  float min_amp = 1e-5f, min_level_db = -140.f;
  DevBuf tc, ts, basis_t, mag, mag2, acc, stage_wav, stage_wav2, stage_mel;
  cudaStream_t own_stream = nullptr;
};

extern "C" int nsw_mel_create(int32_t device, int32_t n_bins, int32_t win, int32_t hop, int32_t n_mel,
                              const float* twiddle_cos, const float* twiddle_sin, const float* mel_basis,
                              float min_amp, float min_level_db, nsw_mel** out) {
  NSW_CHECK(twiddle_cos && twiddle_sin && mel_basis && out, NSW_EINVAL, "nsw_mel_create: null argument");
  NSW_CHECK(n_bins >= 2 && win >= 2 && win % 2 == 0 && hop >= 1 && n_mel >= 1 && n_mel <= 1024, NSW_EINVAL,
            "bad mel geometry (%d bins, window %d, hop %d, %d mel)", n_bins, win, hop, n_mel);
  NSW_CHECK(min_amp > 0.f && min_level_db < 0.f, NSW_EINVAL, "bad min_amp / min_level_db");
  NSW_CUDA(cudaSetDevice(device));
  nsw_mel* h = new nsw_mel();
  h->device = device;
  h->g.n_bins = n_bins;
  h->g.win = win;
  h->g.hop = hop;
  h->g.n_mel = n_mel;
  h->g.shift = -win / 2;   // centre=True: frame j is centred on sample j*hop, the window on the frame
  h->g.reflect = 1;
  h->min_amp = min_amp;
  h->min_level_db = min_level_db;
  std::vector<float> bt((size_t)n_bins * n_mel);
  for (int m = 0; m < n_mel; ++m)
    for (int f = 0; f < n_bins; ++f) bt[(size_t)f * n_mel + m] = mel_basis[(size_t)m * n_bins + f];
  int rc = upload(h->tc, twiddle_cos, (size_t)win * n_bins * 4);
  if (rc == NSW_OK) rc = upload(h->ts, twiddle_sin, (size_t)win * n_bins * 4);
  if (rc == NSW_OK) rc = upload(h->basis_t, bt.data(), bt.size() * 4);
  const size_t smem = ((size_t)(MF_FRAMES - 1) * hop + win + 2 * MF_KC * MF_BINS) * sizeof(float);
  if (rc == NSW_OK) {
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024)
      e = cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      set_error("nsw_mel_create: %s", cudaGetErrorString(e));
      rc = NSW_ECUDA;
    }
  }
  if (rc != NSW_OK) {
    nsw_mel_destroy(h);
    return rc;
  }
  *out = h;
  return NSW_OK;
}

extern "C" void nsw_mel_destroy(nsw_mel* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

// frames = 1 + N / hop (librosa.stft with center=True)
extern "C" int nsw_mel_frames(nsw_mel* h, int32_t n_samples) {
  if (!h || n_samples < 0) return NSW_EINVAL;
  return 1 + n_samples / h->g.hop;
}

extern "C" int nsw_mel_device(nsw_mel* h, const float* d_wav, int32_t B, int32_t N, float* d_mel, void* stream) {
  NSW_CHECK(h && d_wav && d_mel, NSW_EINVAL, "null argument");
  // np.pad(mode='reflect') by n_fft/2 needs N > n_fft/2; only the window half matters to the arithmetic
  NSW_CHECK(B >= 1 && N > h->g.n_bins - 1, NSW_EINVAL, "mel: need more than %d samples per clip (got %d)",
            h->g.n_bins - 1, N);
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int frames = 1 + N / h->g.hop;
  NSW_TRY(h->mag.ensure((size_t)B * frames * h->g.n_bins * sizeof(float)));
  const size_t smem = ((size_t)(MF_FRAMES - 1) * h->g.hop + h->g.win + 2 * MF_KC * MF_BINS) * sizeof(float);
  dim3 grid((frames + MF_FRAMES - 1) / MF_FRAMES, (h->g.n_bins + MF_BINS - 1) / MF_BINS, B);
  NSW_CHECK(h->g.reflect == 1, NSW_EINVAL, "this handle was switched to tf.contrib.signal.stft framing (nsw_mel_set_framing)");
  stft_mag_kernel<<<grid, MF_THREADS, smem, st>>>(d_wav, h->tc.as<float>(), h->ts.as<float>(), h->mag.as<float>(),
                                                  h->g, N, frames, N, 0);
  count_launch();
  mel_log_kernel<<<(unsigned)((size_t)B * frames), 128, h->g.n_bins * sizeof(float), st>>>(
      h->mag.as<float>(), h->basis_t.as<float>(), d_mel, h->g, h->min_amp, h->min_level_db);
  count_launch();
  NSW_CUDA(cudaGetLastError());
  return NSW_OK;
}

extern "C" int nsw_mel_host(nsw_mel* h, const float* wav, int32_t B, int32_t N, float* mel) {
  NSW_CHECK(h && wav && mel, NSW_EINVAL, "null argument");
  NSW_CHECK(B >= 1 && N >= 1, NSW_EINVAL, "bad batch/length %d/%d", B, N);
  NSW_CUDA(cudaSetDevice(h->device));
  const int frames = 1 + N / h->g.hop;
  const size_t nw = (size_t)B * N * 4, nm = (size_t)B * frames * h->g.n_mel * 4;
  NSW_TRY(h->stage_wav.ensure(nw));
  NSW_TRY(h->stage_mel.ensure(nm));
  NSW_CUDA(cudaMemcpyAsync(h->stage_wav.p, wav, nw, cudaMemcpyHostToDevice, h->own_stream));
  NSW_TRY(nsw_mel_device(h, h->stage_wav.as<float>(), B, N, h->stage_mel.as<float>(), h->own_stream));
  NSW_CUDA(cudaMemcpyAsync(mel, h->stage_mel.p, nm, cudaMemcpyDeviceToHost, h->own_stream));
  NSW_CUDA(cudaStreamSynchronize(h->own_stream));
  return NSW_OK;
}


// ---- power-loss STFT ------------------------------------------------------------------------------
// mel_extractor._tf_stft (mel_extractor.py:111-121) = tf.contrib.signal.stft(y, frame_length, frame_step, fft_length,
// pad_end=True): frame j starts AT sample j*hop (not centred), zeros past the end, ceil(N / hop) frames, the periodic
// hann window on the first frame_length samples of the fft_length frame.  The twiddle tables of such a handle are
// built for tap n at position n (the Python side does that) and the framing is switched here.
extern "C" int nsw_mel_set_framing(nsw_mel* h, int32_t shift, int32_t reflect) {
  NSW_CHECK(h, NSW_EINVAL, "null handle");
  h->g.shift = shift;
  h->g.reflect = reflect ? 1 : 0;
  return NSW_OK;
}

static int stft_frames_of(const nsw_mel* h, int N) {
  return h->g.reflect ? 1 + N / h->g.hop : (N + h->g.hop - 1) / h->g.hop;
}

// |STFT| of wav [B][N] -> mag [B][frames][n_bins]; frames = ceil(N / hop) (tf framing) or 1 + N / hop (librosa)
extern "C" int nsw_stft_mag_device(nsw_mel* h, const float* d_wav, int32_t B, int32_t N, float* d_mag, void* stream) {
  NSW_CHECK(h && d_wav && d_mag, NSW_EINVAL, "null argument");
  NSW_CHECK(B >= 1 && N >= 1, NSW_EINVAL, "bad batch/length %d/%d", B, N);
  NSW_CUDA(cudaSetDevice(h->device));
  const int frames = stft_frames_of(h, N);
  const size_t smem = ((size_t)(MF_FRAMES - 1) * h->g.hop + h->g.win + 2 * MF_KC * MF_BINS) * sizeof(float);
  dim3 grid((frames + MF_FRAMES - 1) / MF_FRAMES, (h->g.n_bins + MF_BINS - 1) / MF_BINS, B);
  stft_mag_kernel<<<grid, MF_THREADS, smem, (cudaStream_t)stream>>>(d_wav, h->tc.as<float>(), h->ts.as<float>(), d_mag,
                                                                    h->g, N, frames, N, 0);
  count_launch();
  NSW_CUDA(cudaGetLastError());
  return NSW_OK;
}

// replaces: ParallelWavenet.power_loss (parallel_wavenet.py:459-479) with the shipped switches (SPEC_ENHANCE_FACTOR 1:
// |STFT|; USE_L1_LOSS False: squared difference; USE_PRIORITY_FREQ True: 0.5 mean(all bins) + 0.5 mean(bins below
// priority_freq); NORM_FEAT False).  The longer wave is centre-cropped (_trim, :430-435).  result[0] = power_loss,
// [1] = mean squared difference over all bins, [2] = over the priority bins (host doubles; synchronises).
extern "C" int nsw_power_loss_device(nsw_mel* h, const float* d_orig, int32_t N_orig, const float* d_pred,
                                     int32_t N_pred, int32_t B, int32_t priority_freq, double* result, void* stream) {
  NSW_CHECK(h && d_orig && d_pred && result, NSW_EINVAL, "null argument");
  NSW_CHECK(B >= 1 && N_orig >= 1 && N_pred >= 1, NSW_EINVAL, "bad shapes");
  NSW_CHECK(h->g.reflect == 0, NSW_EINVAL, "power loss needs a handle with tf.contrib.signal.stft framing");
  NSW_CHECK(priority_freq >= 0 && priority_freq <= h->g.n_bins, NSW_EINVAL, "bad priority_freq %d", priority_freq);
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = std::min(N_orig, N_pred);
  const int off_o = (N_orig - N) / 2, off_p = (N_pred - N) / 2;   // tf.slice(x, [0, trim_len // 2], ...)
  const int frames = stft_frames_of(h, N);
  const size_t nm = (size_t)B * frames * h->g.n_bins;
  NSW_TRY(h->mag.ensure(nm * sizeof(float)));
  NSW_TRY(h->mag2.ensure(nm * sizeof(float)));
  NSW_TRY(h->acc.ensure(2 * sizeof(double)));
  const size_t smem = ((size_t)(MF_FRAMES - 1) * h->g.hop + h->g.win + 2 * MF_KC * MF_BINS) * sizeof(float);
  dim3 grid((frames + MF_FRAMES - 1) / MF_FRAMES, (h->g.n_bins + MF_BINS - 1) / MF_BINS, B);
  stft_mag_kernel<<<grid, MF_THREADS, smem, st>>>(d_orig, h->tc.as<float>(), h->ts.as<float>(), h->mag.as<float>(), h->g,
                                                  N, frames, N_orig, off_o);
  stft_mag_kernel<<<grid, MF_THREADS, smem, st>>>(d_pred, h->tc.as<float>(), h->ts.as<float>(), h->mag2.as<float>(), h->g,
                                                  N, frames, N_pred, off_p);
  count_launch(2);
  NSW_CUDA(cudaMemsetAsync(h->acc.p, 0, 2 * sizeof(double), st));
  const int blocks = (int)std::min<size_t>(592, (nm + 255) / 256);
  power_diff_kernel<<<blocks, 256, 0, st>>>(h->mag.as<float>(), h->mag2.as<float>(), (size_t)B * frames, h->g.n_bins,
                                           priority_freq, h->acc.as<double>());
  count_launch();
  NSW_CUDA(cudaGetLastError());
  double sums[2];
  NSW_CUDA(cudaMemcpyAsync(sums, h->acc.p, sizeof(sums), cudaMemcpyDeviceToHost, st));
  NSW_CUDA(cudaStreamSynchronize(st));
  const double all = sums[0] / (double)nm;
  const double pri = priority_freq > 0 ? sums[1] / ((double)B * frames * priority_freq) : 0.0;
  result[1] = all;
  result[2] = pri;
  result[0] = priority_freq > 0 ? 0.5 * all + 0.5 * pri : all;
  return NSW_OK;
}
