"""Mel feature extractor with the reference's interface (auxilaries/mel_extractor.py), computed on the GPU.

The reference computes mels on the CPU with librosa (mel_extractor.py:31-90): centred STFT (n_fft 2048,
hop 200, hann window 800 zero-padded to n_fft, reflect padding), Slaney mel filterbank (80 bins,
125-7600 Hz, area-normalised), 20*log10(max(1e-5, .)), normalise to [0,1] against -140 dB.  Here the
contraction runs in `csrc/nsw_mel.cu` behind `nsw_mel_*` (include/nsw.h); this module only prepares the
constant tables (window-folded twiddles, filterbank) in float64, like the weight repacking of the engines.
There is no CPU fallback: the NumPy restatement lives under `oracle/` and is what the tests compare against."""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np

from .. import _lib as L

mel_params = SimpleNamespace(  # mel_extractor.py:14-25
    sample_rate=16000, num_freq=1025, num_mel=80, frame_shift_ms=12.5, frame_length_ms=50,
    preemphasis=0.97, min_level_db=-140, ref_level_db=40, mel_fmin=125, mel_fmax=7600,
    min_amp=1e-5)

PRIORITY_FREQ = int(3000 / (mel_params.sample_rate * 0.5) * mel_params.num_freq)
FRAME_SHIFT = int(mel_params.frame_shift_ms * mel_params.sample_rate / 1000.)

_extractors = {}


def _hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def _build_mel_basis(p=mel_params):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with Slaney normalisation
    (mel_extractor.py:83-86) -> [num_mel, num_freq] float32."""
    fftfreqs = np.linspace(0, p.sample_rate / 2.0, p.num_freq)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(p.mel_fmin), _hz_to_mel(p.mel_fmax), p.num_mel + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((p.num_mel, p.num_freq))
    for i in range(p.num_mel):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:p.num_mel + 2] - mel_f[:p.num_mel])
    return (weights * enorm[:, None]).astype(np.float32)


def _build_twiddles(p=mel_params, centred=True):
    """Window-folded DFT tables of the non-zero window taps: [win, num_freq] cos and sin (float32 from float64).
    centred: tap n of the window sits at position lpad + n of the n_fft frame (librosa pads the window to n_fft
    centred, mel_extractor.py:68-72); not centred: at position n (tf.contrib.signal.stft zero-pads the windowed frame
    at its END to fft_length, mel_extractor.py:111-121)."""
    n_fft = (p.num_freq - 1) * 2
    win_length = int(p.frame_length_ms / 1000.0 * p.sample_rate)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)       # periodic hann
    lpad = (n_fft - win_length) // 2 if centred else 0
    pos = (np.arange(win_length) + lpad)[:, None] * np.arange(p.num_freq)[None, :]    # exact integers
    ang = 2.0 * np.pi * (pos % n_fft) / n_fft
    return (win[:, None] * np.cos(ang)).astype(np.float32), (win[:, None] * np.sin(ang)).astype(np.float32)


class MelExtractor:
    """Device handle: tables uploaded once, reusable across calls."""

    def __init__(self, device=0, p=mel_params):
        self.lib = L.load()
        self.p = p
        self.dev_index = device
        tc, ts = _build_twiddles(p)
        basis = np.ascontiguousarray(_build_mel_basis(p))
        self.win = tc.shape[0]
        self.hop = int(p.frame_shift_ms / 1000.0 * p.sample_rate)
        h = C.c_void_p()
        L.check(self.lib.nsw_mel_create(device, p.num_freq, self.win, self.hop, p.num_mel, L.ptr(tc), L.ptr(ts),
                                        L.ptr(basis), float(p.min_amp), float(p.min_level_db), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self.lib.nsw_mel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frames(self, n_samples):
        return 1 + n_samples // self.hop

    def host(self, wav):
        """wav [B, N] float -> [B, frames, num_mel] float32."""
        wav = np.ascontiguousarray(wav, np.float32)
        B, N = wav.shape
        out = np.empty((B, self.frames(N), self.p.num_mel), np.float32)
        L.check(self.lib.nsw_mel_host(self._h, L.ptr(wav), B, N, L.ptr(out)))
        return out

    def device(self, wav):
        """torch CUDA wav [B, N] -> torch CUDA [B, frames, num_mel] on the current stream."""
        import torch
        from ..engine import _dev_tensor
        _dev_tensor(wav, 'wav', (None, None), self.dev_index)
        B, N = wav.shape
        out = torch.empty((B, self.frames(N), self.p.num_mel), dtype=torch.float32, device=wav.device)
        st = torch.cuda.current_stream(wav.device).cuda_stream
        L.check(self.lib.nsw_mel_device(self._h, L.ptr(wav), B, N, L.ptr(out), st))
        return out


class TfStft(MelExtractor):
    """mel_extractor._tf_stft (mel_extractor.py:111-121) on the GPU: tf.contrib.signal.stft(frame_length 800,
    frame_step 200, fft_length 2048, pad_end=True) -> magnitudes, and the power loss built on it
    (parallel_wavenet.py:459-479)."""

    def __init__(self, device=0, p=mel_params):
        self.lib = L.load()
        self.p = p
        self.dev_index = device
        tc, ts = _build_twiddles(p, centred=False)
        basis = np.ascontiguousarray(_build_mel_basis(p))
        self.win = tc.shape[0]
        self.hop = int(p.frame_shift_ms / 1000.0 * p.sample_rate)
        h = C.c_void_p()
        L.check(self.lib.nsw_mel_create(device, p.num_freq, self.win, self.hop, p.num_mel, L.ptr(tc), L.ptr(ts),
                                        L.ptr(basis), float(p.min_amp), float(p.min_level_db), C.byref(h)))
        self._h = h
        L.check(self.lib.nsw_mel_set_framing(self._h, 0, 0))

    def frames(self, n_samples):
        return -(-n_samples // self.hop)      # pad_end=True: ceil

    def host(self, wav):
        raise NotImplementedError('TfStft produces STFT magnitudes (stft_mag) and the power loss, not mels')

    device = host

    def stft_mag(self, wav):
        """torch CUDA wav [B, N] -> |STFT| [B, ceil(N / 200), 1025] on the current stream."""
        import torch
        from ..engine import _dev_tensor
        _dev_tensor(wav, 'wav', (None, None), self.dev_index)
        B, N = wav.shape
        out = torch.empty((B, self.frames(N), self.p.num_freq), dtype=torch.float32, device=wav.device)
        st = torch.cuda.current_stream(wav.device).cuda_stream
        L.check(self.lib.nsw_stft_mag_device(self._h, L.ptr(wav), B, N, L.ptr(out), st))
        return out

    def power_loss(self, orig_wav, pred_wav, priority_freq=PRIORITY_FREQ):
        """ParallelWavenet.power_loss (parallel_wavenet.py:459-479), torch CUDA [B, N_orig] / [B, N_pred] ->
        dict(power_loss, all_bins, priority_bins)."""
        import torch
        from ..engine import _dev_tensor
        _dev_tensor(orig_wav, 'orig_wav', (None, None), self.dev_index)
        _dev_tensor(pred_wav, 'pred_wav', (orig_wav.shape[0], None), self.dev_index)
        res = (C.c_double * 3)()
        st = torch.cuda.current_stream(orig_wav.device).cuda_stream
        L.check(self.lib.nsw_power_loss_device(self._h, L.ptr(orig_wav), orig_wav.shape[1], L.ptr(pred_wav),
                                               pred_wav.shape[1], orig_wav.shape[0], int(priority_freq),
                                               C.byref(res), st))
        return {'power_loss': res[0], 'all_bins': res[1], 'priority_bins': res[2]}


def _extractor(device=0):
    if device not in _extractors:
        _extractors[device] = MelExtractor(device)
    return _extractors[device]


def melspectrogram(y, device=0):
    """mel_extractor.py:31-35 -> [frames, 80] float32, frames = 1 + len(y)//200."""
    return _extractor(device).host(np.asarray(y, np.float32)[None, :])[0]


def batch_melspectrogram(y, device=0):
    """mel_extractor.py:38-44."""
    assert len(y.shape) == 2
    return _extractor(device).host(y)
