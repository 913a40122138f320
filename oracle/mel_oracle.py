"""TEST INFRASTRUCTURE (CPU oracle): the reference's mel front-end restated with NumPy float64.

Follows auxilaries/mel_extractor.py:31-90 of the reference, whose arithmetic lives in librosa (un-vendored,
un-pinned dependency; not installable here), so librosa's published algorithm is restated: centred STFT
(n_fft 2048, hop 200, periodic hann window of 800 zero-padded to n_fft, reflect padding — librosa.stft, called at
mel_extractor.py:68-72), Slaney mel filterbank (80 bins, 125-7600 Hz, area-normalised — librosa.filters.mel,
:83-86), 20*log10(max(1e-5, .)) (:76-77), normalise to [0,1] against -140 dB (:80-81).  PARITY UNPINNED against
librosa itself; the STFT is pinned against scipy.signal.stft in tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

mel_params = SimpleNamespace(  # mel_extractor.py:14-25
    sample_rate=16000, num_freq=1025, num_mel=80, frame_shift_ms=12.5, frame_length_ms=50,
    preemphasis=0.97, min_level_db=-140, ref_level_db=40, mel_fmin=125, mel_fmax=7600,
    min_amp=1e-5)

PRIORITY_FREQ = int(3000 / (mel_params.sample_rate * 0.5) * mel_params.num_freq)
FRAME_SHIFT = int(mel_params.frame_shift_ms * mel_params.sample_rate / 1000.)

_mel_basis = None


def _hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def _build_mel_basis(p=mel_params):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with Slaney normalisation
    (mel_extractor.py:83-86)."""
    n_fft = (p.num_freq - 1) * 2
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


def _stft(y, p=mel_params):
    """librosa.stft(center=True, reflect padding, hann) (mel_extractor.py:68-72)."""
    n_fft = (p.num_freq - 1) * 2
    hop = int(p.frame_shift_ms / 1000.0 * p.sample_rate)
    win_length = int(p.frame_length_ms / 1000.0 * p.sample_rate)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)  # periodic hann
    lpad = (n_fft - win_length) // 2
    window = np.zeros(n_fft)
    window[lpad:lpad + win_length] = win
    y = np.pad(np.asarray(y, np.float64), n_fft // 2, mode='reflect')
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = y[idx] * window[None, :]
    return np.fft.rfft(frames, axis=1).T  # [num_freq, frames]


def _amp_to_db(x):
    return 20 * np.log10(np.maximum(mel_params.min_amp, x))


def _normalize(S, min_level_db):
    return np.clip((S - min_level_db) / -min_level_db, 0, 1)


def melspectrogram(y):
    """mel_extractor.py:31-35 -> [frames, 80] float32, frames = 1 + len(y)//200."""
    global _mel_basis
    if _mel_basis is None:
        _mel_basis = _build_mel_basis()
    D = _stft(y)
    S = _amp_to_db(np.dot(_mel_basis, np.abs(D)))
    return _normalize(S, mel_params.min_level_db).T.astype(np.float32)


def batch_melspectrogram(y):
    """mel_extractor.py:38-44."""
    assert len(y.shape) == 2
    return np.array([melspectrogram(y[b]) for b in range(y.shape[0])])


# ---- power-loss STFT (mel_extractor.py:111-121, parallel_wavenet.py:430-435, 56-70, 459-479) -------------------
def tf_stft(y, p=mel_params):
    """mel_extractor._tf_stft = tf.contrib.signal.stft(y, frame_length, frame_step, fft_length, pad_end=True):
    frame j = y[j*step : j*step + frame_length] (zeros past the end), times the periodic hann window
    (tf.contrib.signal.hann_window, periodic=True by default), zero-padded at its end to fft_length, rfft.
    pad_end=True gives ceil(N / frame_step) frames.  y [B, N] -> complex [B, frames, num_freq].
    (TensorFlow's tf.contrib.signal is an un-vendored dependency; restated from its documented framing.)"""
    step = int(p.frame_shift_ms * p.sample_rate / 1000)
    length = int(p.frame_length_ms * p.sample_rate / 1000)
    n_fft = int(2 * (p.num_freq - 1))
    y = np.asarray(y, np.float64)
    B, N = y.shape
    frames = -(-N // step)
    ypad = np.concatenate([y, np.zeros((B, (frames - 1) * step + length - N if (frames - 1) * step + length > N else 0))], 1)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(length) / length)
    idx = np.arange(length)[None, :] + step * np.arange(frames)[:, None]
    return np.fft.rfft(ypad[:, idx] * win[None, None, :], n=n_fft, axis=2)


def trim(x, trim_len):
    """ParallelWavenet._trim (parallel_wavenet.py:430-435)."""
    left = int(trim_len // 2)
    return x[:, left:left + x.shape[1] - trim_len]


def power_loss(orig_wav, pred_wav, priority_freq=PRIORITY_FREQ):
    """ParallelWavenet.power_loss (parallel_wavenet.py:459-479) with the shipped module switches (:11-30):
    SPEC_ENHANCE_FACTOR = 1 (features = |STFT|), USE_MEL = False, NORM_FEAT = False, USE_L1_LOSS = False (squared
    difference), USE_PRIORITY_FREQ = True (0.5 mean over all bins + 0.5 mean over the bins below PRIORITY_FREQ)."""
    orig_wav, pred_wav = np.asarray(orig_wav, np.float64), np.asarray(pred_wav, np.float64)
    lp, lo = pred_wav.shape[1], orig_wav.shape[1]
    if lp > lo:
        pred_wav = trim(pred_wav, lp - lo)
    elif lp < lo:
        orig_wav = trim(orig_wav, lo - lp)
    diff = (np.abs(tf_stft(orig_wav)) - np.abs(tf_stft(pred_wav))) ** 2
    return 0.5 * diff.mean() + 0.5 * diff[:, :, :priority_freq].mean()
