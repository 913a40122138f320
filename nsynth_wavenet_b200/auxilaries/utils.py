"""Host-side signal helpers with the reference's names (auxilaries/utils.py).
Only the NumPy halves are needed on the host; the TF halves run inside the CUDA
kernels (iaf_head_kernel, fastgen sampler)."""
from __future__ import annotations

import os

import numpy as np
from scipy.io import wavfile


def shell_path(path):
    """utils.shell_path (utils.py:33-52): expand ~ and make absolute."""
    return os.path.abspath(os.path.expanduser(os.path.expandvars(path)))


def load_audio(path, sample_length=64000, sr=16000):
    """utils.load_audio (utils.py:55-69) without librosa: 16 kHz mono wav only."""
    rate, data = wavfile.read(path)
    if rate != sr:
        raise ValueError('{}: sample rate {} != {} (resampling needs librosa/sox, see '
                         'tools/sox_downsample.py in the reference)'.format(path, rate, sr))
    if data.ndim > 1:
        data = data.mean(axis=1)
    if data.dtype == np.uint8:
        # 8-bit PCM is unsigned with the zero line at 128 (the only unsigned WAV sample format)
        data = (data.astype(np.float32) - 128.0) / 128.0
    elif np.issubdtype(data.dtype, np.integer):
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    audio = np.asarray(data, np.float32)
    if sample_length > 0:
        audio = audio[:sample_length]
    return audio


def mu_law_numpy(x, mu=255, int8=False):
    """utils.py:90-105."""
    out = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    out = np.floor(out * 128)
    return out.astype(np.int8) if int8 else out


def inv_mu_law_numpy(x, mu=255.0):
    """utils.py:125-139."""
    x = np.array(x).astype(np.float32)
    out = (x + 0.5) * 2. / (mu + 1)
    out = np.sign(out) / mu * ((1 + mu) ** np.abs(out) - 1)
    return np.where(np.equal(x, 0), x, out)


def cast_quantize_numpy(x, quant_chann):
    """utils.py:162-164."""
    return (x * quant_chann / 2).astype(np.int32)


def inv_cast_quantize_numpy(x_quantized, quant_chann):
    """utils.py:167-169."""
    return x_quantized.astype(np.float32) / (quant_chann / 2)
