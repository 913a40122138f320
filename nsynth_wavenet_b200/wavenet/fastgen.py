"""Drop-in for the reference's wavenet/fastgen.py: same public names and contracts;
the per-sample Python/Session loop (fastgen.py:156-168) is one persistent CUDA kernel."""
from __future__ import annotations

import logging
import os
import time

import numpy as np
from scipy.io import wavfile

from .. import checkpoint as ckpt
from ..auxilaries import mel_extractor, utils
from ..engine import FastgenEngine

log = logging.getLogger('nsynth_wavenet_b200')


def get_ema_shadow_dict(tf_vars):
    """fastgen.py:12-14."""
    names = [getattr(v, 'name', v) for v in tf_vars]
    names = [n[:-2] if n.endswith(':0') else n for n in names]
    return {'{}/ExponentialMovingAverage'.format(n): v for n, v in zip(names, tf_vars)}


def load_batch(files, sample_length=64000):
    """fastgen.py:17-52: list of .wav/.npy paths -> padded np array [B, Lmax(, dims)]."""
    batch_data = []
    max_length = 0
    is_npy = (os.path.splitext(files[0])[1] == '.npy')
    for f in files:
        data = np.load(f) if is_npy else utils.load_audio(f, sample_length, sr=16000)
        batch_data.append(data)
        max_length = max(max_length, data.shape[0])
    for i, data in enumerate(batch_data):
        if data.shape[0] < max_length:
            if is_npy:
                padded = np.zeros([max_length, data.shape[1]])
                padded[:data.shape[0], :] = data
            else:
                padded = np.zeros([max_length])
                padded[:data.shape[0]] = data
            batch_data[i] = padded
    return np.vstack(batch_data)


def save_batch(batch_audio, batch_save_paths):
    """fastgen.py:55-58."""
    for audio, name in zip(batch_audio, batch_save_paths):
        log.info('Saving: %s' % name)
        wavfile.write(name, 16000, np.asarray(audio, np.float32))


def load_deconv_stack(hparams, batch_size=1, mel_length=320, num_mel=80, weights=None,
                      device=0, engine=None):
    """fastgen.py:61-66."""
    if weights is None:
        raise ValueError('load_deconv_stack needs `weights`')
    eng = FastgenEngine(hparams, weights, device=device, num_mel=num_mel, engine=engine)
    return {'engine': eng, 'mel_in': (batch_size, mel_length, num_mel), 'encoding': 'encoding'}


def encode(hparams, wav_data, checkpoint_path, device=0, engine=None):
    """fastgen.py:69-88: wav [B,L] -> mel (host) -> deconv stack -> [B, F*200, 256]."""
    if wav_data.ndim == 1:
        wav_data = np.expand_dims(wav_data, 0)
    mel_val = mel_extractor.batch_melspectrogram(wav_data)
    batch_size, mel_length, num_mel = mel_val.shape
    eng = ckpt.cached_engine(FastgenEngine, 'fastgen', hparams, checkpoint_path, device=device,
                             num_mel=num_mel, engine=engine)
    return eng.encode_host(mel_val)


def load_cond_layers(hparams, batch_size=1, en_length=320, weights=None, device=0, engine=None):
    """fastgen.py:91-97: the handle that evaluates Fastgen.cond_vars on an encoding placeholder."""
    if weights is None:
        raise ValueError('load_cond_layers needs `weights`')
    eng = FastgenEngine(hparams, weights, device=device, engine=engine)
    return {'engine': eng, 'encoding_in': (batch_size, en_length, hparams.deconv_width),
            'cond_vars': 'cond_vars'}


def calculate_cond_vars(hparams, encoding, checkpoint_path, device=0, engine=None):
    """fastgen.py:100-115: encoding [B, L, deconv_width] -> dict 'mel_cond_i' / 'mel_cond_out1' -> np[B, L, width]
    (the hoisted conditioning GEMM the persistent kernels consume, in the reference's channel order)."""
    eng = ckpt.cached_engine(FastgenEngine, 'fastgen', hparams, checkpoint_path, device=device,
                             num_mel=80, engine=engine)
    return eng.cond_vars_host(encoding)


def load_fastgen(hparams, batch_size=1, weights=None, device=0, engine=None):
    """fastgen.py:118-125."""
    if weights is None:
        raise ValueError('load_fastgen needs `weights`')
    eng = FastgenEngine(hparams, weights, device=device, engine=engine)
    return {'engine': eng, 'wav_in': (batch_size, 1),
            'encoding_in': (batch_size, hparams.deconv_width),
            'sample': 'sample', 'init_ops': (), 'push_ops': ()}


def synthesis(hparams, mel_encoding, save_paths, checkpoint_path, seed=None, device=0,
              engine=None):
    """fastgen.py:128-169: one sample per encoding step, queues start at zero, the
    previous dequantised sample is fed back; writes len(save_paths) wavs."""
    eng = ckpt.cached_engine(FastgenEngine, 'fastgen', hparams, checkpoint_path, device=device,
                             num_mel=80, engine=engine)
    if seed is None:
        seed = int(time.time_ns() & 0x7FFFFFFFFFFFFFFF)
    audio_batch = eng.run_host(np.asarray(mel_encoding, np.float32), seed=seed)
    save_batch(audio_batch, save_paths)
