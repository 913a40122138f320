"""Drop-in for the reference's wavenet/parallelgen.py (same names, arguments and
side effects), running the IAF student on B200 through libnsw_b200.so.

reference                                   here
---------                                   ----
load_parallelgen -> dict of TF tensors      dict with the same keys holding the engine
                                            ('x', 'mean_tot', ... map to output names)
synthesis(hparams, mel, save_paths, ckpt)   identical contract: writes 16 kHz float32 wavs
                                            of length (F*200//512)*512, logs the "Delay"
"""
from __future__ import annotations

import logging
import time

import numpy as np

from .. import checkpoint as ckpt
from ..engine import IAFEngine
from . import fastgen

log = logging.getLogger('nsynth_wavenet_b200')


def get_default_shadow_dict(tf_vars):
    """parallelgen.py:7-8 (variables are plain names here)."""
    return {getattr(v, 'name', v): v for v in tf_vars}


def _unshadowed(hparams):
    # parallel_wavenet.py:166-170: with use_teacher_deconv the deconv stack is frozen and
    # restored from its plain (non-EMA) name (parallelgen.py:32-39)
    if getattr(hparams, 'use_teacher_deconv', False):
        name = 'resize_conv' if getattr(hparams, 'use_resize_conv', False) else 'trans_conv'
        return ('iaf_share/{}'.format(name),)
    return ()


def load_parallelgen(hparams, batch_size=1, length=7680, num_mel=80, weights=None, device=0,
                     engine=None):
    """parallelgen.py:11-19.  `length` is the number of mel frames (the reference's
    placeholder is [batch_size, length, num_mel]).  Returns a dict with the reference's
    keys; the tensors are replaced by the engine that produces them."""
    if weights is None:
        raise ValueError('load_parallelgen needs `weights` (dict of TF-named arrays); '
                         'synthesis() loads them from checkpoint_path')
    eng = IAFEngine(hparams, weights, device=device, num_mel=num_mel, engine=engine)
    return {'engine': eng, 'mel_in': (batch_size, length, num_mel),
            'x': 'x', 'mean_tot': 'mean_tot', 'scale_tot': 'scale_tot',
            'log_scale_tot': 'log_scale_tot', 'rand_input': 'rand_input'}


def synthesis(hparams, mel, save_paths, checkpoint_path, seed=None, device=0, engine=None):
    """parallelgen.py:22-51."""
    batch_size, length, num_mel = mel.shape
    # graph build + Saver.restore of the reference (parallelgen.py:24-41): done once per checkpoint, then cached
    eng = ckpt.cached_engine(IAFEngine, 'iaf', hparams, checkpoint_path, _unshadowed(hparams),
                             device=device, num_mel=num_mel, engine=engine)
    if seed is None:
        seed = int(time.time_ns() & 0x7FFFFFFFFFFFFFFF)
    start = time.time()
    audio = eng.forward_host(np.asarray(mel, np.float32), z=None, seed=seed, quantize=True,
                             want=('x',))['x']
    cost = time.time() - start
    wave_length = audio.shape[1] / 16000
    log.info('Target waveform length {:.5f}, '
             'Session run consume {:.5f} secs, '
             'Delay {:.2f}'.format(wave_length, cost, cost / wave_length))
    fastgen.save_batch(audio, save_paths)
