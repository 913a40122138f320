"""Weight containers for the generation path.

The reference restores TF-V2 checkpoint bundles through tf.train.Saver with an
EMA-shadow name map (wavenet/fastgen.py:12-14,81-84; wavenet/parallelgen.py:30-41).
TensorFlow is not available here, so `checkpoint_path` is an ``.npz`` archive (or a
directory / prefix next to which ``<prefix>.npz`` exists) whose keys are the same TF
variable names, with or without the ``/ExponentialMovingAverage`` suffix.  When both
spellings are present the EMA shadow wins, except for variables the reference reads
un-shadowed (the frozen teacher deconv stack under use_teacher_deconv,
parallelgen.py:32-39)."""
from __future__ import annotations

import os

import numpy as np

EMA_SUFFIX = '/ExponentialMovingAverage'


def get_ema_shadow_dict(names):
    """fastgen.get_ema_shadow_dict (fastgen.py:12-14) on plain names."""
    return {'{}{}'.format(n, EMA_SUFFIX): n for n in names}


def get_default_shadow_dict(names):
    """parallelgen.get_default_shadow_dict (parallelgen.py:7-8)."""
    return {n: n for n in names}


def resolve_checkpoint(checkpoint_path):
    p = os.fspath(checkpoint_path)
    cands = [p, p + '.npz']
    if os.path.isdir(p):
        cands += sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith('.npz'))[::-1]
    for c in cands:
        if os.path.isfile(c) and c.endswith('.npz'):
            return c
    raise FileNotFoundError(
        'no .npz weight archive found for checkpoint_path={!r}; TF-V2 bundles are not '
        'readable without TensorFlow (export with tools/export_npz.py on the training '
        'side)'.format(checkpoint_path))


def load_weights(checkpoint_path, unshadowed_substrings=()):
    """-> dict plain TF variable name -> float32 array."""
    path = resolve_checkpoint(checkpoint_path)
    raw = np.load(path)
    out = {}
    for key in raw.files:
        name = key[:-2] if key.endswith(':0') else key
        if name.endswith(EMA_SUFFIX):
            base = name[:-len(EMA_SUFFIX)]
            if any(s in base for s in unshadowed_substrings) and base in raw.files:
                continue
            out[base] = np.asarray(raw[key], np.float32)
        elif name not in out:
            shadow = name + EMA_SUFFIX
            if shadow in raw.files and not any(s in name for s in unshadowed_substrings):
                continue
            out[name] = np.asarray(raw[key], np.float32)
    return out


def save_weights(path, weights, ema=True):
    """Write an .npz the loaders above accept (EMA-suffixed names by default)."""
    if not path.endswith('.npz'):
        path = path + '.npz'
    np.savez(path, **{(k + EMA_SUFFIX if ema else k): v for k, v in weights.items()})
    return path
