"""Weight containers for the generation path.

The reference restores TF-V2 checkpoint bundles through tf.train.Saver with an
EMA-shadow name map (wavenet/fastgen.py:12-14,81-84; wavenet/parallelgen.py:30-41) and resolves
directories with tf.train.latest_checkpoint (eval_wavenet.py:22, eval_parallel_wavenet.py:22).
`checkpoint_path` may therefore be

* a TF-V2 bundle prefix (``model.ckpt-200000`` next to ``.index`` / ``.data-*``) or a directory
  holding one (its ``checkpoint`` state file picks the latest) — read without TensorFlow by
  ``tf_bundle.py``;
* an ``.npz`` archive (or a directory / prefix next to which ``<prefix>.npz`` exists),

with keys that are the TF variable names, with or without the ``/ExponentialMovingAverage`` suffix.
When both spellings are present the EMA shadow wins, except for variables the reference reads
un-shadowed (the frozen teacher deconv stack under use_teacher_deconv, parallelgen.py:32-39).
Optimizer slots and counters of a training checkpoint are skipped."""
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
    """-> ('npz', file) or ('bundle', prefix)."""
    from . import tf_bundle
    p = os.fspath(checkpoint_path)
    if os.path.isfile(p) and p.endswith('.npz'):
        return 'npz', p
    if tf_bundle.is_bundle_prefix(p):
        return 'bundle', p
    if os.path.isfile(p) and p.endswith('.index'):
        return 'bundle', p[:-len('.index')]
    if os.path.isfile(p + '.npz'):
        return 'npz', p + '.npz'
    if os.path.isdir(p):
        prefix = tf_bundle.latest_checkpoint(p)
        if prefix is not None:
            return 'bundle', prefix
        cands = sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith('.npz'))[::-1]
        if cands:
            return 'npz', cands[0]
    raise FileNotFoundError(
        'no TF-V2 checkpoint bundle (<prefix>.index + .data-*) and no .npz weight archive found for '
        'checkpoint_path={!r}'.format(checkpoint_path))


_SKIP_SUFFIXES = ('/Adam', '/Adam_1', '/Momentum', '/RMSProp', '/RMSProp_1')
_SKIP_NAMES = ('global_step', 'beta1_power', 'beta2_power')


def _is_model_variable(name):
    return not (name.endswith(_SKIP_SUFFIXES) or name.split('/')[-1] in _SKIP_NAMES)


def load_weights(checkpoint_path, unshadowed_substrings=()):
    """-> dict plain TF variable name -> float32 array."""
    kind, path = resolve_checkpoint(checkpoint_path)
    if kind == 'bundle':
        from . import tf_bundle
        raw = tf_bundle.read_bundle(path, names=_is_model_variable)
        files = list(raw)
    else:
        raw = np.load(path)
        files = raw.files
    out = {}
    for key in files:
        name = key[:-2] if key.endswith(':0') else key
        if name.endswith(EMA_SUFFIX):
            base = name[:-len(EMA_SUFFIX)]
            if any(s in base for s in unshadowed_substrings) and base in files:
                continue
            out[base] = np.asarray(raw[key], np.float32)
        elif name not in out:
            shadow = name + EMA_SUFFIX
            if shadow in files and not any(s in name for s in unshadowed_substrings):
                continue
            out[name] = np.asarray(raw[key], np.float32)
    return out


def save_weights(path, weights, ema=True):
    """Write an .npz the loaders above accept (EMA-suffixed names by default)."""
    if not path.endswith('.npz'):
        path = path + '.npz'
    np.savez(path, **{(k + EMA_SUFFIX if ema else k): v for k, v in weights.items()})
    return path


# ---- engine cache --------------------------------------------------------------------------------------------
# The reference rebuilds its graph and restores the checkpoint on EVERY synthesis / encode call
# (parallelgen.py:24-41, fastgen.py:73-84, 143-147).  Here a call costs about a millisecond of GPU time, so that
# per-call set-up (reading the bundle, repacking weights, tensor maps, workspace) would be all of it: engines are kept,
# keyed on what determines them, and dropped when the checkpoint on disk changes.
_ENGINES = {}
_MAX_ENGINES = 4


def _stamp(kind, path):
    files = [path] if kind == 'npz' else None
    if files is None:
        d, base = os.path.split(path)
        files = sorted(os.path.join(d or '.', f) for f in os.listdir(d or '.')
                       if f == base + '.index' or f.startswith(base + '.data-'))
    return tuple((f, os.path.getmtime(f), os.path.getsize(f)) for f in files)


def cached_engine(factory, tag, hparams, checkpoint_path, unshadowed=(), **kw):
    """factory(hparams, weights, **kw) -> engine; one engine per (tag, checkpoint files + mtimes, hparams, kw)."""
    kind, path = resolve_checkpoint(checkpoint_path)
    key = (tag, kind, os.path.realpath(path), _stamp(kind, path),
           tuple(sorted((k, repr(v)) for k, v in vars(hparams).items())), tuple(sorted(kw.items())))
    eng = _ENGINES.get(key)
    if eng is None or getattr(eng, '_h', None) is None:
        for old in [k for k in _ENGINES if k[:3] == key[:3] and k != key]:   # same file, stale contents
            _ENGINES.pop(old).close()
        while len(_ENGINES) >= _MAX_ENGINES:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = factory(hparams, load_weights(checkpoint_path, unshadowed), **kw)
        _ENGINES[key] = eng
    return eng


def clear_engine_cache():
    while _ENGINES:
        _ENGINES.popitem()[1].close()
