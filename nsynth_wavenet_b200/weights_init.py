"""Random-init weight dicts with the reference's variable names, shapes and initialisers.

The reference creates its variables inside the graph builders: kernels N(0, 0.05) and zero
biases (wavenet/masked.py:166-167, 255-260), `out2_scale` bias -0.3
(wavenet/parallel_wavenet.py:87-103).  bench.py, smoke() and the engine cache use these
dicts where no checkpoint exists (there is no network for real ones); the draw order is
fixed so that a seed names one model (tests/test_host.py pins it against the test-side
generator).
"""
from __future__ import annotations

import numpy as np


def _get(hp, name, default):
    return getattr(hp, name, default)


def _conv(w, rng, name, k, cin, cout, std, bias_init=0.0):
    w[name + '/W'] = rng.normal(0.0, std, size=(1, k, cin, cout)).astype(np.float32)
    w[name + '/biases'] = np.full((cout,), bias_init, dtype=np.float32)


def _deconv(w, rng, prefix, hp, num_mel, std):
    cin = num_mel
    for i, (fl, _stride) in enumerate(hp.deconv_config):
        if _get(hp, 'use_resize_conv', False):      # masked.resize_conv1d -> conv1d variables (masked.py:294-322)
            base = '{}resize_conv_{:d}'.format(prefix, i + 1)
            w[base + '/W'] = rng.normal(0.0, std, size=(1, fl, cin, hp.deconv_width)).astype(np.float32)
            w[base + '/biases'] = np.zeros((hp.deconv_width,), np.float32)
        else:
            base = '{}trans_conv_{:d}'.format(prefix, i + 1)
            w[base + '/kernel'] = rng.normal(0.0, std, size=(1, fl, hp.deconv_width, cin)).astype(np.float32)
            w[base + '/bias'] = np.zeros((hp.deconv_width,), np.float32)
        cin = hp.deconv_width


def _jitter_biases(w, rng, bias_std):
    if bias_std > 0:
        for name in w:
            if name.endswith('/biases') or name.endswith('/bias'):
                w[name] = (w[name] + rng.normal(0, bias_std, w[name].shape)).astype(np.float32)


def init_student_weights(hp, seed=12345, num_mel=80, std=0.05, bias_std=0.0):
    """Variables of ParallelWavenet.feed_forward (parallel_wavenet.py:200-345)."""
    rng = np.random.default_rng(seed)
    w = {}
    width, k = hp.width, hp.filter_length
    share = _get(hp, 'use_share_deconv', False) or _get(hp, 'use_teacher_deconv', False)
    if share:
        _deconv(w, rng, 'iaf_share/', hp, num_mel, std)
    for f, nl in enumerate(hp.num_iaf_layers):
        p = 'iaf_{:d}'.format(f + 1)
        if not share:
            _deconv(w, rng, p + '/', hp, num_mel, std)
        _conv(w, rng, p + '/start_conv', k, 1, width, std)
        for i in range(nl):
            _conv(w, rng, '{}/dilated_conv_{:d}'.format(p, i + 1), k, width, width, std)
            _conv(w, rng, '{}/mel_cond_{:d}'.format(p, i + 1), 1, hp.deconv_width, width, std)
            _conv(w, rng, '{}/res_{:d}'.format(p, i + 1), 1, width // 2, width, std)
        _conv(w, rng, p + '/out1', 1, width, width, std)
        _conv(w, rng, p + '/mel_cond_out1', 1, hp.deconv_width, width, std)
        _conv(w, rng, p + '/out2_mean', 1, width, 1, std)
        _conv(w, rng, p + '/out2_scale', 1, width, 1, std, bias_init=-0.3)
    _jitter_biases(w, rng, bias_std)
    return w


def teacher_widths(hp):
    """(gate_width, out_width) of Wavenet / Fastgen (wavenet.py:106,117-129,204)."""
    gate = 2 * hp.width if _get(hp, 'double_gate_width', True) else hp.width
    qc = 2 ** 8 if hp.use_mu_law else 2 ** 16
    out_w = {'ce': qc, 'mol': 3 * _get(hp, 'mol_mix', 10), 'gauss': 2}[hp.loss_type]
    return gate, out_w


def init_teacher_weights(hp, seed=12345, num_mel=80, std=0.05, bias_std=0.0):
    """Variables of Wavenet.feed_forward / Fastgen.sample (wavenet.py:180-291, 379-514)."""
    rng = np.random.default_rng(seed)
    w = {}
    width, skip, k = hp.width, hp.skip_width, hp.filter_length
    gate, out_w = teacher_widths(hp)
    _deconv(w, rng, '', hp, num_mel, std)
    _conv(w, rng, 'conv_start', k, 1, width, std)
    _conv(w, rng, 'skip_start', 1, width, skip, std)
    for i in range(hp.num_layers):
        _conv(w, rng, 'dilated_conv_%d' % (i + 1), k, width, gate, std)
        _conv(w, rng, 'mel_cond_%d' % (i + 1), 1, hp.deconv_width, gate, std)
        _conv(w, rng, 'res_%d' % (i + 1), 1, gate // 2, width, std)
        _conv(w, rng, 'skip_%d' % (i + 1), 1, gate // 2, skip, std)
    _conv(w, rng, 'out1', 1, skip, skip, std)
    _conv(w, rng, 'mel_cond_out1', 1, hp.deconv_width, skip, std)
    _conv(w, rng, 'out2', 1, skip, out_w, std)
    _jitter_biases(w, rng, bias_std)
    return w
