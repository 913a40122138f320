"""CPU oracle for the WaveNet / Parallel-WaveNet generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``nsynth_wavenet_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module,
and only as the checker / CPU baseline, never as the thing shipped.

PARITY UNPINNED (values): the reference's arithmetic lives in TensorFlow 1.x
(un-vendored, un-pinned; call sites ``wavenet/masked.py:209,262,374-376,402``),
TensorFlow is not installable in this image, the reference ships no golden
tensors, and its tests print instead of assert.  What IS pinned, and asserted in
``tests/test_oracle.py``:
  * the closed-form causal conv == the reference's literal
    time_to_batch -> pad -> VALID conv -> batch_to_time pipeline (masked.py:72-232);
  * the queue form (masked.py:352-376) == the closed form (tap order, zero history);
  * trans_conv1d == the adjoint of a SAME/stride-s forward conv (TF's definition of
    conv2d_transpose) checked through torch autograd, and == torch conv_transpose1d
    with padding (k-s)//2;
  * the invariants the reference's tests print: scale_tot > 0,
    x == rand_input*scale_tot + mean_tot (tests/test_parallel_wavenet.py:63-64),
    teacher exp(-loss) ~ 1/65536-ish at random init (tests/test_wavenet.py:66-69),
    output lengths 154600 / 154112 for the 154480-sample fixture (tests/pred_data-*),
    the numpy scale transform of tests/test_scale.py:67-78,
    _clip_quant_scale of tests/test_clip_quant_scale.py:7-18.

Every function cites the reference file:line it restates.  All functions take a
``dtype`` (np.float64 "truth" / np.float32 twin) and explicit noise / weights.

Weight container: ``dict[str, np.ndarray]`` keyed by the TF variable names the
reference creates (checkpoint contract), e.g.
``iaf_1/dilated_conv_3/W`` [1,k,Cin,Cout], ``iaf_1/dilated_conv_3/biases`` [Cout],
``iaf_share/trans_conv_1/kernel`` [1,k,Cout,Cin], ``.../bias`` [Cout].
"""
from __future__ import annotations

import math
from argparse import Namespace

import numpy as np

# --------------------------------------------------------------------------
# hparams helpers (reference: wavenet/wavenet.py:97-128, parallel_wavenet.py:118-147)
# --------------------------------------------------------------------------
LEAKY_ALPHA = 0.4  # masked.py:33-34


def _get(hp, name, default):
    return getattr(hp, name, default)


def quant_chann_of(hp):
    """wavenet.py:117-120 / parallel_wavenet.py:137-140."""
    return 2 ** 8 if hp.use_mu_law else 2 ** 16


def teacher_out_width(hp):
    """wavenet.py:121-129."""
    if hp.loss_type == 'ce':
        return quant_chann_of(hp)
    if hp.loss_type == 'mol':
        return hp.mol_mix * 3
    if hp.loss_type == 'gauss':
        return 2
    raise ValueError('[{}] loss is not supported'.format(hp.loss_type))


def teacher_gate_width(hp):
    """wavenet.py:106,204: double_gate_width defaults True when absent."""
    return 2 * hp.width if _get(hp, 'double_gate_width', True) else hp.width


def upsample_act(name):
    """masked.py:28-36."""
    if name == 'tanh':
        return np.tanh
    if name == 'relu':
        return lambda v: np.maximum(v, 0)
    if name == 'leaky_relu':
        return lambda v: np.where(v >= 0, v, v * v.dtype.type(LEAKY_ALPHA))
    raise ValueError('Unsupported activation function for upsample layer')


# --------------------------------------------------------------------------
# random-init weights with the reference's shapes/names and initialisers
# (masked.py:166-167: N(0, 0.05) kernels, zero biases; parallel_wavenet.py:87-103:
#  out2_scale bias -0.3)
# --------------------------------------------------------------------------
def _conv_vars(w, rng, name, k, cin, cout, bias_init=0.0, std=0.05):
    w[name + '/W'] = rng.normal(0.0, std, size=(1, k, cin, cout)).astype(np.float32)
    w[name + '/biases'] = np.full((cout,), bias_init, dtype=np.float32)


def _deconv_vars(w, rng, prefix, hp, num_mel, std=0.05):
    cin = num_mel
    for i, (fl, s) in enumerate(hp.deconv_config):
        if _get(hp, 'use_resize_conv', False):   # masked.resize_conv1d creates conv1d variables (masked.py:309-318)
            base = '{}resize_conv_{:d}'.format(prefix, i + 1)
            w[base + '/W'] = rng.normal(0.0, std, size=(1, fl, cin, hp.deconv_width)).astype(np.float32)
            w[base + '/biases'] = np.zeros((hp.deconv_width,), np.float32)
        else:
            base = '{}trans_conv_{:d}'.format(prefix, i + 1)
            w[base + '/kernel'] = rng.normal(
                0.0, std, size=(1, fl, hp.deconv_width, cin)).astype(np.float32)
            w[base + '/bias'] = np.zeros((hp.deconv_width,), np.float32)
        cin = hp.deconv_width


def init_student_weights(hp, seed=12345, num_mel=80, std=0.05, bias_std=0.0):
    """Variables created by ParallelWavenet.feed_forward (parallel_wavenet.py:200-345).

    ``bias_std`` > 0 additionally randomises biases so that bias handling is
    exercised by the parity tests (the reference initialises them to 0)."""
    rng = np.random.default_rng(seed)
    w = {}
    width = hp.width
    k = hp.filter_length
    share = _get(hp, 'use_share_deconv', False) or _get(hp, 'use_teacher_deconv', False)
    if share:
        _deconv_vars(w, rng, 'iaf_share/', hp, num_mel, std)
    for f, nl in enumerate(hp.num_iaf_layers):
        p = 'iaf_{:d}'.format(f + 1)
        if not share:
            _deconv_vars(w, rng, p + '/', hp, num_mel, std)
        _conv_vars(w, rng, p + '/start_conv', k, 1, width, std=std)
        for i in range(nl):
            _conv_vars(w, rng, '{}/dilated_conv_{:d}'.format(p, i + 1), k, width, width, std=std)
            _conv_vars(w, rng, '{}/mel_cond_{:d}'.format(p, i + 1), 1, hp.deconv_width, width, std=std)
            _conv_vars(w, rng, '{}/res_{:d}'.format(p, i + 1), 1, width // 2, width, std=std)
        _conv_vars(w, rng, p + '/out1', 1, width, width, std=std)
        _conv_vars(w, rng, p + '/mel_cond_out1', 1, hp.deconv_width, width, std=std)
        _conv_vars(w, rng, p + '/out2_mean', 1, width, 1, std=std)
        _conv_vars(w, rng, p + '/out2_scale', 1, width, 1, bias_init=-0.3, std=std)
    if bias_std > 0:
        for name in w:
            if name.endswith('/biases') or name.endswith('/bias'):
                w[name] = (w[name] + rng.normal(0, bias_std, w[name].shape)).astype(np.float32)
    return w


def init_teacher_weights(hp, seed=12345, num_mel=80, std=0.05, bias_std=0.0):
    """Variables created by Wavenet.feed_forward / Fastgen.sample (wavenet.py:180-291,379-514)."""
    rng = np.random.default_rng(seed)
    w = {}
    width, skip, k = hp.width, hp.skip_width, hp.filter_length
    gate = teacher_gate_width(hp)
    out_w = teacher_out_width(hp)
    _deconv_vars(w, rng, '', hp, num_mel, std)
    _conv_vars(w, rng, 'conv_start', k, 1, width, std=std)
    _conv_vars(w, rng, 'skip_start', 1, width, skip, std=std)
    for i in range(hp.num_layers):
        _conv_vars(w, rng, 'dilated_conv_%d' % (i + 1), k, width, gate, std=std)
        _conv_vars(w, rng, 'mel_cond_%d' % (i + 1), 1, hp.deconv_width, gate, std=std)
        _conv_vars(w, rng, 'res_%d' % (i + 1), 1, gate // 2, width, std=std)
        _conv_vars(w, rng, 'skip_%d' % (i + 1), 1, gate // 2, skip, std=std)
    _conv_vars(w, rng, 'out1', 1, skip, skip, std=std)
    _conv_vars(w, rng, 'mel_cond_out1', 1, hp.deconv_width, skip, std=std)
    _conv_vars(w, rng, 'out2', 1, skip, out_w, std=std)
    if bias_std > 0:
        for name in w:
            if name.endswith('/biases') or name.endswith('/bias'):
                w[name] = (w[name] + rng.normal(0, bias_std, w[name].shape)).astype(np.float32)
    return w


def fold_weight_norm(w):
    """masked.py:131-157: W = g * V / ||V||, norm over axes (0,1,2) for conv
    kernels ([1,k,Cin,Cout]) and (0,1,3) for deconv kernels ([1,k,Cout,Cin]).
    Returns a dict with every ``X_V``/``X_g`` pair replaced by ``X``."""
    out = {}
    for name, v in w.items():
        if name.endswith('_V'):
            base = name[:-2]
            g = w[base + '_g'].astype(np.float64)
            v64 = v.astype(np.float64)
            if base.endswith('/kernel'):
                nrm = np.sqrt((v64 ** 2).sum(axis=(0, 1, 3), keepdims=True))
                out[base] = (v64 / nrm * g.reshape(1, 1, -1, 1)).astype(np.float32)
            else:
                nrm = np.sqrt((v64 ** 2).sum(axis=(0, 1, 2), keepdims=True))
                out[base] = (v64 / nrm * g.reshape(1, 1, 1, -1)).astype(np.float32)
        elif name.endswith('_g'):
            continue
        else:
            out[name] = v
    return out


# --------------------------------------------------------------------------
# L2 ops (wavenet/masked.py)
# --------------------------------------------------------------------------
def shift_right(x):
    """masked.py:39-52: prepend one zero step, drop the last."""
    y = np.zeros_like(x)
    y[:, 1:] = x[:, :-1]
    return y


def time_to_batch(x, block):
    """masked.py:72-101."""
    b, t, c = x.shape
    y = x.reshape(b, t // block, block, c).transpose(0, 2, 1, 3)
    return y.reshape(b * block, t // block, c)


def batch_to_time(x, block):
    """masked.py:104-122."""
    bb, k, c = x.shape
    y = x.reshape(bb // block, block, k, c).transpose(0, 2, 1, 3)
    return y.reshape(bb // block, k * block, c)


def conv1d_literal(x, W, b, dilation=1):
    """The reference's pipeline, step by step (masked.py:160-232):
    time_to_batch -> left pad k-1 -> VALID conv -> bias -> batch_to_time."""
    k = W.shape[1]
    assert x.shape[1] % dilation == 0  # masked.py:188
    xt = time_to_batch(x, dilation)
    if k > 1:
        xt = np.concatenate(
            [np.zeros((xt.shape[0], k - 1, xt.shape[2]), x.dtype), xt], axis=1)
    n_out = xt.shape[1] - (k - 1)
    y = np.zeros((xt.shape[0], n_out, W.shape[3]), x.dtype)
    for j in range(k):
        y += xt[:, j:j + n_out] @ W[0, j].astype(x.dtype)
    y += b.astype(x.dtype)
    return batch_to_time(y, dilation)


def conv1d(x, W, b, dilation=1):
    """Closed form of masked.conv1d (masked.py:160-232):
    y[b,t,o] = bias[o] + sum_j sum_c W[0,j,c,o] * x[b, t-(k-1-j)*d, c], x[t<0]=0."""
    k = W.shape[1]
    B, T, _ = x.shape
    y = np.zeros((B, T, W.shape[3]), x.dtype)
    for j in range(k):
        sh = (k - 1 - j) * dilation
        if sh >= T:
            continue
        if sh == 0:
            y += x @ W[0, j].astype(x.dtype)
        else:
            y[:, sh:] += x[:, :T - sh] @ W[0, j].astype(x.dtype)
    return y + b.astype(x.dtype)


def trans_conv1d(x, K, b, stride, act=None):
    """masked.py:235-291 (tf.nn.conv2d_transpose, SAME, stride s):
    y[b, i*s + j - p, co] += x[b,i,ci] * K[0,j,co,ci], p=(k-s)//2; + bias; act."""
    B, L, cin = x.shape
    k, cout = K.shape[1], K.shape[2]
    p = (k - stride) // 2
    full = np.zeros((B, L * stride + k, cout), x.dtype)  # index i*s + j
    for j in range(k):
        full[:, j:j + L * stride:stride] += x @ K[0, j].astype(x.dtype).T
    y = full[:, p:p + L * stride] + b.astype(x.dtype)
    if act is not None:
        y = act(y)
    return y


def resize_conv1d(x, W, b, stride, act=None):
    """masked.resize_conv1d (masked.py:294-322): tf.image.resize_nearest_neighbor to length L*stride
    (x_up[u] = x[u // stride]), then masked.conv1d(causal=False): tf.nn.conv2d with SAME padding, stride 1, i.e.
    y[o] = b + sum_j x_up[o + j - (k-1)//2] W[0,j] with zeros outside; then the activation."""
    B, L, cin = x.shape
    k = W.shape[1]
    up = np.repeat(x, stride, axis=1)
    pl = (k - 1) // 2
    padded = np.concatenate([np.zeros((B, pl, cin), x.dtype), up, np.zeros((B, k - 1 - pl, cin), x.dtype)], axis=1)
    y = np.zeros((B, L * stride, W.shape[3]), x.dtype)
    for j in range(k):
        y += padded[:, j:j + L * stride] @ W[0, j].astype(x.dtype)
    y = y + b.astype(x.dtype)
    return act(y) if act is not None else y


def deconv_stack(mel, w, hp, prefix='', dtype=np.float32):
    """wavenet._deconv_stack (wavenet.py:46-73): trans_conv1d layers, or resize_conv1d with use_resize_conv."""
    act = upsample_act(_get(hp, 'upsample_act', 'tanh'))
    x = mel.astype(dtype)
    for i, (fl, s) in enumerate(hp.deconv_config):
        if _get(hp, 'use_resize_conv', False):
            base = '{}resize_conv_{:d}'.format(prefix, i + 1)
            x = resize_conv1d(x, w[base + '/W'], w[base + '/biases'], s, act)
        else:
            base = '{}trans_conv_{:d}'.format(prefix, i + 1)
            x = trans_conv1d(x, w[base + '/kernel'], w[base + '/bias'], s, act)
    return x


def condition(x, cond):
    """wavenet._condition (wavenet.py:76-85): centre-trim cond, add."""
    tl = cond.shape[1] - x.shape[1]
    assert tl >= 0
    left = tl // 2
    return x + cond[:, left:left + x.shape[1]]


def sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def softplus(v):
    return np.logaddexp(v.dtype.type(0), v)


# --------------------------------------------------------------------------
# Student IAF (wavenet/parallel_wavenet.py)
# --------------------------------------------------------------------------
def scale_log_scale_fn(scale_params):
    """PWNHelper.scale_log_scale_fn, USE_LOG_SCALE=False branch
    (parallel_wavenet.py:105-114; numpy twin at tests/test_scale.py:67-78)."""
    dt = scale_params.dtype.type
    sp = softplus(scale_params)
    scale = np.clip(sp, dt(math.exp(-9.0)), dt(math.exp(7.0)))
    return scale, np.log(scale)


def iaf_flow(x, mel_en, w, hp, iaf_idx, dtype=np.float32, taps=None):
    """ParallelWavenet._create_iaf (parallel_wavenet.py:200-287).
    x: [B,T,1]; mel_en: [B,Lc,D].  ``taps`` (optional dict) receives intermediate
    tensors for per-kernel parity tests."""
    p = 'iaf_{:d}'.format(iaf_idx + 1)
    nl = hp.num_iaf_layers[iaf_idx]
    l = conv1d(shift_right(x), w[p + '/start_conv/W'], w[p + '/start_conv/biases'])
    if taps is not None:
        taps['{}/l0'.format(p)] = l.copy()
    for i in range(nl):
        d = 2 ** (i % hp.num_stages)
        dd = conv1d(l, w['{}/dilated_conv_{:d}/W'.format(p, i + 1)],
                    w['{}/dilated_conv_{:d}/biases'.format(p, i + 1)], d)
        c = conv1d(mel_en, w['{}/mel_cond_{:d}/W'.format(p, i + 1)],
                   w['{}/mel_cond_{:d}/biases'.format(p, i + 1)])
        dd = condition(dd, c)
        m = dd.shape[2] // 2
        g = sigmoid(dd[:, :, :m]) * np.tanh(dd[:, :, m:])
        l = l + conv1d(g, w['{}/res_{:d}/W'.format(p, i + 1)],
                       w['{}/res_{:d}/biases'.format(p, i + 1)])
        if taps is not None:
            taps['{}/l{}'.format(p, i + 1)] = l.copy()
    l = np.maximum(l, 0)
    l = conv1d(l, w[p + '/out1/W'], w[p + '/out1/biases'])
    c = conv1d(mel_en, w[p + '/mel_cond_out1/W'], w[p + '/mel_cond_out1/biases'])
    l = np.maximum(condition(l, c), 0)
    mean = conv1d(l, w[p + '/out2_mean/W'], w[p + '/out2_mean/biases'])
    sp = conv1d(l, w[p + '/out2_scale/W'], w[p + '/out2_scale/biases'])
    scale, log_scale = scale_log_scale_fn(sp)
    return {'x': x * scale + mean, 'mean': mean, 'scale': scale, 'log_scale': log_scale}


def iaf_length(num_frames, hp):
    """parallel_wavenet.py:294-302."""
    frame_shift = int(np.prod([dc[1] for dc in hp.deconv_config]))
    max_dil = 2 ** (hp.num_stages - 1)
    return (num_frames * frame_shift // max_dil) * max_dil


def logistic_from_uniform(u):
    """parallel_wavenet.py:173-178."""
    return np.log(u) - np.log(1.0 - u)


def student_feed_forward(w, hp, mel, z, dtype=np.float32, taps=None):
    """ParallelWavenet.feed_forward (parallel_wavenet.py:289-345) with the noise
    ``z`` [B,T] given explicitly (TF's Philox stream is not reproducible here)."""
    mel = np.asarray(mel, dtype)
    z = np.asarray(z, dtype)
    B, F, _ = mel.shape
    T = iaf_length(F, hp)
    assert z.shape == (B, T)
    share = _get(hp, 'use_share_deconv', False) or _get(hp, 'use_teacher_deconv', False)
    mel_en = deconv_stack(mel, w, hp, 'iaf_share/', dtype) if share else None
    x = z[:, :, None]
    mean_tot = np.zeros_like(x)
    scale_tot = np.ones_like(x)
    log_scale_tot = np.zeros_like(x)
    for f in range(len(hp.num_iaf_layers)):
        me = mel_en if share else deconv_stack(mel, w, hp, 'iaf_{:d}/'.format(f + 1), dtype)
        if taps is not None:
            taps['mel_en_{}'.format(f)] = me
        fd = iaf_flow(x, me, w, hp, f, dtype, taps)
        x = fd['x']
        mean_tot = fd['mean'] + mean_tot * fd['scale']
        scale_tot = scale_tot * fd['scale']
        log_scale_tot = log_scale_tot + fd['log_scale']
    dt = np.dtype(dtype).type
    mean_tot = mean_tot[:, :, 0]
    scale_tot = np.minimum(scale_tot, dt(math.exp(7.0)))[:, :, 0]
    log_scale_tot = np.minimum(log_scale_tot, dt(7.0))[:, :, 0]
    new_x = z * scale_tot + mean_tot
    return {'x': new_x, 'mean_tot': mean_tot, 'scale_tot': scale_tot,
            'log_scale_tot': log_scale_tot, 'rand_input': z}


# --------------------------------------------------------------------------
# signal utils (auxilaries/utils.py)
# --------------------------------------------------------------------------
def mu_law(x, mu=255):
    """utils.py:72-87."""
    out = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    return np.floor(out * 128)


def inv_mu_law(x, mu=255):
    """utils.py:108-122 / :125-139."""
    x = np.asarray(x, np.float32)
    out = (x + 0.5) * 2. / (mu + 1)
    out = np.sign(out) / mu * ((1 + mu) ** np.abs(out) - 1)
    return np.where(x == 0, x, out).astype(np.float32)


def cast_quantize(x, quant_chann):
    """utils.py:142-154: floor(x * Q / 2) -> int32."""
    return np.floor(x * quant_chann / 2).astype(np.int32)


def inv_cast_quantize(xq, quant_chann):
    """utils.py:157-159 / :167-169."""
    return xq.astype(np.float32) / np.float32(quant_chann / 2)


def clip_quant_scale(x, quant_chann, use_mu_law):
    """ParallelWavenet._clip_quant_scale (parallel_wavenet.py:348-359;
    restated by the reference at tests/test_clip_quant_scale.py:7-18)."""
    x = np.clip(np.asarray(x, np.float32), np.float32(-1.0),
                np.float32(1.0 - 2.0 / quant_chann))
    xq = cast_quantize(x, quant_chann)
    if use_mu_law:
        return inv_mu_law(xq)
    return inv_cast_quantize(xq, quant_chann)


def parallelgen_forward(w, hp, mel, z, dtype=np.float32):
    """parallelgen.load_parallelgen (parallelgen.py:11-19): feed_forward, then
    fg_dict['x'] = _clip_quant_scale(x)."""
    out = student_feed_forward(w, hp, mel, z, dtype)
    out['x_pre_quant'] = out['x']
    out['x'] = clip_quant_scale(out['x'], quant_chann_of(hp), hp.use_mu_law)
    return out


# --------------------------------------------------------------------------
# Teacher WaveNet, full sequence (wavenet/wavenet.py:180-291)
# --------------------------------------------------------------------------
def teacher_feed_forward(w, hp, wav_scaled, mel, dtype=np.float32, mel_en=None):
    wav_scaled = np.asarray(wav_scaled, dtype)
    if mel_en is None:
        mel_en = deconv_stack(np.asarray(mel, dtype), w, hp, '', dtype)
    x = wav_scaled[:, :, None]
    l = conv1d(shift_right(x), w['conv_start/W'], w['conv_start/biases'])
    s = conv1d(l, w['skip_start/W'], w['skip_start/biases'])
    for i in range(hp.num_layers):
        d = 2 ** (i % hp.num_stages)
        dd = conv1d(l, w['dilated_conv_%d/W' % (i + 1)], w['dilated_conv_%d/biases' % (i + 1)], d)
        c = conv1d(mel_en, w['mel_cond_%d/W' % (i + 1)], w['mel_cond_%d/biases' % (i + 1)])
        dd = condition(dd, c)
        m = dd.shape[2] // 2
        g = sigmoid(dd[:, :, :m]) * np.tanh(dd[:, :, m:])
        l = l + conv1d(g, w['res_%d/W' % (i + 1)], w['res_%d/biases' % (i + 1)])
        s = s + conv1d(g, w['skip_%d/W' % (i + 1)], w['skip_%d/biases' % (i + 1)])
    s = np.maximum(s, 0)
    s = conv1d(s, w['out1/W'], w['out1/biases'])
    c = conv1d(mel_en, w['mel_cond_out1/W'], w['mel_cond_out1/biases'])
    s = np.maximum(condition(s, c), 0)
    out = conv1d(s, w['out2/W'], w['out2/biases'])
    return {'encoding': mel_en, 'out_params': out}


# --------------------------------------------------------------------------
# Output distributions (wavenet/loss_func.py)
# --------------------------------------------------------------------------
def mol_sample(out, quant_chann, u1, u2):
    """loss_func.mol_sample (loss_func.py:154-186) with the two uniform draws
    given: out [B,3*nr_mix] (logits|means|log_scales), u1 [B,nr_mix], u2 [B].
    Returns int32 [B] in [-Q/2, Q/2)."""
    nr = out.shape[-1] // 3
    logit, means, ls = out[..., :nr], out[..., nr:2 * nr], out[..., 2 * nr:]
    sel = np.argmax(logit - np.log(-np.log(u1)), axis=-1)
    idx = np.arange(out.shape[0])
    mu = means[idx, sel]
    log_s = np.clip(ls[idx, sel], -7.0, 7.0)
    x = mu + np.exp(log_s) * (np.log(u2) - np.log(1.0 - u2))
    x = np.clip(x, -1.0, 1.0 - 2.0 / quant_chann)
    return cast_quantize(x, quant_chann)


def gauss_sample(out, quant_chann, n):
    """loss_func.gauss_sample (loss_func.py:200-206), mean_std_from_out_params :66-75
    with use_log_scales=True; n ~ N(0,1) given."""
    mean, lp = out[..., 0], out[..., 1]
    std = np.exp(np.maximum(lp, -7.0))
    x = np.clip(mean + std * n, -1.0, 1.0 - 2.0 / quant_chann)
    return cast_quantize(x, quant_chann)


def log_prob_from_logits(x):
    """loss_func.py:7-11."""
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def log_sum_exp(x):
    """loss_func.py:14-19."""
    m = x.max(axis=-1)
    return m + np.log(np.exp(x - m[..., None]).sum(axis=-1))


def mol_log_probs(mol_params, targets, quant_chann):
    """loss_func.mol_log_probs (loss_func.py:22-63), use_log_scales=True."""
    nr = mol_params.shape[-1] // 3
    logit = mol_params[..., :nr]
    means = mol_params[..., nr:2 * nr]
    log_scales = np.maximum(mol_params[..., 2 * nr:], -7.0)
    inv_stdv = np.exp(-log_scales)
    t = targets[..., None] + np.zeros((1, 1, nr), mol_params.dtype)
    cx = t - means
    plus_in = inv_stdv * (cx + 1. / quant_chann)
    min_in = inv_stdv * (cx - 1. / quant_chann)
    cdf_plus = sigmoid(plus_in)
    cdf_min = sigmoid(min_in)
    log_cdf_plus = plus_in - softplus(plus_in)
    log_one_minus_cdf_min = -softplus(min_in)
    cdf_delta = cdf_plus - cdf_min
    max_val = float(quant_chann - 1)
    max_thres = (max_val - 0.5) / (quant_chann / 2.) - 1.0
    min_thres = 0.5 / (quant_chann / 2.) - 1.0
    lp = np.where(t < min_thres, log_cdf_plus,
                  np.where(t > max_thres, log_one_minus_cdf_min,
                           np.log(np.maximum(cdf_delta, 1e-12))))
    lp = lp + log_prob_from_logits(logit)
    return log_sum_exp(lp)


def mol_loss(mol_params, targets, quant_chann):
    """loss_func.py:117-119."""
    return -mol_log_probs(mol_params, targets, quant_chann).mean()


def ce_sample(out, quant_chann, u):
    """loss_func.ce_sample (loss_func.py:140-151): one draw from Categorical(logits=out), shifted to
    [-Q/2, Q/2).  The reference draws through tf.distributions.Categorical -> tf.multinomial, whose noise
    stream cannot be reproduced outside TensorFlow; with the uniform u [B] given, the draw is defined here by
    the inverse CDF, k = min{k : sum_{i<=k} p_i > u * sum_i p_i} with p = exp(out - max(out)), which has the
    same distribution.  out [B, Q] -> int32 [B]."""
    out = np.asarray(out, np.float64)
    p = np.exp(out - out.max(axis=-1, keepdims=True))
    cdf = np.cumsum(p, axis=-1)
    thr = np.asarray(u, np.float64) * cdf[..., -1]
    k = (cdf > thr[..., None]).argmax(axis=-1)
    k = np.where(cdf[..., -1] > thr, k, out.shape[-1] - 1)
    return (k - quant_chann // 2).astype(np.int32)


# --------------------------------------------------------------------------
# Autoregressive fastgen (wavenet/wavenet.py:379-514, masked.py:328-405,
# wavenet/fastgen.py:128-169)
# --------------------------------------------------------------------------
class _Queue:
    """tf.FIFOQueue of depth ``rate`` initialised with zeros (masked.py:352-359)."""

    def __init__(self, rate, shape, dtype):
        self.buf = [np.zeros(shape, dtype) for _ in range(rate)]

    def dequeue(self):
        return self.buf.pop(0)

    def enqueue(self, v):
        self.buf.append(v)


class FastgenOracle:
    """One-timestep graph with queues (Fastgen.sample, wavenet.py:379-514).
    step(x_t, enc_t) returns the pre-sample tensor out[B,out_width]."""

    def __init__(self, w, hp, batch_size, dtype=np.float32):
        assert hp.filter_length == 3  # masked.py:349
        self.w = {k: v.astype(dtype) for k, v in w.items()}
        self.hp = hp
        self.dtype = dtype
        self.B = batch_size
        width = hp.width
        self.q = {}
        self.q['conv_start'] = (_Queue(1, (batch_size, 1), dtype), _Queue(1, (batch_size, 1), dtype))
        for i in range(hp.num_layers):
            rate = 2 ** (i % hp.num_stages)
            self.q['dilated_conv_%d' % (i + 1)] = (
                _Queue(rate, (batch_size, width), dtype), _Queue(rate, (batch_size, width), dtype))

    def _causal_linear(self, x, name):
        """masked.causal_linear (masked.py:328-380): W[:,0] hits the state from
        2*rate steps ago, W[:,1] from rate steps ago, W[:,2] the current input."""
        q1, q2 = self.q[name]
        s1 = q1.dequeue()
        q1.enqueue(x)
        s2 = q2.dequeue()
        q2.enqueue(s1)
        W = self.w[name + '/W']
        return s2 @ W[0, 0] + s1 @ W[0, 1] + x @ W[0, 2] + self.w[name + '/biases']

    def _linear(self, x, name):
        """masked.linear (masked.py:383-405)."""
        return x @ self.w[name + '/W'][0, 0] + self.w[name + '/biases']

    def step(self, wav, enc):
        hp = self.hp
        x = np.asarray(wav, self.dtype).reshape(self.B, 1)
        enc = np.asarray(enc, self.dtype)
        if hp.use_mu_law:
            x = (mu_law(x) / (quant_chann_of(hp) / 2)).astype(self.dtype)  # wavenet.py:411-414
        l = self._causal_linear(x, 'conv_start')
        s = self._linear(l, 'skip_start')
        for i in range(hp.num_layers):
            d = self._causal_linear(l, 'dilated_conv_%d' % (i + 1))
            d = d + self._linear(enc, 'mel_cond_%d' % (i + 1))
            m = d.shape[1] // 2
            g = sigmoid(d[:, :m]) * np.tanh(d[:, m:])
            l = l + self._linear(g, 'res_%d' % (i + 1))
            s = s + self._linear(g, 'skip_%d' % (i + 1))
        s = np.maximum(s, 0)
        s = self._linear(s, 'out1') + self._linear(enc, 'mel_cond_out1')
        s = np.maximum(s, 0)
        return self._linear(s, 'out2')


def fastgen_run(w, hp, encoding, dtype=np.float32, teacher_force=None,
                u1=None, u2=None, n=None, return_out=True):
    """fastgen.synthesis loop (fastgen.py:147-168): zero-initialised queues,
    audio starts at 0, encoding[:, i] fed with NO centre trim, previous sample fed
    back dequantised (inv_cast_quantize_numpy, utils.py:167).

    teacher_force [B,T]: if given, wav fed at step i is teacher_force[:, i-1]
    (0 at i=0) instead of the model's own sample.  Noise (u1 [B,T,nr_mix], u2 [B,T]
    for MoL; n [B,T] for gauss) must be given when free-running."""
    encoding = np.asarray(encoding, dtype)
    B, T, _ = encoding.shape
    Q = quant_chann_of(hp)
    fg = FastgenOracle(w, hp, B, dtype)
    audio = np.zeros((B, 1), dtype)
    outs = []
    samples = np.zeros((B, T), np.float32)
    for i in range(T):
        out = fg.step(audio, encoding[:, i])
        if return_out:
            outs.append(out)
        if teacher_force is not None:
            audio = np.asarray(teacher_force[:, i:i + 1], dtype)
            samples[:, i] = audio[:, 0]
            continue
        if hp.loss_type == 'mol':
            q = mol_sample(out.astype(np.float32), Q, u1[:, i], u2[:, i])
        elif hp.loss_type == 'gauss':
            q = gauss_sample(out.astype(np.float32), Q, n[:, i])
        else:
            q = ce_sample(out.astype(np.float32), Q, n[:, i])   # n carries the uniform of the inverse-CDF draw
        a = inv_mu_law(q) if hp.use_mu_law else inv_cast_quantize(q, Q)
        audio = a.reshape(B, 1).astype(dtype)
        samples[:, i] = a
    res = {'audio': samples}
    if return_out:
        res['out'] = np.stack(outs, axis=1)
    return res


def load_hparams(path):
    import json
    with open(path, 'rt') as f:
        return Namespace(**json.load(f))


# --------------------------------------------------------------------------
# Distillation cross-entropy (wavenet/parallel_wavenet.py:361-402)
# --------------------------------------------------------------------------
def kl_loss_logistic(te_out, mean_tot, scale_tot, log_scale_tot, eps, quant_chann):
    """ParallelWavenet.kl_loss_logistic with the teacher output `te_out` [B,T,3*nr] and the
    logistic draws `eps` [S,B,T] given (the reference tiles with tf_repeat along the batch axis,
    utils.py:175-195, which is np.repeat; averaging over the S copies is order-independent).
    CLIP=False (parallel_wavenet.py:15), so x_xp is not clipped."""
    S = eps.shape[0]
    x_xp = eps * scale_tot[None] + mean_tot[None]                       # :373-377
    te = np.broadcast_to(te_out[None], (S,) + te_out.shape)
    B, T = mean_tot.shape
    lp = mol_log_probs(te.reshape(S * B, T, -1), x_xp.reshape(S * B, T), quant_chann)   # :385-390
    H_Ps_Pt = float(-lp.mean())                                         # :392-396
    H_Ps = float(log_scale_tot.mean() + 2.0)                            # :395
    return {'H_Ps': H_Ps, 'H_Ps_Pt': H_Ps_Pt, 'kl_loss': H_Ps_Pt - H_Ps}


def mean_std_from_out_params(gauss_params, use_log_scales=True):
    """loss_func.mean_std_from_out_params (wavenet/loss_func.py:66-75): [B,T,2] -> mean, std."""
    mean, std_param = gauss_params[..., 0], gauss_params[..., 1]
    if use_log_scales:
        std = np.exp(np.maximum(std_param, gauss_params.dtype.type(-7.0)))      # :71-72
    else:
        sp = np.logaddexp(std_param, 0)                                        # softplus, :74
        std = np.maximum(sp, np.exp(gauss_params.dtype.type(-7.0)))
    return mean, std


def kl_loss_gauss(te_out, mean_tot, scale_tot, log_scale_tot):
    """ParallelWavenet.kl_loss_gauss (wavenet/parallel_wavenet.py:404-428) with the Gaussian teacher's
    output `te_out` [B,T,2] given: closed-form KL(q || p) per sample between the student's N(mean_tot,
    scale_tot) and the teacher's N(mean_p, scale_p), averaged, plus 4 x the mean squared difference of the
    log scales (ClariNet's regulariser).  Arithmetic in the dtype of the inputs, like the TF graph."""
    mean_p, scale_p = mean_std_from_out_params(te_out, use_log_scales=True)     # :416-417
    log_scale_p = np.log(scale_p)                                               # :418
    var_q = scale_tot ** 2                                                      # :420
    var_p = scale_p ** 2                                                        # :421
    kl_bl = (log_scale_p - log_scale_tot +
             (var_q - var_p + (mean_p - mean_tot) ** 2) / (2 * var_p))          # :422-423
    kl = float(kl_bl.mean(dtype=np.float64))                                    # :424
    reg = float(((log_scale_p - log_scale_tot) ** 2).mean(dtype=np.float64))    # :425
    return {'kl': kl, 'reg': reg, 'kl_loss': kl + 4.0 * reg}                    # :426-428
