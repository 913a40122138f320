"""Thin Python handles over the C ABI (include/nsw.h).

`IAFEngine` stands where the reference builds a TF graph for
ParallelWavenet.feed_forward (wavenet/parallelgen.py:11-19) and `FastgenEngine`
where it builds Fastgen.sample + the deconv stack (wavenet/fastgen.py:61-66,118-125).
Weights are dicts keyed by the reference's TF variable names.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


def _get(hp, name, default):
    return getattr(hp, name, default)


def default_engine():
    # tc3 = conv-GEMMs and the IAF residual layers on tcgen05 (fp32-grade via split fp16: hi + lo, three products)
    return os.environ.get('NSW_ENGINE', 'tc3')


def iaf_config(hp, num_mel=80, engine=None):
    cfg = L.nsw_iaf_config()
    layers = list(hp.num_iaf_layers)
    if len(layers) > L.NSW_MAX_FLOWS:
        raise ValueError('too many IAF flows')
    cfg.num_flows = len(layers)
    for i, n in enumerate(layers):
        cfg.num_iaf_layers[i] = n
    cfg.num_stages = hp.num_stages
    cfg.filter_length = hp.filter_length
    cfg.width = hp.width
    cfg.deconv_width = hp.deconv_width
    cfg.num_mel = num_mel
    dc = list(hp.deconv_config)
    cfg.num_deconv = len(dc)
    for i, (fl, s) in enumerate(dc):
        cfg.deconv_filter[i] = fl
        cfg.deconv_stride[i] = s
    # parallel_wavenet.py:130-135
    cfg.share_deconv = int(bool(_get(hp, 'use_share_deconv', False) or
                                _get(hp, 'use_teacher_deconv', False)))
    cfg.loss_type = L.LOSS[_get(hp, 'loss_type', 'logistic')]  # parallel_wavenet.py:129
    cfg.upsample_act = L.ACT[_get(hp, 'upsample_act', 'tanh')]
    cfg.use_mu_law = int(bool(hp.use_mu_law))
    cfg.engine = L.ENGINE[engine or default_engine()]
    # use_resize_conv (masked.py:294-322) needs no flag: the library picks the upsampler from the variable names it is
    # given (resize_conv_i/{W,biases} or trans_conv_i/{kernel,bias})
    return cfg


def wavenet_config(hp, num_mel=80, engine=None):
    cfg = L.nsw_wavenet_config()
    cfg.num_layers = hp.num_layers
    cfg.num_stages = hp.num_stages
    cfg.filter_length = hp.filter_length
    cfg.width = hp.width
    # wavenet.py:106,204
    cfg.gate_width = 2 * hp.width if _get(hp, 'double_gate_width', True) else hp.width
    cfg.skip_width = hp.skip_width
    use_mu_law = bool(hp.use_mu_law)
    qc = 2 ** 8 if use_mu_law else 2 ** 16
    cfg.out_width = {'ce': qc, 'mol': 3 * _get(hp, 'mol_mix', 10), 'gauss': 2}[hp.loss_type]
    cfg.deconv_width = hp.deconv_width
    cfg.num_mel = num_mel
    dc = list(hp.deconv_config)
    cfg.num_deconv = len(dc)
    for i, (fl, s) in enumerate(dc):
        cfg.deconv_filter[i] = fl
        cfg.deconv_stride[i] = s
    cfg.loss_type = L.LOSS[hp.loss_type]
    cfg.upsample_act = L.ACT[_get(hp, 'upsample_act', 'tanh')]
    cfg.use_mu_law = int(use_mu_law)
    cfg.engine = L.ENGINE[engine or default_engine()]
    return cfg


def fold_weight_norm(weights):
    """masked.get_kernel (masked.py:131-157): W = g * V / ||V||.  Returns a dict
    with every (X_V, X_g) pair replaced by X so the engines only see plain kernels."""
    out = {}
    for name, v in weights.items():
        if name.endswith('_V'):
            base = name[:-2]
            g = np.asarray(weights[base + '_g'], np.float64)
            v64 = np.asarray(v, np.float64)
            if base.endswith('/kernel'):  # deconv [1,k,Cout,Cin]: norm over (0,1,3)
                nrm = np.sqrt((v64 ** 2).sum(axis=(0, 1, 3), keepdims=True))
                out[base] = (v64 / nrm * g.reshape(1, 1, -1, 1)).astype(np.float32)
            else:                          # conv [1,k,Cin,Cout]: norm over (0,1,2)
                nrm = np.sqrt((v64 ** 2).sum(axis=(0, 1, 2), keepdims=True))
                out[base] = (v64 / nrm * g.reshape(1, 1, 1, -1)).astype(np.float32)
        elif not name.endswith('_g'):
            out[name] = v
    return out


def _dev_tensor(x, name, shape, device):
    """The C ABI takes raw device pointers: reject anything whose memory is not what the kernels index
    (CUDA, float32, contiguous, on the handle's device, expected shape; None in `shape` = any extent)."""
    import torch
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        raise TypeError('{}: expected a torch CUDA tensor, got {}'.format(name, type(x).__name__))
    if not x.is_cuda:
        raise ValueError('{}: must live on the GPU (got a {} tensor)'.format(name, x.device))
    if x.device.index != device:
        raise ValueError('{}: lives on cuda:{} but the engine was created on cuda:{}'.format(
            name, x.device.index, device))
    if x.dtype != torch.float32:
        raise ValueError('{}: must be float32 (got {})'.format(name, x.dtype))
    if not x.is_contiguous():
        raise ValueError('{}: must be contiguous (call .contiguous())'.format(name))
    if shape is not None:
        if x.dim() != len(shape) or any(e is not None and int(d) != int(e) for d, e in zip(x.shape, shape)):
            raise ValueError('{}: shape {} does not match {}'.format(name, tuple(x.shape), tuple(shape)))
    return x


class IAFEngine:
    """4-flow IAF student on one GPU."""

    def __init__(self, hparams, weights, device=0, num_mel=80, engine=None):
        self.lib = L.load()
        self.hparams = hparams
        self.device = device
        self.num_mel = num_mel
        self.engine = engine or default_engine()
        cfg = iaf_config(hparams, num_mel, self.engine)
        tensors, keep = L.make_tensors(fold_weight_norm(weights))
        h = C.c_void_p()
        L.check(self.lib.nsw_iaf_create(C.byref(cfg), tensors, len(tensors), device, C.byref(h)))
        del keep
        self._h = h
        self.quant_chann = 2 ** 8 if hparams.use_mu_law else 2 ** 16

    def close(self):
        if getattr(self, '_h', None):
            self.lib.nsw_iaf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def length(self, num_frames):
        return int(self.lib.nsw_iaf_length(self._h, num_frames))

    def forward_host(self, mel, z=None, seed=0, quantize=True, want=('x',)):
        """mel np[B,F,num_mel] (+ optional noise z np[B,T]) -> dict of np[B,T].
        H2D and D2H copies happen inside the C call."""
        mel = np.ascontiguousarray(mel, np.float32)
        B, F, M = mel.shape
        assert M == self.num_mel
        T = self.length(F)
        if z is not None:
            z = np.ascontiguousarray(z, np.float32)
            assert z.shape == (B, T), (z.shape, (B, T))
        names = ('x', 'mean_tot', 'scale_tot', 'log_scale_tot', 'rand_input')
        out = {n: (np.empty((B, T), np.float32) if n in want else None) for n in names}
        L.check(self.lib.nsw_iaf_forward_host(
            self._h, L.ptr(mel), L.ptr(z), seed, B, F, int(bool(quantize)),
            *[L.ptr(out[n]) for n in names]))
        return {n: v for n, v in out.items() if v is not None}

    def forward_device(self, mel, z=None, seed=0, quantize=True, out=None, stream=None):
        """torch CUDA tensors in, torch CUDA tensors out; enqueues on the current
        torch stream (or `stream`) and does not synchronise."""
        import torch
        _dev_tensor(mel, 'mel', (None, None, self.num_mel), self.device)
        B, F, _ = mel.shape
        T = self.length(F)
        names = ('x', 'mean_tot', 'scale_tot', 'log_scale_tot', 'rand_input')
        if out is None:
            out = {n: torch.empty((B, T), dtype=torch.float32, device=mel.device)
                   for n in names[:4]}
        for n, v in out.items():
            _dev_tensor(v, 'out[{!r}]'.format(n), (B, T), self.device)
        _dev_tensor(z, 'z', (B, T), self.device)
        st = stream if stream is not None else torch.cuda.current_stream(mel.device).cuda_stream
        L.check(self.lib.nsw_iaf_forward_device(
            self._h, L.ptr(mel), L.ptr(z), seed, B, F, int(bool(quantize)),
            *[L.ptr(out.get(n)) for n in names], st))
        return out

    def deconv_device(self, mel, stack=0):
        import torch
        _dev_tensor(mel, 'mel', (None, None, self.num_mel), self.device)
        B, F, _ = mel.shape
        stride = int(np.prod([dc[1] for dc in self.hparams.deconv_config]))
        enc = torch.empty((B, F * stride, self.hparams.deconv_width), dtype=torch.float32,
                          device=mel.device)
        st = torch.cuda.current_stream(mel.device).cuda_stream
        L.check(self.lib.nsw_iaf_deconv_device(self._h, stack, L.ptr(mel), B, F, L.ptr(enc), st))
        return enc

    def set_tap(self, flow, layer, dst):
        _dev_tensor(dst, 'dst', None, self.device)
        L.check(self.lib.nsw_iaf_set_tap(self._h, flow, layer, L.ptr(dst)))

    def set_profiling(self, on):
        L.check(self.lib.nsw_iaf_set_profiling(self._h, int(on)))

    def last_timing(self):
        ms = (C.c_float * 5)()
        L.check(self.lib.nsw_iaf_last_timing(self._h, C.byref(ms)))
        return dict(zip(('deconv', 'cond', 'layers', 'heads', 'total'), [float(v) for v in ms]))

    def workspace_bytes(self):
        return int(self.lib.nsw_iaf_workspace_bytes(self._h))


class FastgenEngine:
    """Teacher WaveNet: deconv encoder + persistent autoregressive kernel."""

    def __init__(self, hparams, weights, device=0, num_mel=80, engine=None):
        self.lib = L.load()
        self.hparams = hparams
        self.device = device
        self.num_mel = num_mel
        cfg = wavenet_config(hparams, num_mel, engine)
        self.out_width = cfg.out_width
        self.gate_width = cfg.gate_width
        tensors, keep = L.make_tensors(fold_weight_norm(weights))
        h = C.c_void_p()
        L.check(self.lib.nsw_fastgen_create(C.byref(cfg), tensors, len(tensors), device,
                                            C.byref(h)))
        del keep
        self._h = h
        self.frame_shift = int(np.prod([dc[1] for dc in hparams.deconv_config]))

    def close(self):
        if getattr(self, '_h', None):
            self.lib.nsw_fastgen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_host(self, mel):
        mel = np.ascontiguousarray(mel, np.float32)
        B, F, _ = mel.shape
        enc = np.empty((B, F * self.frame_shift, self.hparams.deconv_width), np.float32)
        L.check(self.lib.nsw_fastgen_encode_host(self._h, L.ptr(mel), B, F, L.ptr(enc)))
        return enc

    def cond_vars_host(self, encoding):
        """Fastgen.cond_vars (wavenet.py:353-377): dict layer name -> np[B, T, width] of the mel-conditioning
        projections (mel_cond_1..L: gate_width columns each, mel_cond_out1: skip_width)."""
        encoding = np.ascontiguousarray(encoding, np.float32)
        B, T, _ = encoding.shape
        Lr, G, S = self.hparams.num_layers, self.gate_width, self.hparams.skip_width
        flat = np.empty((B, T, Lr * G + S), np.float32)
        L.check(self.lib.nsw_fastgen_cond_vars_host(self._h, L.ptr(encoding), B, T, L.ptr(flat)))
        out = {'mel_cond_%d' % (i + 1): flat[:, :, i * G:(i + 1) * G] for i in range(Lr)}
        out['mel_cond_out1'] = flat[:, :, Lr * G:]
        return out

    def noise_width(self):
        """floats of sampler noise per step: mol nr_mix + 1 (u1, u2), gauss 1 (n), ce 1 (u)."""
        return self.out_width // 3 + 1 if self.hparams.loss_type == 'mol' else 1

    def set_noise(self, noise):
        """Parity hook (nsw_fastgen_set_noise): the following runs read their random draws from `noise`
        [B, T, noise_width()] (NumPy or torch CUDA) instead of the in-kernel Philox stream; None restores Philox."""
        if noise is None:
            L.check(self.lib.nsw_fastgen_set_noise(self._h, None, 0, 0, 0, 0))
            return
        if isinstance(noise, np.ndarray):
            noise = np.ascontiguousarray(noise, np.float32)
            on_dev = 0
        else:
            _dev_tensor(noise, 'noise', (None, None, self.noise_width()), self.device)
            on_dev = 1
        B, T, nu = noise.shape
        L.check(self.lib.nsw_fastgen_set_noise(self._h, L.ptr(noise), B, T, nu, on_dev))

    def run_host(self, encoding, teacher_force=None, seed=0, want_out=False):
        encoding = np.ascontiguousarray(encoding, np.float32)
        if encoding.ndim != 3 or encoding.shape[2] != self.hparams.deconv_width:
            raise ValueError('encoding must be [B, T, {}], got {}'.format(self.hparams.deconv_width,
                                                                          encoding.shape))
        B, T, _ = encoding.shape
        audio = np.empty((B, T), np.float32)
        out = np.empty((B, T, self.out_width), np.float32) if want_out else None
        if teacher_force is not None:
            teacher_force = np.ascontiguousarray(teacher_force, np.float32)
            assert teacher_force.shape == (B, T)
        L.check(self.lib.nsw_fastgen_run_host(self._h, L.ptr(encoding), B, T,
                                              L.ptr(teacher_force), seed, L.ptr(audio),
                                              L.ptr(out)))
        return (audio, out) if want_out else audio

    def run_device(self, encoding, teacher_force=None, seed=0, want_out=False):
        import torch
        _dev_tensor(encoding, 'encoding', (None, None, self.hparams.deconv_width), self.device)
        B, T, _ = encoding.shape
        _dev_tensor(teacher_force, 'teacher_force', (B, T), self.device)
        audio = torch.empty((B, T), dtype=torch.float32, device=encoding.device)
        out = (torch.empty((B, T, self.out_width), dtype=torch.float32, device=encoding.device)
               if want_out else None)
        st = torch.cuda.current_stream(encoding.device).cuda_stream
        L.check(self.lib.nsw_fastgen_run_device(self._h, L.ptr(encoding), B, T,
                                                L.ptr(teacher_force), seed, L.ptr(audio),
                                                L.ptr(out), st))
        return (audio, out) if want_out else audio

    def last_timing(self):
        ms = C.c_float()
        L.check(self.lib.nsw_fastgen_last_timing(self._h, C.byref(ms)))
        return float(ms.value)


class TeacherEngine:
    """Teacher WaveNet full-sequence forward + student->teacher cross-entropy (distillation
    scoring, BASELINE config 5), all contractions on tcgen05."""

    def __init__(self, hparams, weights, device=0, num_mel=80):
        self.lib = L.load()
        self.hparams = hparams
        self.device = device
        self.num_mel = num_mel
        cfg = wavenet_config(hparams, num_mel, 'tc')
        self.out_width = cfg.out_width
        tensors, keep = L.make_tensors(fold_weight_norm(weights))
        h = C.c_void_p()
        L.check(self.lib.nsw_teacher_create(C.byref(cfg), tensors, len(tensors), device, C.byref(h)))
        del keep
        self._h = h

    def close(self):
        if getattr(self, '_h', None):
            self.lib.nsw_teacher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward_host(self, wav_scaled, mel):
        wav = np.ascontiguousarray(wav_scaled, np.float32)
        mel = np.ascontiguousarray(mel, np.float32)
        B, T = wav.shape
        F = mel.shape[1]
        out = np.empty((B, T, self.out_width), np.float32)
        L.check(self.lib.nsw_teacher_forward_host(self._h, L.ptr(wav), L.ptr(mel), B, T, F, L.ptr(out)))
        return out

    def forward_device(self, wav_scaled, mel):
        import torch
        _dev_tensor(wav_scaled, 'wav_scaled', (None, None), self.device)
        _dev_tensor(mel, 'mel', (wav_scaled.shape[0], None, self.num_mel), self.device)
        B, T = wav_scaled.shape
        F = mel.shape[1]
        out = torch.empty((B, T, self.out_width), dtype=torch.float32, device=wav_scaled.device)
        st = torch.cuda.current_stream(wav_scaled.device).cuda_stream
        L.check(self.lib.nsw_teacher_forward_device(self._h, L.ptr(wav_scaled), L.ptr(mel), B, T, F,
                                                    L.ptr(out), st))
        return out

    def mol_score(self, te_out, mean_tot, scale_tot, log_scale_tot, num_samples=100, eps=None, seed=0):
        """-> dict(H_Ps, H_Ps_Pt, kl_loss) like ParallelWavenet.kl_loss_logistic (torch CUDA inputs)."""
        import torch
        _dev_tensor(mean_tot, 'mean_tot', (None, None), self.device)
        B, T = mean_tot.shape
        _dev_tensor(te_out, 'te_out', (B, T, self.out_width), self.device)
        _dev_tensor(scale_tot, 'scale_tot', (B, T), self.device)
        _dev_tensor(log_scale_tot, 'log_scale_tot', (B, T), self.device)
        _dev_tensor(eps, 'eps', (num_samples, B, T), self.device)
        res = (C.c_double * 3)()
        st = torch.cuda.current_stream(mean_tot.device).cuda_stream
        L.check(self.lib.nsw_mol_score_device(self._h, L.ptr(te_out), L.ptr(mean_tot), L.ptr(scale_tot),
                                              L.ptr(log_scale_tot), L.ptr(eps), seed, num_samples, B, T,
                                              C.byref(res), st))
        return {'H_Ps': res[0], 'H_Ps_Pt': res[1], 'kl_loss': res[2]}

    def gauss_kl(self, te_out, mean_tot, scale_tot, log_scale_tot):
        """-> dict(kl_loss, kl, reg) like ParallelWavenet.kl_loss_gauss (parallel_wavenet.py:404-428) downstream of
        the teacher forward; needs a gauss teacher (wavenet_gauss.json).  torch CUDA inputs."""
        import torch
        _dev_tensor(mean_tot, 'mean_tot', (None, None), self.device)
        B, T = mean_tot.shape
        _dev_tensor(te_out, 'te_out', (B, T, 2), self.device)       # (a mol teacher is refused by the library)
        _dev_tensor(scale_tot, 'scale_tot', (B, T), self.device)
        _dev_tensor(log_scale_tot, 'log_scale_tot', (B, T), self.device)
        res = (C.c_double * 3)()
        st = torch.cuda.current_stream(mean_tot.device).cuda_stream
        L.check(self.lib.nsw_gauss_kl_device(self._h, L.ptr(te_out), L.ptr(mean_tot), L.ptr(scale_tot),
                                             L.ptr(log_scale_tot), B, T, C.byref(res), st))
        return {'kl': res[0], 'reg': res[1], 'kl_loss': res[2]}

    def last_timing(self):
        ms = C.c_float()
        L.check(self.lib.nsw_teacher_last_timing(self._h, C.byref(ms)))
        return float(ms.value)
