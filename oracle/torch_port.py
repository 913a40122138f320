"""Multi-threaded fp32 CPU port of the reference's generation graphs on PyTorch-CPU ops.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/wavenet_oracle.py header): this is the
"reference-equivalent CPU path" timed by bench.py's cpu_baseline leg and by
`bench.py --impl reference`, because the reference's own TensorFlow-1.x graph cannot
run in this image (PARITY UNPINNED; TF absent).  It follows the same file:line map as
wavenet_oracle.py and is checked against it in tests/test_torch_port.py.
Layout here is channels-first [B, C, T] (what oneDNN convs want); weights are
converted once with the rules of SURVEY.md Appendix A."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _conv_w(W):  # TF [1,k,Cin,Cout] -> torch [Cout,Cin,k]
    return torch.from_numpy(np.ascontiguousarray(np.transpose(W[0], (2, 1, 0)))).float()


def _deconv_w(K):  # TF [1,k,Cout,Cin] -> torch conv_transpose1d [Cin,Cout,k]
    return torch.from_numpy(np.ascontiguousarray(np.transpose(K[0], (2, 1, 0)))).float()


def _b(v):
    return torch.from_numpy(np.asarray(v, np.float32))


def causal_conv(x, w, b, d):
    """masked.conv1d (masked.py:160-232): left pad (k-1)*d, dilation d."""
    k = w.shape[-1]
    if k > 1:
        x = F.pad(x, ((k - 1) * d, 0))
    return F.conv1d(x, w, b, dilation=d)


def _act(name):
    return {'tanh': torch.tanh, 'relu': F.relu,
            'leaky_relu': lambda v: F.leaky_relu(v, 0.4)}[name]


class StudentPort:
    """ParallelWavenet.feed_forward + _clip_quant_scale (parallel_wavenet.py:200-359)."""

    def __init__(self, w, hp):
        self.hp = hp
        self.share = bool(getattr(hp, 'use_share_deconv', False) or
                          getattr(hp, 'use_teacher_deconv', False))
        self.act = _act(getattr(hp, 'upsample_act', 'tanh'))
        self.gauss = getattr(hp, 'loss_type', 'logistic') != 'logistic'
        t = {}
        for name, v in w.items():
            if name.endswith('/W'):
                t[name] = _conv_w(v)
            elif name.endswith('/kernel'):
                t[name] = _deconv_w(v)
            else:
                t[name] = _b(v)
        self.t = t

    def deconv(self, mel, prefix):
        x = mel.transpose(1, 2)
        for i, (fl, s) in enumerate(self.hp.deconv_config):
            base = '{}trans_conv_{:d}'.format(prefix, i + 1)
            x = self.act(F.conv_transpose1d(x, self.t[base + '/kernel'], self.t[base + '/bias'],
                                            stride=s, padding=(fl - s) // 2))
        return x

    def flow(self, x, mel_en, f):
        hp, t = self.hp, self.t
        p = 'iaf_{:d}'.format(f + 1)
        T = x.shape[-1]
        left = (mel_en.shape[-1] - T) // 2
        me = mel_en[:, :, left:left + T]  # 1x1 conv commutes with the centre trim
        l = causal_conv(F.pad(x, (1, 0))[:, :, :-1], t[p + '/start_conv/W'], t[p + '/start_conv/biases'], 1)
        for i in range(hp.num_iaf_layers[f]):
            d = 2 ** (i % hp.num_stages)
            dd = causal_conv(l, t['{}/dilated_conv_{:d}/W'.format(p, i + 1)],
                             t['{}/dilated_conv_{:d}/biases'.format(p, i + 1)], d)
            dd = dd + F.conv1d(me, t['{}/mel_cond_{:d}/W'.format(p, i + 1)],
                               t['{}/mel_cond_{:d}/biases'.format(p, i + 1)])
            m = dd.shape[1] // 2
            g = torch.sigmoid(dd[:, :m]) * torch.tanh(dd[:, m:])
            l = l + F.conv1d(g, t['{}/res_{:d}/W'.format(p, i + 1)], t['{}/res_{:d}/biases'.format(p, i + 1)])
        l = F.relu(l)
        l = F.conv1d(l, t[p + '/out1/W'], t[p + '/out1/biases'])
        l = F.relu(l + F.conv1d(me, t[p + '/mel_cond_out1/W'], t[p + '/mel_cond_out1/biases']))
        mean = F.conv1d(l, t[p + '/out2_mean/W'], t[p + '/out2_mean/biases'])
        sp = F.conv1d(l, t[p + '/out2_scale/W'], t[p + '/out2_scale/biases'])
        scale = torch.clamp(F.softplus(sp), math.exp(-9.0), math.exp(7.0))
        return x * scale + mean, mean, scale, torch.log(scale)

    @torch.no_grad()
    def forward(self, mel, z, quantize=True):
        mel = torch.as_tensor(mel, dtype=torch.float32)
        z = torch.as_tensor(z, dtype=torch.float32)
        x = z[:, None, :]
        mel_en = self.deconv(mel, 'iaf_share/') if self.share else None
        mean_tot, scale_tot, ls_tot = 0.0, 1.0, 0.0
        for f in range(len(self.hp.num_iaf_layers)):
            me = mel_en if self.share else self.deconv(mel, 'iaf_{:d}/'.format(f + 1))
            x, mean, scale, ls = self.flow(x, me, f)
            mean_tot = mean + mean_tot * scale
            scale_tot = scale_tot * scale
            ls_tot = ls_tot + ls
        scale_tot = torch.clamp(scale_tot, max=math.exp(7.0))[:, 0]
        ls_tot = torch.clamp(ls_tot, max=7.0)[:, 0]
        mean_tot = mean_tot[:, 0]
        new_x = z * scale_tot + mean_tot
        if quantize:
            Q = 256 if self.hp.use_mu_law else 65536
            new_x = torch.floor(torch.clamp(new_x, -1.0, 1.0 - 2.0 / Q) * (Q / 2)) / (Q / 2)
        return {'x': new_x.numpy(), 'mean_tot': mean_tot.numpy(), 'scale_tot': scale_tot.numpy(),
                'log_scale_tot': ls_tot.numpy()}


class FastgenPort:
    """fastgen.synthesis loop (fastgen.py:147-168) with Fastgen.sample (wavenet.py:379-514):
    one Python iteration per audio sample, like the reference's sess.run loop."""

    def __init__(self, w, hp, batch):
        self.hp = hp
        self.B = batch
        self.w = {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in w.items()}
        self.rates = [2 ** (i % hp.num_stages) for i in range(hp.num_layers)]
        self.reset()

    def reset(self):
        W = self.hp.width
        self.q0 = [torch.zeros(self.B, 1), torch.zeros(self.B, 1)]
        self.q = [[torch.zeros(2 * r, self.B, W), 0] for r in self.rates]  # ring of depth 2*rate

    def _lin(self, x, name):
        return x @ self.w[name + '/W'][0, 0] + self.w[name + '/biases']

    @torch.no_grad()
    def step(self, x, enc):
        w = self.w
        Wc = w['conv_start/W'][0]
        l = self.q0[1] @ Wc[0] + self.q0[0] @ Wc[1] + x @ Wc[2] + w['conv_start/biases']
        self.q0 = [x, self.q0[0]]
        s = self._lin(l, 'skip_start')
        for i, r in enumerate(self.rates):
            ring, pos = self.q[i]
            s1 = ring[(pos - r) % (2 * r)]
            s2 = ring[pos % (2 * r)]  # written 2*r steps ago
            Wd = w['dilated_conv_%d/W' % (i + 1)][0]
            d = s2 @ Wd[0] + s1 @ Wd[1] + l @ Wd[2] + w['dilated_conv_%d/biases' % (i + 1)]
            ring[pos % (2 * r)] = l
            self.q[i][1] = pos + 1
            d = d + self._lin(enc, 'mel_cond_%d' % (i + 1))
            m = d.shape[1] // 2
            g = torch.sigmoid(d[:, :m]) * torch.tanh(d[:, m:])
            l = l + self._lin(g, 'res_%d' % (i + 1))
            s = s + self._lin(g, 'skip_%d' % (i + 1))
        s = F.relu(s)
        s = F.relu(self._lin(s, 'out1') + self._lin(enc, 'mel_cond_out1'))
        return self._lin(s, 'out2')

    @torch.no_grad()
    def run(self, encoding, steps, seed=0):
        """free-running MoL generation for `steps` samples; returns audio [B, steps]."""
        g = torch.Generator().manual_seed(seed)
        enc = torch.as_tensor(encoding, dtype=torch.float32)
        audio = torch.zeros(self.B, 1)
        out_audio = torch.zeros(self.B, steps)
        Q = 256 if self.hp.use_mu_law else 65536
        for i in range(steps):
            out = self.step(audio, enc[:, i])
            nr = out.shape[1] // 3
            u1 = torch.rand(self.B, nr, generator=g) * (1 - 2e-5) + 1e-5
            u2 = torch.rand(self.B, generator=g) * (1 - 2e-5) + 1e-5
            sel = torch.argmax(out[:, :nr] - torch.log(-torch.log(u1)), dim=1)
            idx = torch.arange(self.B)
            mu = out[idx, nr + sel]
            ls = torch.clamp(out[idx, 2 * nr + sel], -7.0, 7.0)
            x = torch.clamp(mu + torch.exp(ls) * (torch.log(u2) - torch.log(1 - u2)), -1.0, 1.0 - 2.0 / Q)
            q = torch.floor(x * (Q / 2))
            audio = (q / (Q / 2))[:, None]
            out_audio[:, i] = audio[:, 0]
        return out_audio.numpy()
