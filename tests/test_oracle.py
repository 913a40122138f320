"""Pins the CPU oracle (oracle/wavenet_oracle.py) against everything the reference's own
tests pin for the generation path and against independent torch ops."""
import os
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wavenet_oracle as O
from conftest import GOLDEN_DIR, synth_inputs

RNG = np.random.default_rng(1234)


@pytest.mark.parametrize('d', [1, 2, 8, 64, 512])
def test_conv_closed_form_equals_literal_pipeline(d):
    # masked.py:160-232: time_to_batch -> pad -> VALID conv -> batch_to_time
    x = RNG.normal(size=(2, 1024, 6))
    W = RNG.normal(size=(1, 3, 6, 5))
    b = RNG.normal(size=5)
    assert np.allclose(O.conv1d(x, W, b, d), O.conv1d_literal(x, W, b, d), atol=1e-12)


def test_conv_matches_torch_causal_dilated():
    x = RNG.normal(size=(2, 256, 4))
    W = RNG.normal(size=(1, 3, 4, 7))
    b = RNG.normal(size=7)
    for d in (1, 4, 32):
        xt = F.pad(torch.tensor(x).permute(0, 2, 1), (2 * d, 0))
        wt = torch.tensor(W[0]).permute(2, 1, 0)  # w_t[o,c,j] = W[0,j,c,o]  
        y = F.conv1d(xt, wt, torch.tensor(b), dilation=d).permute(0, 2, 1).numpy()
        assert np.allclose(O.conv1d(x, W, b, d), y, atol=1e-10)


def test_queue_form_equals_closed_form_tap_order_and_zero_history():
    # masked.py:352-376: W[:,0] <- state from 2*rate ago, W[:,1] <- rate ago, W[:,2] <- now
    T, rate, C = 40, 4, 3
    x = RNG.normal(size=(1, T, C))
    W = RNG.normal(size=(1, 3, C, 2))
    b = RNG.normal(size=2)
    q1, q2 = O._Queue(rate, (1, C), np.float64), O._Queue(rate, (1, C), np.float64)
    ys = []
    for t in range(T):
        s1 = q1.dequeue(); q1.enqueue(x[:, t]); s2 = q2.dequeue(); q2.enqueue(s1)
        ys.append(s2 @ W[0, 0] + s1 @ W[0, 1] + x[:, t] @ W[0, 2] + b)
    assert np.allclose(np.stack(ys, 1), O.conv1d(x, W, b, rate), atol=1e-12)


def test_shift_right():
    x = RNG.normal(size=(2, 5, 1))
    y = O.shift_right(x)
    assert np.all(y[:, 0] == 0) and np.array_equal(y[:, 1:], x[:, :-1])


@pytest.mark.parametrize('k,s', [(40, 10), (80, 20)])
def test_trans_conv_is_adjoint_of_same_strided_conv(k, s):
    # TF defines conv2d_transpose as the input-gradient of conv2d (SAME, stride s).
    L, cin, cout = 7, 3, 4
    x = RNG.normal(size=(1, L, cin))
    K = RNG.normal(size=(1, k, cout, cin))
    y = O.trans_conv1d(x, K, np.zeros(cout), s)
    assert y.shape == (1, s * L, cout)
    # forward conv SAME: pad_total = k - s, pad_left = (k - s)//2
    inp = torch.zeros(1, cout, s * L, dtype=torch.float64, requires_grad=True)
    pl = (k - s) // 2
    padded = F.pad(inp, (pl, k - s - pl))
    wt = torch.tensor(K[0]).permute(2, 1, 0)  # [cin(out of fwd), cout(in of fwd), k]
    fwd = F.conv1d(padded, wt, stride=s)
    assert fwd.shape[-1] == L
    (fwd * torch.tensor(x).permute(0, 2, 1)).sum().backward()
    assert np.allclose(inp.grad.permute(0, 2, 1).numpy(), y, atol=1e-10)
    # and equals conv_transpose1d(stride=s, padding=(k-s)//2)  
    yt = F.conv_transpose1d(torch.tensor(x).permute(0, 2, 1), torch.tensor(K[0]).permute(2, 1, 0),
                            stride=s, padding=pl).permute(0, 2, 1).numpy()
    assert np.allclose(yt, y, atol=1e-10)


def test_condition_centre_trim():
    x = np.zeros((1, 7680, 2)); c = np.arange(7800, dtype=np.float64)[None, :, None] + np.zeros((1, 1, 2))
    assert O.condition(x, c)[0, 0, 0] == 60 and O.condition(x, c)[0, -1, 0] == 7739


def test_scale_transform_matches_reference_numpy_twin():
    # tests/test_scale.py:67-78
    d = RNG.normal(size=20000)
    ref = np.clip(np.log(1.0 + np.exp(d)), np.exp(-9.0), np.exp(7.0))
    s, ls = O.scale_log_scale_fn(d)
    assert np.allclose(s, ref) and np.allclose(ls, np.log(ref))


def test_clip_quant_scale():
    # tests/test_clip_quant_scale.py:7-18
    x = np.array([-2.0, -1.0, -0.5, 0.0, 1e-5, 0.99999, 1.0, 3.0], np.float32)
    y = O.clip_quant_scale(x, 65536, False)
    assert y.min() == -1.0 and y.max() == np.float32(1 - 2 / 65536)
    assert np.all(y * 32768 == np.floor(y * 32768))
    assert np.all(np.abs(y - np.clip(x, -1, 1 - 2 / 65536)) <= 1 / 32768 + 1e-7)
    ym = O.clip_quant_scale(x, 256, True)
    assert ym.min() >= -1.0 and ym.max() <= 1.0


def test_lengths_of_reference_fixture():
    # tests/pred_data-*: 154480-sample source -> 773 frames -> fastgen 154600, parallelgen 154112
    hp = O.load_hparams(os.path.join(os.path.dirname(GOLDEN_DIR), '..', 'nsynth_wavenet_b200',
                                     'config_jsons', 'parallel_wavenet.json'))
    frames = 1 + 154480 // 200
    assert frames == 773 and frames * 200 == 154600
    assert O.iaf_length(frames, hp) == 154112
    assert O.iaf_length(39, hp) == 7680 and O.iaf_length(21, hp) == 4096


def test_student_invariants(student_hp):
    # tests/test_parallel_wavenet.py:63-73
    w = O.init_student_weights(student_hp, seed=12345)
    mel, z = synth_inputs(student_hp, 1, 6)
    o = O.student_feed_forward(w, student_hp, mel, z, np.float64)
    assert np.all(o['scale_tot'] > 0)
    assert np.allclose(o['x'], o['rand_input'] * o['scale_tot'] + o['mean_tot'])
    assert abs(o['mean_tot'].mean()) < 0.1          # "should be close to 0.0"
    assert 0.03 < o['scale_tot'].mean() < 0.2       # softplus(-0.3)^4 ~ 0.096
    assert np.allclose(o['log_scale_tot'], np.log(o['scale_tot']), atol=1e-9)


def test_student_fp32_twin_close_to_fp64(student_hp):
    w = O.init_student_weights(student_hp, seed=12345)
    mel, z = synth_inputs(student_hp, 1, 6)
    a = O.student_feed_forward(w, student_hp, mel, z, np.float64)
    b = O.student_feed_forward(w, student_hp, mel, z, np.float32)
    for k in ('mean_tot', 'scale_tot', 'log_scale_tot'):
        assert np.abs(a[k] - b[k]).max() < 5e-6


def test_golden_iaf_vectors_reproduce(student_hp, clarinet_hp):
    for hp, name in ((student_hp, 'iaf_logistic_1x21.npz'), (clarinet_hp, 'iaf_gauss_2x6.npz')):
        g = np.load(os.path.join(GOLDEN_DIR, name))
        w = O.init_student_weights(hp, seed=12345, bias_std=0.02)
        if g['z'].shape[1] > 2048:
            o = O.parallelgen_forward(w, hp, g['mel'], g['z'], np.float32)
            tol = 2e-5
        else:
            o = O.parallelgen_forward(w, hp, g['mel'], g['z'], np.float64)
            tol = 1e-6
        for k in ('mean_tot', 'scale_tot', 'log_scale_tot'):
            assert np.abs(o[k] - g[k]).max() < tol, (name, k)


def test_teacher_random_init_likelihood(teacher_hp):
    # tests/test_wavenet.py:66-69: exp(-loss) is of the order 1/65536 at random init
    hp = teacher_hp
    w = O.init_teacher_weights(hp, seed=12345)
    rng = np.random.default_rng(3)
    mel = rng.uniform(0, 1, (1, 3, 80)).astype(np.float32)
    wav = rng.uniform(-0.3, 0.3, (1, 512)).astype(np.float32)
    enc = O.deconv_stack(mel, w, hp, '', np.float32)
    out = O.teacher_feed_forward(w, hp, wav, None, np.float32, mel_en=enc[:, :572])['out_params']
    assert out.shape == (1, 512, 30)
    loss = O.mol_loss(out.astype(np.float64), wav.astype(np.float64), 65536)
    assert 0.05 / 65536 < np.exp(-loss) < 20.0 / 65536


def test_fastgen_equals_full_sequence_teacher_under_teacher_forcing(teacher_hp):
    # Fastgen.sample (wavenet.py:379-514) step-by-step == Wavenet.feed_forward (:180-291)
    # when cond is untrimmed (fastgen.py:157 feeds encoding[:, i] with no centre trim).
    hp = teacher_hp
    w = O.init_teacher_weights(hp, seed=5, bias_std=0.02)
    rng = np.random.default_rng(4)
    T = 24
    enc = rng.uniform(-1, 1, (1, T, 256))
    wav = rng.uniform(-0.5, 0.5, (1, T))
    r = O.fastgen_run(w, hp, enc, np.float64, teacher_force=wav)
    # full-sequence teacher needs T % 512 == 0 only for the literal ttb path; closed form does not
    full = O.teacher_feed_forward(w, hp, wav, None, np.float64, mel_en=enc)['out_params']
    assert np.allclose(r['out'], full, atol=1e-10)


def test_golden_fastgen_vector_reproduces(teacher_hp):
    g = np.load(os.path.join(GOLDEN_DIR, 'fastgen_tf_1x96.npz'))
    w = O.init_teacher_weights(teacher_hp, seed=12345, bias_std=0.02)
    r = O.fastgen_run(w, teacher_hp, g['enc'][:, :16], np.float64, teacher_force=g['wav'][:, :16])
    assert np.abs(r['out'] - g['out'][:, :16]).max() < 1e-6


def test_mol_sample_is_deterministic_given_uniforms():
    out = RNG.normal(size=(5, 30))
    u1 = RNG.uniform(1e-5, 1 - 1e-5, (5, 10)); u2 = RNG.uniform(1e-5, 1 - 1e-5, 5)
    q = O.mol_sample(out, 65536, u1, u2)
    assert q.dtype == np.int32 and q.min() >= -32768 and q.max() <= 32767
    assert np.array_equal(q, O.mol_sample(out, 65536, u1, u2))


def test_weight_norm_folding():
    V = RNG.normal(size=(1, 3, 4, 5)).astype(np.float32)
    g = RNG.uniform(0.5, 2, 5).astype(np.float32)
    w = O.fold_weight_norm({'a/W_V': V, 'a/W_g': g, 'a/biases': np.zeros(5, np.float32)})
    nrm = np.sqrt((w['a/W'].astype(np.float64) ** 2).sum(axis=(0, 1, 2)))
    assert np.allclose(nrm, g, rtol=1e-5) and 'a/W_V' not in w


def test_kl_loss_gauss_matches_torch_distributions():
    """kl_loss_gauss (parallel_wavenet.py:404-428): the per-sample term is KL(N(m_q, s_q) || N(m_p, s_p)); an
    independent implementation of that closed form exists in torch.distributions, so this pins the restatement
    (including the -7 floor on the teacher's log-scale parameter, loss_func.py:71)."""
    import torch
    from torch.distributions import Normal, kl_divergence
    rng = np.random.default_rng(41)
    B, T = 3, 500
    te = np.stack([rng.normal(0, 0.3, (B, T)), rng.uniform(-9, -1, (B, T))], axis=-1)      # some below the floor
    mean = rng.normal(0, 0.3, (B, T))
    ls = rng.uniform(-6, -1, (B, T))
    scale = np.exp(ls)
    got = O.kl_loss_gauss(te, mean, scale, ls)
    lp = np.maximum(te[..., 1], -7.0)
    q = Normal(torch.from_numpy(mean), torch.from_numpy(scale))
    p = Normal(torch.from_numpy(te[..., 0]), torch.from_numpy(np.exp(lp)))
    kl = float(kl_divergence(q, p).mean())
    reg = float(((lp - ls) ** 2).mean())
    assert abs(got['kl'] - kl) < 1e-9 * max(1.0, abs(kl))
    assert abs(got['reg'] - reg) < 1e-12
    assert abs(got['kl_loss'] - (kl + 4 * reg)) < 1e-9 * max(1.0, abs(kl))
    # identical distributions: every term vanishes
    same = O.kl_loss_gauss(np.stack([mean, ls], -1), mean, scale, ls)
    assert abs(same['kl_loss']) < 1e-12
    # float32 twin stays close to the float64 truth on the same inputs
    g32 = O.kl_loss_gauss(te.astype(np.float32), mean.astype(np.float32), scale.astype(np.float32),
                          ls.astype(np.float32))
    assert abs(g32['kl_loss'] - got['kl_loss']) < 1e-4 * abs(got['kl_loss'])


def test_mel_oracle_stft_matches_scipy():
    """The oracle's centred, reflect-padded, window-zero-padded STFT (librosa.stft as called at
    auxilaries/mel_extractor.py:68-72) against scipy.signal.stft, an independent implementation of the framing
    and the FFT, fed the same padded signal and the same 2048-point window."""
    from scipy import signal
    from oracle import mel_oracle
    rng = np.random.default_rng(3)
    y = rng.uniform(-0.5, 0.5, 5000)
    D = mel_oracle._stft(y)                                            # [1025, frames]
    win = np.zeros(2048)
    win[624:1424] = signal.get_window('hann', 800, fftbins=True)       # periodic hann, centred in n_fft
    ypad = np.pad(y, 1024, mode='reflect')
    _, _, Z = signal.stft(ypad, window=win, nperseg=2048, noverlap=2048 - 200, nfft=2048, boundary=None,
                          padded=False, return_onesided=True)
    Z = Z * win.sum()                                                  # scipy scales by 1 / sum(window)
    assert Z.shape == D.shape == (1025, 1 + 5000 // 200)
    assert np.abs(Z - D).max() < 1e-9


def test_power_loss_stft_oracle_is_the_documented_framing():
    """oracle.mel_oracle.tf_stft against a literal DFT of hand-cut frames (tf.contrib.signal.stft with pad_end=True:
    frames start at j*step, ceil(N/step) of them, zeros past the end, periodic hann on the first frame_length samples
    of the fft_length frame), and against scipy's STFT for the frames scipy also has."""
    import scipy.signal
    from oracle import mel_oracle as MO
    rng = np.random.default_rng(3)
    N = 1530
    y = rng.standard_normal((2, N))
    S = MO.tf_stft(y)
    assert S.shape == (2, 8, 1025)                       # ceil(1530 / 200) = 8 frames
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(800) / 800)
    for j in (0, 3, 7):
        seg = np.zeros(800)
        chunk = y[1, j * 200:j * 200 + 800]
        seg[:len(chunk)] = chunk
        k = np.asarray([0, 1, 17, 512, 1024])
        n = np.arange(800)
        lit = (seg * win)[None, :] * np.exp(-2j * np.pi * k[:, None] * n[None, :] / 2048)
        assert np.abs(S[1, j, k] - lit.sum(1)).max() < 1e-9
    f, t, Z = scipy.signal.stft(y[0], window=win, nperseg=800, noverlap=600, nfft=2048, boundary=None, padded=False)
    Z = Z * win.sum()                                    # scipy normalises by the window sum
    assert np.abs(S[0, :Z.shape[1]].T - Z).max() < 1e-9   # the frames that lie fully inside the signal
    # power loss: trims the longer wave about its centre, squared magnitude difference, priority bins weighted 0.5
    a, b = rng.standard_normal((2, 1600)), rng.standard_normal((2, 1536))
    pl = MO.power_loss(a, b)
    d = (np.abs(MO.tf_stft(a[:, 32:32 + 1536])) - np.abs(MO.tf_stft(b))) ** 2
    assert abs(pl - (0.5 * d.mean() + 0.5 * d[:, :, :MO.PRIORITY_FREQ].mean())) < 1e-12
    assert MO.PRIORITY_FREQ == 384 and MO.power_loss(a, a) == 0.0


def test_resize_conv_oracle_equals_nearest_upsampling_plus_same_conv():
    """oracle.resize_conv1d (masked.resize_conv1d, masked.py:294-322) against torch: nearest-neighbour interpolation to
    L*stride followed by a stride-1 convolution with TensorFlow's SAME padding ((k-1)//2 zeros on the left, the rest on
    the right) -- for an even filter length the padding is asymmetric, which is what this pins."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    for (k, s, cin, cout, L) in ((40, 10, 8, 6, 5), (80, 20, 6, 4, 3), (5, 3, 4, 4, 7), (4, 2, 4, 4, 6)):
        x = rng.standard_normal((2, L, cin))
        W = rng.standard_normal((1, k, cin, cout))
        b = rng.standard_normal(cout)
        got = O.resize_conv1d(x, W, b, s, None)
        xt = torch.from_numpy(x).permute(0, 2, 1)                        # [B, C, L]
        up = F.interpolate(xt, scale_factor=s, mode='nearest')
        pl = (k - 1) // 2
        up = F.pad(up, (pl, k - 1 - pl))
        wt = torch.from_numpy(W[0]).permute(2, 1, 0)                     # [cout, cin, k]
        ref = F.conv1d(up, wt, torch.from_numpy(b)).permute(0, 2, 1).numpy()
        assert got.shape == (2, L * s, cout)
        assert np.abs(got - ref).max() < 1e-10, (k, s)
    hp = Namespace(deconv_config=[[40, 10], [80, 20]], deconv_width=256, use_resize_conv=True, upsample_act='leaky_relu',
                   width=512, skip_width=256, filter_length=3, num_layers=2, num_stages=2, use_mu_law=False,
                   loss_type='mol', mol_mix=10, double_gate_width=False)
    w = O.init_teacher_weights(hp, seed=3)
    assert 'resize_conv_1/W' in w and w['resize_conv_1/W'].shape == (1, 40, 80, 256) and 'trans_conv_1/kernel' not in w
    enc = O.deconv_stack(rng.uniform(0, 1, (1, 3, 80)), w, hp, '', np.float64)
    assert enc.shape == (1, 600, 256)
