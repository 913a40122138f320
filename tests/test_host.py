"""Host-side logic of the drop-in modules (no GPU)."""
import os
from argparse import Namespace

import numpy as np
import pytest
from scipy.io import wavfile

from nsynth_wavenet_b200 import checkpoint as ckpt
from nsynth_wavenet_b200 import engine
from nsynth_wavenet_b200.auxilaries import mel_extractor, utils
from nsynth_wavenet_b200.wavenet import fastgen, parallelgen


def test_checkpoint_ema_shadow_preferred(tmp_path):
    w = {'iaf_1/out1/W': np.ones((1, 1, 2, 2), np.float32)}
    p = ckpt.save_weights(str(tmp_path / 'model.ckpt-1'), w, ema=True)
    raw = dict(np.load(p))
    raw['iaf_1/out1/W'] = np.zeros((1, 1, 2, 2), np.float32)  # stale non-EMA copy
    np.savez(p, **raw)
    got = ckpt.load_weights(str(tmp_path / 'model.ckpt-1'))
    assert got['iaf_1/out1/W'].sum() == 4
    assert ckpt.load_weights(str(tmp_path))['iaf_1/out1/W'].sum() == 4  # directory form


def test_checkpoint_unshadowed_teacher_deconv(tmp_path):
    # parallelgen.py:32-39: frozen deconv vars are restored from their plain names
    p = str(tmp_path / 'c.npz')
    np.savez(p, **{'iaf_share/trans_conv_1/kernel': np.full(3, 2.0, np.float32),
                   'iaf_share/trans_conv_1/kernel/ExponentialMovingAverage': np.zeros(3, np.float32)})
    hp = Namespace(use_teacher_deconv=True, use_resize_conv=False)
    got = ckpt.load_weights(p, parallelgen._unshadowed(hp))
    assert got['iaf_share/trans_conv_1/kernel'][0] == 2.0


def test_missing_checkpoint_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        ckpt.load_weights(str(tmp_path / 'nope'))


def test_shadow_dict_names():
    class V:  # mimics tf.Variable.name
        def __init__(self, n): self.name = n
    d = fastgen.get_ema_shadow_dict([V('conv_start/W:0')])
    assert list(d) == ['conv_start/W/ExponentialMovingAverage']
    assert list(parallelgen.get_default_shadow_dict(['a/b'])) == ['a/b']


def test_load_batch_pads_and_save_batch_roundtrip(tmp_path):
    a = (np.sin(np.arange(1000) / 10) * 0.5).astype(np.float32)
    b = a[:600]
    pa, pb = str(tmp_path / 'a.wav'), str(tmp_path / 'b.wav')
    wavfile.write(pa, 16000, (a * 32767).astype(np.int16))
    wavfile.write(pb, 16000, (b * 32767).astype(np.int16))
    batch = fastgen.load_batch([pa, pb], sample_length=-1)
    assert batch.shape == (2, 1000) and np.all(batch[1, 600:] == 0)
    assert fastgen.load_batch([pa, pb], sample_length=500).shape == (2, 500)
    out = [str(tmp_path / 'o0.wav'), str(tmp_path / 'o1.wav')]
    fastgen.save_batch(batch.astype(np.float32), out)
    rate, data = wavfile.read(out[0])
    assert rate == 16000 and data.dtype == np.float32 and len(data) == 1000
    np.save(str(tmp_path / 'e.npy'), np.zeros((7, 3)))
    np.save(str(tmp_path / 'f.npy'), np.zeros((5, 3)))
    assert fastgen.load_batch([str(tmp_path / 'e.npy'), str(tmp_path / 'f.npy')]).shape == (14, 3)


def test_mel_tables_reproduce_the_oracle_stft_and_filterbank():
    """Host-side table preparation of the GPU mel front-end (no GPU needed): the window-folded twiddle tables
    contract a frame's 800 window taps to the same magnitudes as the oracle's zero-padded 2048-point rfft, and the
    filterbank equals the oracle's."""
    from oracle import mel_oracle
    tc, ts = mel_extractor._build_twiddles()
    assert tc.shape == (800, 1025) and ts.shape == (800, 1025)
    rng = np.random.default_rng(0)
    y = rng.uniform(-0.5, 0.5, 4000)
    D = np.abs(mel_oracle._stft(y))                                   # [1025, frames], float64
    ypad = np.pad(y, 1024, mode='reflect')
    for j in (0, 3, D.shape[1] - 1):
        taps = ypad[j * 200 + 624:j * 200 + 624 + 800]
        re, im = taps @ tc.astype(np.float64), taps @ ts.astype(np.float64)
        assert np.abs(np.sqrt(re * re + im * im) - D[:, j]).max() < 2e-5 * max(1.0, D[:, j].max())
    basis = mel_extractor._build_mel_basis()
    assert basis.shape == (80, 1025) and np.all(basis >= 0) and np.all(basis.sum(1) > 0)
    assert np.array_equal(basis, mel_oracle._build_mel_basis())
    m = mel_oracle.melspectrogram(rng.uniform(-0.5, 0.5, 154480).astype(np.float32))
    assert m.shape == (773, 80) and m.dtype == np.float32 and m.min() >= 0 and m.max() <= 1
    assert mel_oracle.batch_melspectrogram(y[None, :].astype(np.float32)).shape == (1, 21, 80)


def test_mu_law_roundtrip_numpy():
    x = np.linspace(-0.99, 0.99, 101)
    q = utils.mu_law_numpy(x)
    assert q.min() >= -128 and q.max() <= 127
    assert np.abs(utils.inv_mu_law_numpy(q) - x).max() < 0.05


def test_config_mapping(student_hp, clarinet_hp, teacher_hp):
    c = engine.iaf_config(student_hp, engine='ffma')
    assert (c.num_flows, list(c.num_iaf_layers)[:4], c.share_deconv, c.loss_type) == (4, [10, 10, 10, 30], 1, 0)
    assert (c.deconv_filter[1], c.deconv_stride[1], c.upsample_act) == (80, 20, 2)
    g = engine.iaf_config(clarinet_hp, engine='tc')
    assert (g.share_deconv, g.loss_type, g.engine) == (0, 1, 1)   # parallel_wavenet.py:129-134
    t = engine.wavenet_config(teacher_hp)
    assert (t.gate_width, t.out_width, t.skip_width) == (512, 30, 256)
    ce = Namespace(**{**vars(teacher_hp), 'loss_type': 'ce', 'use_mu_law': True})
    delattr(ce, 'double_gate_width')
    t2 = engine.wavenet_config(ce)
    assert (t2.gate_width, t2.out_width) == (1024, 256)           # wavenet.py:106 default True


def test_engine_fold_weight_norm_matches_oracle():
    from oracle import wavenet_oracle as O
    rng = np.random.default_rng(2)
    w = {'x/W_V': rng.normal(size=(1, 3, 4, 5)).astype(np.float32),
         'x/W_g': rng.uniform(0.5, 2, 5).astype(np.float32),
         'y/kernel_V': rng.normal(size=(1, 4, 6, 3)).astype(np.float32),
         'y/kernel_g': rng.uniform(0.5, 2, 6).astype(np.float32)}
    a, b = engine.fold_weight_norm(w), O.fold_weight_norm(w)
    assert set(a) == set(b) == {'x/W', 'y/kernel'}
    for k in a:
        assert np.allclose(a[k], b[k])


def test_load_audio_sample_formats(tmp_path):
    """utils.load_audio (utils.py:55-69 goes through librosa / soundfile): int16, int32, uint8 (offset 128) and
    float32 wavs of the same signal load to the same [-1, 1) floats."""
    x = np.sin(np.arange(400) / 7.0) * 0.5
    p16, p32, p8, pf = (str(tmp_path / n) for n in ('a16.wav', 'a32.wav', 'a8.wav', 'af.wav'))
    wavfile.write(p16, 16000, np.round(x * 32767).astype(np.int16))
    wavfile.write(p32, 16000, np.round(x * (2 ** 31 - 1)).astype(np.int32))
    wavfile.write(p8, 16000, np.round(x * 127 + 128).astype(np.uint8))
    wavfile.write(pf, 16000, x.astype(np.float32))
    ref = utils.load_audio(pf, -1)
    assert ref.dtype == np.float32 and np.abs(ref - x).max() < 1e-6
    assert np.abs(utils.load_audio(p16, -1) - x).max() < 1e-4
    assert np.abs(utils.load_audio(p32, -1) - x).max() < 1e-6
    a8 = utils.load_audio(p8, -1)
    assert np.abs(a8 - x).max() < 1.0 / 128 and abs(a8.mean()) < 0.01     # no DC offset of +1
    assert len(utils.load_audio(p16, 100)) == 100
    with pytest.raises(ValueError, match='sample rate'):
        wavfile.write(str(tmp_path / 'r.wav'), 22050, x.astype(np.float32))
        utils.load_audio(str(tmp_path / 'r.wav'))


def test_weights_init_matches_the_test_side_generator():
    """nsynth_wavenet_b200.weights_init (bench.py / smoke use it) draws the same tensors as the oracle's initialiser,
    for the student, the MoL teacher and the double-gate CE teacher."""
    from oracle import wavenet_oracle as O
    from nsynth_wavenet_b200 import weights_init as WI
    from conftest import load_hparams
    for cfg, a, b in (('parallel_wavenet.json', WI.init_student_weights, O.init_student_weights),
                      ('parallel_wavenet_gauss.json', WI.init_student_weights, O.init_student_weights),
                      ('wavenet_mol.json', WI.init_teacher_weights, O.init_teacher_weights),
                      ('wavenet_ce.json', WI.init_teacher_weights, O.init_teacher_weights)):
        hp = load_hparams(cfg)
        if cfg == 'wavenet_ce.json':
            hp = Namespace(**{**vars(hp), 'num_layers': 2})
        wa, wb = a(hp, seed=7, bias_std=0.01), b(hp, seed=7, bias_std=0.01)
        assert set(wa) == set(wb)
        assert all(np.array_equal(wa[k], wb[k]) for k in wa), cfg


def test_engine_cache_key_follows_the_checkpoint_file(tmp_path, monkeypatch):
    """checkpoint.cached_engine: same checkpoint -> same engine object; rewritten checkpoint -> a new one."""
    made = []

    class Fake:
        def __init__(self, hp, weights, **kw):
            self._h, self.w = 1, weights
            made.append(self)

        def close(self):
            self._h = None
    hp = Namespace(a=1)
    p = ckpt.save_weights(str(tmp_path / 'm'), {'x/W': np.ones(3, np.float32)})
    e1 = ckpt.cached_engine(Fake, 'iaf', hp, p, device=0)
    e2 = ckpt.cached_engine(Fake, 'iaf', hp, p, device=0)
    assert e1 is e2 and len(made) == 1
    assert ckpt.cached_engine(Fake, 'iaf', Namespace(a=2), p, device=0) is not e1      # other hparams
    os.utime(p, (1, 1))
    ckpt.save_weights(str(tmp_path / 'm'), {'x/W': np.zeros(3, np.float32)})
    e3 = ckpt.cached_engine(Fake, 'iaf', hp, p, device=0)
    assert e3 is not e1 and e1._h is None and e3.w['x/W'].sum() == 0
    ckpt.clear_engine_cache()
    assert e3._h is None
