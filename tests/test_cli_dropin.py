"""The drop-in boundary: the reference's eval CLIs import `tensorflow` (for tf.gfile / tf.logging /
tf.train.latest_checkpoint only), `wavenet.fastgen`, `wavenet.parallelgen`, `auxilaries.utils`,
`auxilaries.mel_extractor`.  With this repo (and its shims/ directory) on the path, the UNMODIFIED scripts
/root/reference/eval_parallel_wavenet.py and /root/reference/eval_wavenet.py must run: checkpoint directory holding a
TF-V2 bundle + `checkpoint` state file + one config json, a directory of wavs in, `gen_*.wav` out.

* CPU (here, where /root/reference exists): the scripts are executed with runpy; the three device classes are replaced
  by shape-faithful fakes, everything above the C ABI is the real code (shim, checkpoint resolution through the state
  file, bundle reader with EMA shadows, load_batch, padding, output lengths, wav writing).
* GPU box (no /root/reference): tests/cli_driver.py -- the same call sequence written out again -- drives the real
  engines; where the reference IS present next to a GPU the unmodified scripts run with the real engines too."""
import json
import os
import runpy
import sys

import numpy as np
import pytest
from scipy.io import wavfile

from conftest import CONFIG_DIR, ROOT
from oracle import wavenet_oracle as O

REF = '/root/reference'
SHIMS = os.path.join(ROOT, 'shims')


@pytest.fixture
def tf_shim(monkeypatch):
    try:
        import tensorflow  # noqa: F401
        if 'shim' not in getattr(tensorflow, '__version__', ''):
            pytest.skip('a real TensorFlow is installed; the shim is not in play')
    except ImportError:
        pass
    monkeypatch.syspath_prepend(ROOT)
    monkeypatch.setattr(sys, 'path', sys.path + [SHIMS])
    for m in [m for m in sys.modules if m == 'tensorflow' or m.startswith('tensorflow.')]:
        monkeypatch.delitem(sys.modules, m)
    yield
    from nsynth_wavenet_b200 import checkpoint
    checkpoint._ENGINES.clear()


def make_case(tmp_path, config, weights, n_samples=(154480, 3000)):
    """ckpt_dir (bundle with EMA shadows + raw variables + optimizer slots + state file + config json), wav dir."""
    from nsynth_wavenet_b200 import checkpoint as ckpt
    from tf_bundle_writer import write_bundle
    d = tmp_path / 'ckpt'
    d.mkdir()
    bundle = {}
    for k, v in weights.items():
        bundle[k + ckpt.EMA_SUFFIX] = v
        if v.size <= 4096 or k.endswith('start/W'):     # (the pure-Python CRC of the writer is slow: small tensors only)
            bundle[k] = np.zeros_like(v)                # raw variable: must NOT be the one used (fastgen.py:12-14)
            bundle[k + '/Adam'] = np.zeros_like(v)
    bundle['global_step'] = np.asarray(200000, np.int64)
    write_bundle(str(d / 'model.ckpt-200000'), bundle, num_shards=2, block_size=4096)
    (d / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-200000"\n'
                                  'all_model_checkpoint_paths: "model.ckpt-200000"\n')
    with open(os.path.join(CONFIG_DIR, config)) as f:
        (d / config).write_text(f.read())
    src = tmp_path / 'wavs'
    src.mkdir()
    rng = np.random.default_rng(0)
    for i, n in enumerate(n_samples):
        t = np.arange(n) / 16000.0
        x = 0.4 * np.sin(2 * np.pi * (220 + 110 * i) * t) + 0.02 * rng.standard_normal(n)
        wavfile.write(str(src / ('clip%d.wav' % i)), 16000, (np.clip(x, -1, 1) * 32767).astype(np.int16))
    return str(d), str(src), str(tmp_path / 'out')


class FakeIAF:
    def __init__(self, hparams, weights, device=0, num_mel=80, engine=None):
        assert 'iaf_1/start_conv/W' in weights and float(np.abs(weights['iaf_1/start_conv/W']).sum()) > 0   # EMA shadow, not the zeros
        self.hp, self._h = hparams, 1

    def length(self, F):
        return (F * 200 // 512) * 512

    def forward_host(self, mel, z=None, seed=0, quantize=True, want=('x',)):
        B, F, M = mel.shape
        assert M == 80
        return {'x': np.zeros((B, self.length(F)), np.float32)}

    def close(self):
        self._h = None


class FakeFastgen:
    def __init__(self, hparams, weights, device=0, num_mel=80, engine=None):
        assert float(np.abs(weights['conv_start/W']).sum()) > 0
        self.hp, self._h = hparams, 1

    def encode_host(self, mel):
        return np.zeros((mel.shape[0], mel.shape[1] * 200, self.hp.deconv_width), np.float32)

    def run_host(self, encoding, teacher_force=None, seed=0, want_out=False):
        return np.zeros(encoding.shape[:2], np.float32)

    def close(self):
        self._h = None


def run_cli(script, argv, monkeypatch):
    monkeypatch.setattr(sys, 'argv', [script] + argv)
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', os.environ.get('CUDA_VISIBLE_DEVICES', '0'))
    runpy.run_path(script, run_name='__main__')


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference checkout is not present on this machine')
def test_unmodified_eval_parallel_wavenet_runs_on_the_dropin_modules(tmp_path, tf_shim, monkeypatch):
    from nsynth_wavenet_b200.wavenet import parallelgen
    from nsynth_wavenet_b200.auxilaries import mel_extractor
    monkeypatch.setattr(parallelgen, 'IAFEngine', FakeIAF)
    monkeypatch.setattr(mel_extractor, 'batch_melspectrogram',
                        lambda y, device=0: np.zeros((y.shape[0], 1 + y.shape[1] // 200, 80), np.float32))
    import auxilaries.mel_extractor as shim_mel
    monkeypatch.setattr(shim_mel, 'batch_melspectrogram', mel_extractor.batch_melspectrogram)
    hp = O.load_hparams(os.path.join(CONFIG_DIR, 'parallel_wavenet.json'))
    ck, src, out = make_case(tmp_path, 'parallel_wavenet.json', O.init_student_weights(hp, seed=1))
    run_cli(os.path.join(REF, 'eval_parallel_wavenet.py'),
            ['--ckpt_dir', ck, '--source_path', src, '--save_path', out, '--batch_size', '2'], monkeypatch)
    rate, a = wavfile.read(os.path.join(out, 'gen_clip0.wav'))
    assert rate == 16000 and a.dtype == np.float32
    assert len(a) == 154112                      # (773 * 200 // 512) * 512 for the 154 480-sample file
    assert len(wavfile.read(os.path.join(out, 'gen_clip1.wav'))[1]) == 154112   # batch rows are padded to the longest


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference checkout is not present on this machine')
def test_unmodified_eval_wavenet_runs_on_the_dropin_modules(tmp_path, tf_shim, monkeypatch):
    from nsynth_wavenet_b200.wavenet import fastgen
    from nsynth_wavenet_b200.auxilaries import mel_extractor
    monkeypatch.setattr(fastgen, 'FastgenEngine', FakeFastgen)
    monkeypatch.setattr(mel_extractor, 'batch_melspectrogram',
                        lambda y, device=0: np.zeros((y.shape[0], 1 + y.shape[1] // 200, 80), np.float32))
    hp = O.load_hparams(os.path.join(CONFIG_DIR, 'wavenet_mol.json'))
    ck, src, out = make_case(tmp_path, 'wavenet_mol.json', O.init_teacher_weights(hp, seed=1))
    run_cli(os.path.join(REF, 'eval_wavenet.py'),
            ['--ckpt_dir', ck, '--source_path', src, '--save_path', out, '--sample_length', '154480'], monkeypatch)
    assert len(wavfile.read(os.path.join(out, 'gen_clip0.wav'))[1]) == 154600   # 773 frames x 200
    assert len(wavfile.read(os.path.join(out, 'gen_clip1.wav'))[1]) == 3200     # 16 frames x 200 (batch_size 1)


def test_tf_shim_surface(tmp_path, tf_shim):
    import tensorflow as tf
    assert 'shim' in tf.__version__
    (tmp_path / 'b.txt').write_text('x')
    (tmp_path / 'a.txt').write_text('x')
    assert tf.gfile.IsDirectory(str(tmp_path)) and not tf.gfile.IsDirectory(str(tmp_path / 'a.txt'))
    assert tf.gfile.ListDirectory(str(tmp_path)) == ['a.txt', 'b.txt']
    tf.logging.set_verbosity('INFO')
    tf.logging.set_verbosity(tf.logging.WARN)
    tf.logging.info('not shown')
    assert tf.train.latest_checkpoint(str(tmp_path)) is None
    assert not tf.train.checkpoint_exists(None) and not tf.train.checkpoint_exists(str(tmp_path / 'model.ckpt-1'))
    (tmp_path / 'model.ckpt-7.index').write_bytes(b'')
    (tmp_path / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-7"\n')
    assert tf.train.latest_checkpoint(str(tmp_path)) == str(tmp_path / 'model.ckpt-7')
    assert tf.train.checkpoint_exists(str(tmp_path / 'model.ckpt-7'))


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.parametrize('kind', ['parallel_wavenet', 'wavenet'])
def test_eval_cli_sequence_with_real_engines(tmp_path, tf_shim, monkeypatch, kind):
    """GPU: wav directory -> gen_*.wav through the real engines.  Uses the unmodified reference script when the
    reference checkout is next to the GPU, else tests/cli_driver.py (the same call sequence)."""
    if kind == 'parallel_wavenet':
        hp = O.load_hparams(os.path.join(CONFIG_DIR, 'parallel_wavenet.json'))
        ck, src, out = make_case(tmp_path, 'parallel_wavenet.json', O.init_student_weights(hp, seed=1), (154480, 30000))
        argv = ['--ckpt_dir', ck, '--source_path', src, '--save_path', out, '--batch_size', '2']
        want = {'gen_clip0.wav': 154112, 'gen_clip1.wav': 154112}
    else:
        hp = O.load_hparams(os.path.join(CONFIG_DIR, 'wavenet_mol.json'))
        ck, src, out = make_case(tmp_path, 'wavenet_mol.json', O.init_teacher_weights(hp, seed=1), (6000, 3000))
        argv = ['--ckpt_dir', ck, '--source_path', src, '--save_path', out, '--sample_length', '4000']
        want = {'gen_clip0.wav': 4200, 'gen_clip1.wav': 3200}     # (1 + 4000 // 200) * 200, (1 + 3000 // 200) * 200
    script = os.path.join(REF, 'eval_%s.py' % kind)
    if os.path.exists(script):
        run_cli(script, argv, monkeypatch)
    else:
        import cli_driver
        a = dict(zip(argv[::2], argv[1::2]))
        cli_driver.generate(kind, a['--ckpt_dir'], a['--source_path'], a['--save_path'],
                            sample_length=int(a.get('--sample_length', -1)), batch_size=int(a.get('--batch_size', 1)))
    for name, n in want.items():
        rate, a = wavfile.read(os.path.join(out, name))
        assert rate == 16000 and a.dtype == np.float32 and len(a) == n, (name, len(a))
        assert np.all(np.isfinite(a)) and np.abs(a).max() <= 1.0 and np.abs(a).max() > 0
