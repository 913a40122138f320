"""GPU parity of the persistent autoregressive kernel (fastgen) against the CPU oracle.
Sampling is stochastic, so parity is on the pre-sample tensor out[B,T,30] under teacher
forcing (BASELINE north_star: within 1e-4)."""
import os

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_engine(hp, engine='ffma', seed=12345, bias_std=0.02):
    from nsynth_wavenet_b200 import FastgenEngine
    w = O.init_teacher_weights(hp, seed=seed, bias_std=bias_std)
    return FastgenEngine(hp, w, device=0, engine=engine), w


@pytest.mark.timeout(300)
def test_teacher_forced_out_matches_golden(teacher_hp):
    eng, _ = make_engine(teacher_hp)
    g = np.load(os.path.join(GOLDEN_DIR, 'fastgen_tf_1x96.npz'))
    audio, out = eng.run_host(g['enc'], teacher_force=g['wav'], want_out=True)
    err = np.abs(out - g['out']).max()
    print('fastgen teacher-forced max-abs err', err, 'kernel ms', eng.last_timing())
    assert err < TOL, err
    assert np.array_equal(audio, g['wav'])


@pytest.mark.timeout(600)
def test_long_run_ring_wraparound_matches_oracle(teacher_hp):
    # 2100 steps > 2*1024+1: every history ring (d up to 512) wraps at least once
    hp = teacher_hp
    eng, w = make_engine(hp)
    rng = np.random.default_rng(21)
    T = 2100
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
    _, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    err = np.abs(out - ref).max()
    print('fastgen 2100-step max-abs err', err)
    assert err < TOL, err


@pytest.mark.timeout(300)
def test_free_running_feedback_is_self_consistent(teacher_hp):
    hp = teacher_hp
    eng, w = make_engine(hp)
    rng = np.random.default_rng(22)
    T = 300
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    a1, o1 = eng.run_host(enc, seed=5, want_out=True)
    a2, o2 = eng.run_host(enc, seed=5, want_out=True)
    assert np.array_equal(a1, a2) and np.array_equal(o1, o2)          # deterministic per seed
    a3 = eng.run_host(enc, seed=6)
    assert not np.array_equal(a1, a3)
    assert a1.min() >= -1.0 and a1.max() <= 1.0 - 2.0 / 65536
    assert np.all(a1 * 32768 == np.floor(a1 * 32768))                 # on the 16-bit grid
    # feeding the kernel's own samples back under teacher forcing reproduces its outputs,
    # and the oracle agrees with them: the feedback path carries the right sample
    _, o_tf = eng.run_host(enc, teacher_force=a1, want_out=True)
    assert np.abs(o_tf - o1).max() < 1e-6
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=a1)['out']
    assert np.abs(o1 - ref).max() < TOL


@pytest.mark.timeout(300)
def test_batch_rows_and_tc_cond_engine(teacher_hp):
    hp = teacher_hp
    eng, w = make_engine(hp, engine='tc')
    rng = np.random.default_rng(23)
    T = 64
    enc = rng.uniform(-1, 1, (2, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (2, T)).astype(np.float32)
    _, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    assert np.abs(out - ref).max() < TOL


@pytest.mark.timeout(300)
def test_encode_matches_oracle_deconv(teacher_hp):
    hp = teacher_hp
    eng, w = make_engine(hp)
    rng = np.random.default_rng(24)
    mel = rng.uniform(0, 1, (2, 4, 80)).astype(np.float32)
    enc = eng.encode_host(mel)
    ref = O.deconv_stack(mel, w, hp, '', np.float64)
    assert enc.shape == (2, 800, 256)
    assert np.abs(enc - ref).max() < 2e-5


@pytest.mark.timeout(300)
def test_gauss_head_teacher_forced(teacher_hp):
    from argparse import Namespace
    hp = Namespace(**{**vars(teacher_hp), 'loss_type': 'gauss'})
    from nsynth_wavenet_b200 import FastgenEngine
    w = O.init_teacher_weights(hp, seed=7, bias_std=0.02)
    eng = FastgenEngine(hp, w, device=0, engine='ffma')
    rng = np.random.default_rng(25)
    T = 48
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
    _, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    assert out.shape == (1, T, 2) and np.abs(out - ref).max() < TOL
    a = eng.run_host(enc, seed=1)
    assert np.all(np.isfinite(a)) and a.min() >= -1.0 and a.max() < 1.0


@pytest.mark.timeout(600)
@pytest.mark.parametrize('env', [
    {'NSW_FASTGEN_GENERIC': '1'},                                   # run-time-switch build, default switches
    {'NSW_FASTGEN_FLAGS': '0', 'NSW_FASTGEN_L2LAST': '0'},          # round-start switches (st publish, LDG history)
    {'NSW_FASTGEN_FLAGS': '2564'},                                  # one replica
    {'NSW_FASTGEN_FLAGS': '3584', 'NSW_FASTGEN_GENERIC': '1'},      # bulk-copy polling (run-time build only)
])
def test_switch_variants_agree_bit_for_bit_with_the_product_build(teacher_hp, env, monkeypatch):
    """Every publish / poll / prefetch variant moves the same fp32 values through the same arithmetic, so
    audio and pre-sample outputs must be IDENTICAL to the compile-time-flag product kernel; 1100 steps
    wrap every history ring with d <= 256."""
    hp = teacher_hp
    eng, _ = make_engine(hp)
    rng = np.random.default_rng(23)
    T = 1100
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    for k in ('NSW_FASTGEN_GENERIC', 'NSW_FASTGEN_FLAGS', 'NSW_FASTGEN_L2LAST', 'NSW_FASTGEN_DEBUG'):
        monkeypatch.delenv(k, raising=False)
    a0, o0 = eng.run_host(enc, seed=9, want_out=True)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    a1, o1 = eng.run_host(enc, seed=9, want_out=True)
    assert np.array_equal(a0, a1)
    assert np.array_equal(o0, o1)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('T', [1, 2, 3, 5, 33])
def test_very_short_utterances_match_oracle(teacher_hp, T):
    """Edge cases of the one-phase-ahead history prefetch and the weight ring: utterances shorter than the
    prefetch distance, than the first dilation cycle, and just past one ring of the d=16 layers."""
    hp = teacher_hp
    eng, w = make_engine(hp)
    rng = np.random.default_rng(100 + T)
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
    _, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    assert np.abs(out - ref).max() < TOL
