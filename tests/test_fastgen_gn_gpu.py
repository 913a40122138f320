"""GPU parity of the batched autoregressive engine (nsw_fastgen_gn.cu) and of the samplers of both fastgen engines.

* pre-sample tensor out[B,T,O] under teacher forcing vs the oracle, within 1e-4 (BASELINE north_star), for
  gate 512 (wavenet_mol.json), batch rows 1..9 (dead lanes, two weight passes), chunked conditioning, and the
  double-gate mu-law CE model (wavenet_ce.json: gate 1024, 256-way head);
* samplers with SUPPLIED noise (nsw_fastgen_set_noise): the int32 sample of every free-running step must equal
  loss_func.mol_sample / gauss_sample / ce_sample (oracle restatements, loss_func.py:140-206) evaluated on the
  kernel's own pre-sample tensor and the same draws.  "Equal" is bit-exact except where float64 arithmetic puts the
  value within rounding distance of a decision boundary (a 1/32768 bin edge, a Gumbel-max tie, a CDF step): there
  either neighbour is accepted, the number of such steps is bounded and printed."""
import os
from argparse import Namespace

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from conftest import GOLDEN_DIR, load_hparams

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_engine(hp, engine='ffma', seed=12345, bias_std=0.02, edit=None):
    from nsynth_wavenet_b200 import FastgenEngine
    w = O.init_teacher_weights(hp, seed=seed, bias_std=bias_std)
    if edit:
        edit(w)
    return FastgenEngine(hp, w, device=0, engine=engine), w


@pytest.fixture
def gn(monkeypatch):
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', 'gn')
    monkeypatch.delenv('NSW_FASTGEN_CHUNK', raising=False)


@pytest.mark.timeout(300)
def test_gn_teacher_forced_matches_golden_and_latency_engine(teacher_hp, monkeypatch):
    g = np.load(os.path.join(GOLDEN_DIR, 'fastgen_tf_1x96.npz'))
    eng, _ = make_engine(teacher_hp)
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', 'latency')
    _, o_lat = eng.run_host(g['enc'], teacher_force=g['wav'], want_out=True)
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', 'gn')
    audio, o_gn = eng.run_host(g['enc'], teacher_force=g['wav'], want_out=True)
    print('gn vs golden', np.abs(o_gn - g['out']).max(), 'gn vs latency engine', np.abs(o_gn - o_lat).max(),
          'ms', eng.last_timing())
    assert np.abs(o_gn - g['out']).max() < TOL
    assert np.abs(o_gn - o_lat).max() < 2e-5
    assert np.array_equal(audio, g['wav'])


@pytest.mark.timeout(600)
@pytest.mark.parametrize('B', [2, 3, 8, 9])
def test_gn_batch_rows_match_oracle(teacher_hp, gn, B):
    """2 and 8 fill the batch template exactly, 3 leaves a dead lane, 9 needs a second weight pass."""
    eng, w = make_engine(teacher_hp, engine='tc')
    rng = np.random.default_rng(40 + B)
    T = 40
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (B, T)).astype(np.float32)
    audio, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, teacher_hp, enc, np.float32, teacher_force=wav)['out']
    err = np.abs(out - ref).max()
    print('B', B, 'err', err)
    assert err < TOL and np.array_equal(audio, wav)


@pytest.mark.timeout(600)
def test_gn_chunked_conditioning_carries_state_across_launches(teacher_hp, gn, monkeypatch):
    """Chunk boundaries (37 steps) cut through ring wrap-arounds (2d+1 <= 129 for d <= 64 inside 300 steps) and
    through the conv_start queues; the chunked run must be bit-identical to the unchunked one."""
    eng, w = make_engine(teacher_hp)
    rng = np.random.default_rng(50)
    B, T = 2, 300
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    a0, o0 = eng.run_host(enc, seed=11, want_out=True)
    monkeypatch.setenv('NSW_FASTGEN_CHUNK', '37')
    eng2, _ = make_engine(teacher_hp)
    a1, o1 = eng2.run_host(enc, seed=11, want_out=True)
    assert np.array_equal(a0, a1) and np.array_equal(o0, o1)
    ref = O.fastgen_run(w, teacher_hp, enc, np.float32, teacher_force=a0)['out']
    assert np.abs(o0 - ref).max() < TOL


@pytest.mark.timeout(600)
def test_ce_double_gate_mu_law_model_matches_oracle():
    """config_jsons/wavenet_ce.json as shipped: 30 layers, gate 1024 (double_gate_width default), mu-law input,
    256-way categorical head (wavenet.py:106,117-122,411-414)."""
    hp = load_hparams('wavenet_ce.json')
    eng, w = make_engine(hp, seed=77)
    assert eng.out_width == 256
    rng = np.random.default_rng(60)
    B, T = 2, 48
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    codes = rng.integers(-128, 128, (B, T))
    wav = O.inv_mu_law(codes)                      # what fastgen.synthesis feeds back (fastgen.py:163-164)
    audio, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    err = np.abs(out - ref).max()
    print('ce double-gate teacher-forced err', err, 'ms', eng.last_timing())
    assert out.shape == (B, T, 256) and err < TOL
    a = eng.run_host(enc, seed=3)                  # free running: every sample is an inverse mu-law code
    codes_back = O.mu_law(a.astype(np.float64))
    assert np.all(np.abs(O.inv_mu_law(codes_back) - a) < 1e-6) and codes_back.min() >= -128 and codes_back.max() <= 127


def _check_mol(out, audio, u1, u2, Q=65536):
    """-> (#steps that differ from the float64 restatement, #of those not explained by a rounding tie)."""
    out = out.astype(np.float64)
    nr = out.shape[-1] // 3
    v = out[..., :nr] - np.log(-np.log(u1.astype(np.float64)))
    srt = np.sort(v, axis=-1)
    gap = srt[..., -1] - srt[..., -2]
    sel = v.argmax(-1)
    mu = np.take_along_axis(out[..., nr:2 * nr], sel[..., None], -1)[..., 0]
    ls = np.clip(np.take_along_axis(out[..., 2 * nr:], sel[..., None], -1)[..., 0], -7.0, 7.0)
    nz = np.log(u2.astype(np.float64)) - np.log1p(-u2.astype(np.float64))
    x = mu + np.exp(ls) * nz
    y = np.clip(x, -1.0, 1.0 - 2.0 / Q) * (Q / 2)
    q = np.floor(y)
    got = np.round(audio.astype(np.float64) * (Q / 2))
    bad = got != q
    tol = (Q / 2) * (2e-6 * (np.abs(mu) + np.abs(np.exp(ls) * nz)) + 1e-7)
    near_edge = np.abs(y - np.round(y)) <= tol
    tie = gap < 1e-5
    unexplained = bad & ~((near_edge & (np.abs(got - q) <= 1)) | tie)
    return int(bad.sum()), int(unexplained.sum()), {'clip_hi': int((x >= 1.0 - 2.0 / Q).sum()), 'clip_lo': int((x <= -1.0).sum()),
                                                   'ls_hi': int((ls >= 7.0).sum()), 'ls_lo': int((ls <= -7.0).sum())}


def _edge_biases(w):
    """Push the head into every clip branch of mol_sample: log-scales past +-7 (loss_func.py:175-177), means past
    +-1 so that x leaves [-1, 1 - 2/Q] (:183)."""
    b = w['out2/biases'].copy()
    nr = b.shape[0] // 3
    b[nr + 0] += 1.5; b[nr + 1] -= 1.5                      # means beyond +-1
    b[2 * nr + 2] += 9.0; b[2 * nr + 3] -= 9.0              # log-scales beyond +-7
    b[2 * nr + 4:3 * nr] -= 4.0                             # the rest: narrow components, samples near the means
    w['out2/biases'] = b


@pytest.mark.timeout(900)
@pytest.mark.parametrize('which', ['latency', 'gn'])
def test_mol_sampler_is_bit_exact_against_the_oracle_on_supplied_noise(teacher_hp, monkeypatch, which):
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', which)
    eng, w = make_engine(teacher_hp, edit=_edge_biases)
    rng = np.random.default_rng(70)
    B, T = (1, 2200) if which == 'latency' else (3, 800)    # 2200 > 2*1024+1: every ring wraps while free-running
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    u = rng.uniform(1e-5, 1 - 1e-5, (B, T, 11)).astype(np.float32)
    eng.set_noise(u)
    audio, out = eng.run_host(enc, seed=1, want_out=True)
    a2, _ = eng.run_host(enc, seed=999, want_out=True)      # supplied noise: the seed is irrelevant
    assert np.array_equal(audio, a2)
    nbad, nunexpl, cover = _check_mol(out, audio, u[..., :10], u[..., 10])
    print(which, 'steps', B * T, 'differ from float64 restatement', nbad, 'unexplained', nunexpl, cover)
    assert nunexpl == 0 and nbad <= 0.02 * B * T
    assert min(cover.values()) > 0, cover                   # every clip branch was taken
    # and the trajectory the kernel followed is the oracle's: same outputs when the oracle is fed the kernel's samples
    ref = O.fastgen_run(w, teacher_hp, enc[:, :400], np.float32, teacher_force=audio[:, :400])['out']
    assert np.abs(out[:, :400] - ref).max() < TOL
    # free-running oracle with the same draws, step-locked: identical samples until the first rounding tie
    fr = O.fastgen_run(w, teacher_hp, enc[:1, :64], np.float32, u1=u[:1, :64, :10], u2=u[:1, :64, 10])['audio']
    agree = (fr == audio[:1, :64])
    first = int(np.argmin(agree)) if not agree.all() else 64
    print('free-running oracle agrees for the first', first, 'of 64 steps')
    assert first >= 8
    eng.set_noise(None)
    a3 = eng.run_host(enc[:, :50], seed=1)
    assert not np.array_equal(a3, audio[:, :50])            # back on Philox


@pytest.mark.timeout(600)
@pytest.mark.parametrize('which', ['latency', 'gn'])
def test_gauss_sampler_is_bit_exact_against_the_oracle_on_supplied_noise(teacher_hp, monkeypatch, which):
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', which)
    hp = Namespace(**{**vars(teacher_hp), 'loss_type': 'gauss'})

    def edit(w):
        w['out2/biases'] = w['out2/biases'] + np.asarray([0.3, -2.0], np.float32)   # std ~ e^-2: in range, some clips
    eng, w = make_engine(hp, seed=7, edit=edit)
    rng = np.random.default_rng(71)
    B, T = (1, 1200) if which == 'latency' else (2, 600)
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    n = (rng.standard_normal((B, T, 1)) * 3).astype(np.float32)   # x3: reaches both clip edges
    eng.set_noise(n)
    audio, out = eng.run_host(enc, want_out=True)
    o = out.astype(np.float64)
    x = o[..., 0] + np.exp(np.maximum(o[..., 1], -7.0)) * n[..., 0]
    y = np.clip(x, -1.0, 1.0 - 2.0 / 65536) * 32768
    got = np.round(audio.astype(np.float64) * 32768)
    bad = got != np.floor(y)
    near = np.abs(y - np.round(y)) <= 32768 * (2e-6 * (np.abs(o[..., 0]) + np.abs(x - o[..., 0])) + 1e-7)
    print(which, 'gauss: differ', int(bad.sum()), 'of', B * T, 'clipped', int((x <= -1).sum()), int((x >= 1 - 2 / 65536).sum()))
    assert not np.any(bad & ~(near & (np.abs(got - np.floor(y)) <= 1)))
    assert (x <= -1).sum() > 0 and (x >= 1 - 2 / 65536).sum() > 0


@pytest.mark.timeout(600)
def test_ce_sampler_is_bit_exact_against_the_oracle_on_supplied_noise():
    hp = Namespace(**{**vars(load_hparams('wavenet_ce.json')), 'num_layers': 10})

    def edit(w):
        w['out2/W'] = w['out2/W'] * 6.0     # peaked categorical: the CDF has large and tiny steps
    eng, w = make_engine(hp, seed=9, edit=edit)
    rng = np.random.default_rng(72)
    B, T = 2, 500
    enc = rng.uniform(-1, 1, (B, T, 256)).astype(np.float32)
    u = rng.uniform(0, 1, (B, T, 1)).astype(np.float32)
    eng.set_noise(u)
    audio, out = eng.run_host(enc, want_out=True)
    o = out.astype(np.float64)
    p = np.exp(o - o.max(-1, keepdims=True))
    cdf = np.cumsum(p, -1)
    thr = u[..., 0].astype(np.float64) * cdf[..., -1]
    k = O.ce_sample(out.reshape(B * T, 256), 256, u.reshape(B * T)).reshape(B, T) + 128
    got = O.mu_law(audio.astype(np.float64)) + 128          # audio = inv_mu_law(code): recover the code
    bad = got != k
    # a differing step must sit on a CDF step within fp32 rounding of u * total
    lo = np.take_along_axis(cdf, np.minimum(got, k).astype(np.int64)[..., None], -1)[..., 0]
    near = np.abs(lo - thr) <= 3e-6 * cdf[..., -1]
    print('ce: differ', int(bad.sum()), 'of', B * T, 'distinct codes', len(np.unique(got)))
    assert not np.any(bad & ~(near & (np.abs(got - k) <= 1)))
    assert len(np.unique(got)) > 20
    ref = O.fastgen_run(w, hp, enc[:, :64], np.float32, teacher_force=audio[:, :64])['out']
    assert np.abs(out[:, :64] - ref).max() < TOL


@pytest.mark.timeout(300)
def test_mu_law_input_on_the_latency_engine(teacher_hp, monkeypatch):
    """mu-law input encoding + inverse mu-law feedback with a MoL head (hparams the reference allows,
    wavenet.py:117-125,411-414): quant_chann 256."""
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', 'latency')
    hp = Namespace(**{**vars(teacher_hp), 'use_mu_law': True})
    eng, w = make_engine(hp, seed=3)
    rng = np.random.default_rng(73)
    T = 80
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    wav = O.inv_mu_law(rng.integers(-128, 128, (1, T)))
    _, out = eng.run_host(enc, teacher_force=wav, want_out=True)
    ref = O.fastgen_run(w, hp, enc, np.float32, teacher_force=wav)['out']
    assert np.abs(out - ref).max() < TOL
    a = eng.run_host(enc, seed=5)
    codes = O.mu_law(a.astype(np.float64))
    assert np.all(np.abs(O.inv_mu_law(codes) - a) < 1e-6)


@pytest.mark.timeout(300)
def test_philox_gaussian_draws_have_untruncated_normal_moments(teacher_hp, monkeypatch):
    """Device-drawn N(0,1) of the Gaussian head (Box-Muller on an unclipped uniform): with out2 frozen to
    (mean 0, log std log 0.05) the samples ARE the noise; moments and the 4-sigma tail mass of 40 000 draws."""
    monkeypatch.setenv('NSW_FASTGEN_ENGINE', 'latency')
    hp = Namespace(**{**vars(teacher_hp), 'loss_type': 'gauss'})

    def edit(w):
        w['out2/W'] = np.zeros_like(w['out2/W'])
        w['out2/biases'] = np.asarray([0.0, np.log(0.05)], np.float32)
    eng, _ = make_engine(hp, seed=7, edit=edit)
    T = 40000
    enc = np.zeros((1, T, 256), np.float32)
    n = eng.run_host(enc, seed=123)[0] / 0.05
    m, s = n.mean(), n.std()
    kurt = ((n - m) ** 4).mean() / s ** 4
    tail = (np.abs(n) > 3.0).mean()
    print('gauss draws: mean', m, 'std', s, 'kurtosis', kurt, 'P(|n|>3)', tail, 'max', np.abs(n).max())
    assert abs(m) < 0.02 and abs(s - 1) < 0.02 and abs(kurt - 3) < 0.15 and 0.0015 < tail < 0.004
