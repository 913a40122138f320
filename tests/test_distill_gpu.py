"""The rest of BASELINE configs[4]: power-loss STFT (mel_extractor.py:111-121), power loss
(parallel_wavenet.py:459-479), contrastive term (:481-490) and the combined ParallelWavenet.calculate_loss (:492-510),
each against the oracle with shared noise."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as MO
from oracle import wavenet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(300)
def test_tf_stft_magnitudes_match_oracle():
    from nsynth_wavenet_b200.auxilaries.mel_extractor import TfStft
    st = TfStft(0)
    rng = np.random.default_rng(1)
    for B, N in ((1, 7680), (3, 1530), (2, 200), (2, 801), (1, 154480)):
        y = (0.3 * rng.standard_normal((B, N))).astype(np.float32)
        got = st.stft_mag(torch.from_numpy(y).cuda()).cpu().numpy()
        ref = np.abs(MO.tf_stft(y))
        assert got.shape == ref.shape == (B, -(-N // 200), 1025)
        err = np.abs(got - ref).max()
        print('tf stft', B, N, 'max-abs err', err, 'max |S|', ref.max())
        assert err < 1e-4 * max(1.0, ref.max() / 50)          # fp32 contraction of 800 taps, |S| up to ~30


@pytest.mark.timeout(300)
def test_power_loss_matches_oracle_including_the_centre_crop():
    from nsynth_wavenet_b200.auxilaries.mel_extractor import TfStft, PRIORITY_FREQ
    st = TfStft(0)
    assert PRIORITY_FREQ == MO.PRIORITY_FREQ == 384
    rng = np.random.default_rng(2)
    for (B, No, Np) in ((7, 7680, 7680), (2, 7800, 7680), (2, 7680, 7745), (1, 1000, 1000)):
        o = (0.3 * rng.standard_normal((B, No))).astype(np.float32)
        p = (0.3 * rng.standard_normal((B, Np))).astype(np.float32)
        got = st.power_loss(torch.from_numpy(o).cuda(), torch.from_numpy(p).cuda())
        ref = MO.power_loss(o, p)
        print('power loss', B, No, Np, got['power_loss'], ref)
        assert abs(got['power_loss'] - ref) < 2e-5 * ref
        assert abs(got['power_loss'] - (0.5 * got['all_bins'] + 0.5 * got['priority_bins'])) < 1e-12
    same = st.power_loss(torch.from_numpy(o).cuda(), torch.from_numpy(o).cuda())
    assert same['power_loss'] == 0.0


@pytest.mark.timeout(900)
def test_calculate_loss_logistic_kl_power_and_contrastive_terms(student_hp, teacher_hp):
    """ParallelWavenet.calculate_loss for parallel_wavenet.json (power_loss_factor 1.0, contrastive_loss_factor 0.3,
    num_samples 100) at reduced batch, with the logistic draws of the KL shared with the oracle."""
    from nsynth_wavenet_b200.wavenet.distill import DistillForward
    sw = O.init_student_weights(student_hp, seed=12345)
    tw = O.init_teacher_weights(teacher_hp, seed=12345, bias_std=0.02)
    df = DistillForward(student_hp, sw, teacher_hp, tw, device=0)
    rng = np.random.default_rng(3)
    B, F = 2, 6
    mel = rng.uniform(0, 1, (B, F, 80)).astype(np.float32)
    T = df.student.length(F)
    z = O.logistic_from_uniform(rng.uniform(1e-5, 1 - 1e-5, (B, T))).astype(np.float32)
    wav = (0.2 * rng.standard_normal((B, F * 200))).astype(np.float32)       # ground truth: longer than T, cropped
    S = student_hp.num_samples
    eps = O.logistic_from_uniform(rng.uniform(1e-5, 1 - 1e-5, (S, B, T))).astype(np.float32)
    ff = df.feed_forward(torch.from_numpy(mel).cuda(), torch.from_numpy(z).cuda())
    ff['wav'] = torch.from_numpy(wav).cuda()
    ff['mel_rand'] = torch.from_numpy(mel[::-1].copy()).cuda()               # the other clip's conditioning
    got = df.calculate_loss(ff, eps=torch.from_numpy(eps).cuda())
    # oracle, on the engine's own student outputs (the student forward has its own parity tests)
    x, mt = ff['x'].cpu().numpy(), ff['mean_tot'].cpu().numpy()
    sc, ls = ff['scale_tot'].cpu().numpy(), ff['log_scale_tot'].cpu().numpy()
    te = O.teacher_feed_forward(tw, teacher_hp, x, mel, np.float32)['out_params']
    te_r = O.teacher_feed_forward(tw, teacher_hp, x, mel[::-1], np.float32)['out_params']
    kl = O.kl_loss_logistic(te, mt, sc, ls, eps, 65536)
    cl = -O.kl_loss_logistic(te_r, mt, sc, ls, eps, 65536)['kl_loss']
    pl = MO.power_loss(wav, x)
    ref_loss = kl['kl_loss'] + student_hp.power_loss_factor * pl + student_hp.contrastive_loss_factor * cl
    print('calculate_loss', got, {'kl': kl, 'power_loss': pl, 'contrastive_loss': cl, 'loss': ref_loss})
    assert abs(got['kl_loss'] - kl['kl_loss']) < 2e-4 * abs(kl['kl_loss'])
    assert abs(got['H_Ps'] - kl['H_Ps']) < 1e-4
    assert abs(got['power_loss'] - pl) < 2e-5 * pl
    assert abs(got['contrastive_loss'] - cl) < 2e-4 * abs(cl)
    assert abs(got['loss'] - ref_loss) < 2e-4 * abs(ref_loss)
    assert got['contrastive_loss'] != -got['kl_loss']                          # the permuted mel changes the teacher
    df.close()


@pytest.mark.timeout(600)
def test_calculate_loss_gauss_student_has_no_contrastive_term(clarinet_hp):
    from nsynth_wavenet_b200.wavenet.distill import DistillForward
    thp = O.load_hparams(os.path.join(os.path.dirname(__file__), '..', 'nsynth_wavenet_b200', 'config_jsons',
                                      'wavenet_gauss.json'))
    sw = O.init_student_weights(clarinet_hp, seed=12345)
    tw = O.init_teacher_weights(thp, seed=12345, bias_std=0.02)
    df = DistillForward(clarinet_hp, sw, thp, tw, device=0)
    rng = np.random.default_rng(4)
    mel = rng.uniform(0, 1, (2, 6, 80)).astype(np.float32)
    ff = df.feed_forward(torch.from_numpy(mel).cuda(), None, seed=3)
    ff['wav'] = torch.from_numpy((0.2 * rng.standard_normal((2, 1024))).astype(np.float32)).cuda()
    got = df.calculate_loss(ff)
    assert sorted(got) == ['kl_loss', 'loss', 'power_loss']
    x = ff['x'].cpu().numpy()
    te = O.teacher_feed_forward(tw, thp, x, mel, np.float32)['out_params']
    kl = O.kl_loss_gauss(te, ff['mean_tot'].cpu().numpy(), ff['scale_tot'].cpu().numpy(), ff['log_scale_tot'].cpu().numpy())
    pl = MO.power_loss(ff['wav'].cpu().numpy(), x)
    assert abs(got['kl_loss'] - kl['kl_loss']) < 2e-3 * max(1.0, abs(kl['kl_loss']))
    assert abs(got['power_loss'] - pl) < 2e-5 * pl
    assert abs(got['loss'] - (got['kl_loss'] + clarinet_hp.power_loss_factor * got['power_loss'])) < 1e-9
    from nsynth_wavenet_b200.weights_init import init_teacher_weights
    mol = O.load_hparams(os.path.join(os.path.dirname(__file__), '..', 'nsynth_wavenet_b200', 'config_jsons', 'wavenet_mol.json'))
    with pytest.raises(ValueError, match="'gauss' teacher"):
        DistillForward(clarinet_hp, sw, mol, init_teacher_weights(mol, seed=1), device=0)
    df.close()
