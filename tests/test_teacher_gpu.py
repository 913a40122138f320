"""GPU parity of the teacher full-sequence forward (a14) and the distillation cross-entropy
(a15) against the CPU oracle.  Tolerance 1e-4 on out_params against the fp64 oracle (split-fp16 tensor cores; the fp32 twin of the oracle is printed
next to it for scale)."""
import os

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_teacher(hp, seed=12345):
    from nsynth_wavenet_b200 import TeacherEngine
    w = O.init_teacher_weights(hp, seed=seed, bias_std=0.02)
    return TeacherEngine(hp, w, device=0), w


@pytest.mark.timeout(600)
def test_teacher_forward_matches_oracle_small(teacher_hp):
    hp = teacher_hp
    eng, w = make_teacher(hp)
    rng = np.random.default_rng(31)
    mel = rng.uniform(0, 1, (1, 3, 80)).astype(np.float32)       # 600 cond steps, trimmed to 512
    wav = rng.uniform(-0.5, 0.5, (1, 512)).astype(np.float32)
    out = eng.forward_host(wav, mel)
    ref = O.teacher_feed_forward(w, hp, wav, mel, np.float64)['out_params']
    err = np.abs(out - ref).max()
    print('teacher forward max-abs err', err, 'ms', eng.last_timing())
    assert out.shape == (1, 512, 30) and err < TOL, err


@pytest.mark.timeout(900)
def test_teacher_forward_batch_and_deep_dilations(teacher_hp):
    hp = teacher_hp
    eng, w = make_teacher(hp, seed=7)
    rng = np.random.default_rng(32)
    mel = rng.uniform(0, 1, (2, 8, 80)).astype(np.float32)       # 1600 -> 1536 samples (d up to 512 twice)
    wav = rng.uniform(-0.5, 0.5, (2, 1536)).astype(np.float32)
    out = eng.forward_host(wav, mel)
    ref = O.teacher_feed_forward(w, hp, wav, mel, np.float64)['out_params']
    twin = O.teacher_feed_forward(w, hp, wav, mel, np.float32)['out_params']
    err = np.abs(out - ref).max()
    print('teacher forward 2x1536 max-abs err vs fp64', err, ' fp32 twin vs fp64', np.abs(twin - ref).max())
    assert err < TOL, err
    # batch rows are independent
    single = eng.forward_host(wav[1:2], mel[1:2])
    assert np.abs(single[0] - out[1]).max() < 1e-6


@pytest.mark.timeout(600)
def test_mol_score_matches_oracle_with_shared_noise(teacher_hp):
    hp = teacher_hp
    eng, _ = make_teacher(hp)
    rng = np.random.default_rng(33)
    B, T, S = 2, 256, 7
    te = rng.normal(0, 1.0, (B, T, 30)).astype(np.float32)
    te[..., 20:] = rng.uniform(-9, -2, (B, T, 10))                # log-scales incl. values below the -7 floor
    mean = rng.uniform(-1.2, 1.2, (B, T)).astype(np.float32)      # some targets hit the +-1 edge branches
    scale = rng.uniform(0.01, 0.2, (B, T)).astype(np.float32)
    ls = np.log(scale)
    u = rng.uniform(1e-5, 1 - 1e-5, (S, B, T))
    eps = O.logistic_from_uniform(u).astype(np.float32)
    # the reference evaluates mol_log_probs in float32, where saturated sigmoids hit the 1e-12
    # clamp (loss_func.py:57-59) earlier than in float64 (14.72 vs 15.04 on this input): the
    # float32 oracle is the one to match
    ref = O.kl_loss_logistic(te, mean, scale, ls, eps, 65536)
    got = eng.mol_score(torch.from_numpy(te).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(scale).cuda(),
                        torch.from_numpy(ls).cuda(), num_samples=S, eps=torch.from_numpy(eps).cuda())
    print('mol score', got, ref)
    for k in ('H_Ps', 'H_Ps_Pt', 'kl_loss'):
        assert abs(got[k] - ref[k]) < 3e-4 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


@pytest.mark.timeout(900)
def test_distillation_forward_pipeline(student_hp, teacher_hp):
    # BASELINE config 5 (reduced batch): student forward -> teacher forward on x -> 100-sample scoring
    from nsynth_wavenet_b200 import IAFEngine
    st_w = O.init_student_weights(student_hp, seed=12345)
    st = IAFEngine(student_hp, st_w, device=0)
    te, te_w = make_teacher(teacher_hp)
    rng = np.random.default_rng(34)
    mel = torch.from_numpy(rng.uniform(0, 1, (2, 6, 80)).astype(np.float32)).cuda()
    out = st.forward_device(mel, None, seed=5, quantize=False)
    te_out = te.forward_device(out['x'], mel)
    torch.cuda.synchronize()
    x_np, mel_np = out['x'].cpu().numpy(), mel.cpu().numpy()
    ref_te = O.teacher_feed_forward(te_w, teacher_hp, x_np, mel_np, np.float32)['out_params']
    ref64 = O.teacher_feed_forward(te_w, teacher_hp, x_np, mel_np, np.float64)['out_params']
    err = np.abs(te_out.cpu().numpy() - ref64).max()
    print('teacher on student output: max-abs err vs fp64', err, ' fp32 twin vs fp64', np.abs(ref_te - ref64).max())
    assert err < TOL, err
    got = te.mol_score(te_out, out['mean_tot'], out['scale_tot'], out['log_scale_tot'], num_samples=100, seed=9)
    S = 100
    eps = O.logistic_from_uniform(rng.uniform(1e-5, 1 - 1e-5, (S, 2, 1024)))
    ref = O.kl_loss_logistic(ref_te, out['mean_tot'].cpu().numpy(), out['scale_tot'].cpu().numpy(),
                             out['log_scale_tot'].cpu().numpy(), eps.astype(np.float32), 65536)
    print('distillation losses', got, ref)
    assert abs(got['H_Ps'] - ref['H_Ps']) < 1e-4
    assert abs(got['H_Ps_Pt'] - ref['H_Ps_Pt']) < 0.02 * abs(ref['H_Ps_Pt'])   # different noise draws


@pytest.mark.timeout(600)
def test_gauss_kl_matches_oracle():
    """nsw_gauss_kl_device == oracle kl_loss_gauss (parallel_wavenet.py:404-428) on random inputs, including
    teacher log-scale parameters below the -7 floor and a length that is not a multiple of the block size."""
    hp = O.load_hparams(os.path.join(os.path.dirname(__file__), '..', 'nsynth_wavenet_b200', 'config_jsons',
                                     'wavenet_gauss.json'))
    from nsynth_wavenet_b200 import TeacherEngine
    w = O.init_teacher_weights(hp, seed=12345, bias_std=0.02)
    eng = TeacherEngine(hp, w, device=0)
    rng = np.random.default_rng(43)
    for B, T in ((1, 77), (3, 7680), (8, 61440)):
        te = np.stack([rng.normal(0, 0.3, (B, T)), rng.uniform(-9, -1, (B, T))], axis=-1).astype(np.float32)
        mean = rng.normal(0, 0.3, (B, T)).astype(np.float32)
        ls = rng.uniform(-6, -1, (B, T)).astype(np.float32)
        scale = np.exp(ls)
        ref = O.kl_loss_gauss(te, mean, scale, ls)
        got = eng.gauss_kl(torch.from_numpy(te).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(scale).cuda(),
                           torch.from_numpy(ls).cuda())
        print('gauss kl', B, T, got, ref)
        # tolerance: fp32 per-sample terms (expf/logf within 2 ulp of NumPy's), fp64 accumulation on both sides
        for k in ('kl', 'reg', 'kl_loss'):
            assert abs(got[k] - ref[k]) < 2e-5 * max(1.0, abs(ref[k])), (B, T, k, got[k], ref[k])
    # a mol teacher refuses
    from nsynth_wavenet_b200._lib import NswError
    mol_hp = O.load_hparams(os.path.join(os.path.dirname(__file__), '..', 'nsynth_wavenet_b200', 'config_jsons',
                                         'wavenet_mol.json'))
    mol = TeacherEngine(mol_hp, O.init_teacher_weights(mol_hp, seed=1), device=0)
    z = torch.zeros((1, 128), device='cuda')
    with pytest.raises(NswError):
        mol.gauss_kl(torch.zeros((1, 128, 2), device='cuda'), z, z, z)


@pytest.mark.timeout(900)
def test_clarinet_distillation_forward_pipeline():
    """ClariNet (BASELINE configs[3] model) distillation forward at reduced batch: Gaussian IAF student forward ->
    Gaussian teacher forward on x -> closed-form KL, each stage against the oracle."""
    from nsynth_wavenet_b200 import IAFEngine, TeacherEngine
    cj = os.path.join(os.path.dirname(__file__), '..', 'nsynth_wavenet_b200', 'config_jsons')
    shp = O.load_hparams(os.path.join(cj, 'parallel_wavenet_gauss.json'))
    thp = O.load_hparams(os.path.join(cj, 'wavenet_gauss.json'))
    st_w = O.init_student_weights(shp, seed=12345)
    te_w = O.init_teacher_weights(thp, seed=12345, bias_std=0.02)
    st = IAFEngine(shp, st_w, device=0)
    te = TeacherEngine(thp, te_w, device=0)
    rng = np.random.default_rng(44)
    mel = torch.from_numpy(rng.uniform(0, 1, (2, 6, 80)).astype(np.float32)).cuda()
    out = st.forward_device(mel, None, seed=5, quantize=False)
    te_out = te.forward_device(out['x'], mel)
    torch.cuda.synchronize()
    x = out['x'].cpu().numpy()
    ref_te = O.teacher_feed_forward(te_w, thp, x, mel.cpu().numpy(), np.float32)['out_params']
    assert ref_te.shape[-1] == 2
    assert np.abs(te_out.cpu().numpy() - ref_te).max() < TOL
    got = te.gauss_kl(te_out, out['mean_tot'], out['scale_tot'], out['log_scale_tot'])
    ref = O.kl_loss_gauss(ref_te, out['mean_tot'].cpu().numpy(), out['scale_tot'].cpu().numpy(),
                          out['log_scale_tot'].cpu().numpy())
    print('clarinet distillation', got, ref)
    # the KL divides by var_p ~ exp(2*log-scale): a 1e-4 difference in the teacher's log-scale moves it by ~2e-4 relative
    assert abs(got['kl_loss'] - ref['kl_loss']) < 2e-3 * max(1.0, abs(ref['kl_loss']))
