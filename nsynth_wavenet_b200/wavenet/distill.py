"""Forward half of the teacher-student distillation step (BASELINE configs[4]), with the reference's names:
ParallelWavenet.kl_loss_logistic / kl_loss_gauss / power_loss / contrastive_loss / calculate_loss
(wavenet/parallel_wavenet.py:361-510) on top of the three device engines.  Everything stays on the GPU; the losses
come back as Python floats.

The reference builds this as part of its TF training graph; here it is the scoring path only (no gradients): what
`train_parallel_wavenet.py` logs every step, and what an evaluation of a student against its teacher needs."""
from __future__ import annotations

from ..auxilaries import mel_extractor
from ..engine import IAFEngine, TeacherEngine


class DistillForward:
    """student (IAFEngine) + teacher (TeacherEngine) + power-loss STFT on one GPU."""

    def __init__(self, student_hparams, student_weights, teacher_hparams, teacher_weights, device=0, engine=None):
        self.hparams = student_hparams
        self.loss_type = getattr(student_hparams, 'loss_type', 'logistic')     # parallel_wavenet.py:129
        self.student = IAFEngine(student_hparams, student_weights, device=device, engine=engine)
        self.teacher = TeacherEngine(teacher_hparams, teacher_weights, device=device)
        self.stft = mel_extractor.TfStft(device)
        want = 'mol' if self.loss_type == 'logistic' else 'gauss'
        if teacher_hparams.loss_type != want:
            raise ValueError("a '{}' student is scored by a '{}' teacher (parallel_wavenet.py:492-501), got '{}'".format(
                self.loss_type, want, teacher_hparams.loss_type))

    def close(self):
        for e in (self.student, self.teacher, self.stft):
            e.close()

    def feed_forward(self, mel, z=None, seed=0):
        """ParallelWavenet.feed_forward (parallel_wavenet.py:289-345): torch CUDA mel [B,F,80] -> ff_dict."""
        out = self.student.forward_device(mel, z, seed=seed, quantize=False)
        out['mel'] = mel
        return out

    def kl_loss_logistic(self, ff_dict, num_samples=100, eps=None, seed=0, mel_key='mel'):
        """parallel_wavenet.py:361-402 (CLIP = False: the teacher sees x unclipped)."""
        te_out = self.teacher.forward_device(ff_dict['x'], ff_dict[mel_key])
        return self.teacher.mol_score(te_out, ff_dict['mean_tot'], ff_dict['scale_tot'], ff_dict['log_scale_tot'],
                                      num_samples=num_samples, eps=eps, seed=seed)

    def kl_loss_gauss(self, ff_dict):
        """parallel_wavenet.py:404-428."""
        te_out = self.teacher.forward_device(ff_dict['x'], ff_dict['mel'])
        return {'kl_loss': self.teacher.gauss_kl(te_out, ff_dict['mean_tot'], ff_dict['scale_tot'],
                                                 ff_dict['log_scale_tot'])['kl_loss']}

    def power_loss(self, wav_dict):
        """parallel_wavenet.py:459-479: wav_dict['x'] (student output) against wav_dict['wav'] (ground truth)."""
        return {'power_loss': self.stft.power_loss(wav_dict['wav'], wav_dict['x'])['power_loss']}

    def contrastive_loss(self, ff_dict, num_samples=100, eps=None, seed=0):
        """parallel_wavenet.py:481-490: minus the KL against the teacher conditioned on ANOTHER clip's mel."""
        return {'contrastive_loss': -self.kl_loss_logistic(ff_dict, num_samples, eps=eps, seed=seed,
                                                           mel_key='mel_rand')['kl_loss']}

    def calculate_loss(self, ff_dict, eps=None, seed=0):
        """parallel_wavenet.py:492-510."""
        plf = self.hparams.power_loss_factor
        if self.loss_type == 'logistic':
            clf = getattr(self.hparams, 'contrastive_loss_factor', 0.0)
            num_samples = getattr(self.hparams, 'num_samples', 0)
            loss_dict = dict(self.kl_loss_logistic(ff_dict, num_samples, eps=eps, seed=seed))
        else:
            clf, num_samples = 0.0, 0
            loss_dict = dict(self.kl_loss_gauss(ff_dict))
        loss = loss_dict['kl_loss']
        if plf > 0.0:
            pl = self.power_loss(ff_dict)
            loss += plf * pl['power_loss']
            loss_dict.update(pl)
        if clf > 0.0:
            cl = self.contrastive_loss(ff_dict, num_samples, eps=eps, seed=seed)
            loss += clf * cl['contrastive_loss']
            loss_dict.update(cl)
        loss_dict['loss'] = loss
        return loss_dict
