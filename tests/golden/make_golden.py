"""Regenerates tests/golden/*.npz from the fp64 oracle (oracle/wavenet_oracle.py).

The reference itself cannot be imported here (TensorFlow 1.x absent), so these are
oracle outputs, not reference outputs: they pin the oracle against silent drift and
give the GPU parity tests fixed expected values.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import wavenet_oracle as O  # noqa: E402
from conftest import load_hparams, synth_inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def iaf_case(cfg, name, B, F, gauss):
    hp = load_hparams(cfg)
    w = O.init_student_weights(hp, seed=12345, bias_std=0.02)
    mel, z = synth_inputs(hp, B, F, gauss=gauss)
    o = O.parallelgen_forward(w, hp, mel, z, np.float64)
    np.savez_compressed(
        os.path.join(OUT, name), mel=mel, z=z,
        mean_tot=o['mean_tot'].astype(np.float32), scale_tot=o['scale_tot'].astype(np.float32),
        log_scale_tot=o['log_scale_tot'].astype(np.float32),
        x_pre_quant=o['x_pre_quant'].astype(np.float32), x=o['x'].astype(np.float32))
    print(name, 'T =', z.shape[1], 'mean|scale_tot| =', float(np.abs(o['scale_tot']).mean()))


def deconv_case():
    hp = load_hparams('parallel_wavenet.json')
    w = O.init_student_weights(hp, seed=12345, bias_std=0.02)
    rng = np.random.default_rng(7)
    mel = rng.uniform(0, 1, (2, 5, 80)).astype(np.float32)
    enc = O.deconv_stack(mel, w, hp, 'iaf_share/', np.float64)
    # a strided subset keeps the fixture small; tests index the same way
    np.savez_compressed(os.path.join(OUT, 'deconv_2x5.npz'), mel=mel,
                        enc_sub=enc[:, ::3, ::4].astype(np.float32),
                        enc_sum=enc.sum(axis=(1, 2)))
    print('deconv', enc.shape)


def fastgen_case():
    hp = load_hparams('wavenet_mol.json')
    w = O.init_teacher_weights(hp, seed=12345, bias_std=0.02)
    rng = np.random.default_rng(11)
    T = 96
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    tf_wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
    r = O.fastgen_run(w, hp, enc, np.float64, teacher_force=tf_wav)
    np.savez_compressed(os.path.join(OUT, 'fastgen_tf_1x96.npz'), enc=enc, wav=tf_wav,
                        out=r['out'].astype(np.float32))
    print('fastgen', r['out'].shape)


if __name__ == '__main__':
    iaf_case('parallel_wavenet.json', 'iaf_logistic_1x21.npz', 1, 21, False)
    iaf_case('parallel_wavenet_gauss.json', 'iaf_gauss_2x6.npz', 2, 6, True)
    deconv_case()
    fastgen_case()
