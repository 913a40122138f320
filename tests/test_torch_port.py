"""The torch-CPU baseline port agrees with the numpy oracle (so the CPU number bench.py
reports is for the same arithmetic the GPU path is checked against)."""
import numpy as np

from oracle import torch_port, wavenet_oracle as O
from conftest import synth_inputs


def test_student_port_matches_oracle(student_hp):
    w = O.init_student_weights(student_hp, seed=12345, bias_std=0.02)
    mel, z = synth_inputs(student_hp, 2, 6)
    ref = O.parallelgen_forward(w, student_hp, mel, z, np.float64)
    got = torch_port.StudentPort(w, student_hp).forward(mel, z)
    for k in ('mean_tot', 'scale_tot', 'log_scale_tot'):
        assert np.abs(got[k] - ref[k]).max() < 2e-5, k
    assert np.abs(got['x'] - ref['x']).max() <= 1.0 / 32768 + 1e-6


def test_clarinet_port_matches_oracle(clarinet_hp):
    w = O.init_student_weights(clarinet_hp, seed=3, bias_std=0.02)
    mel, z = synth_inputs(clarinet_hp, 1, 6, gauss=True)
    ref = O.student_feed_forward(w, clarinet_hp, mel, z, np.float64)
    got = torch_port.StudentPort(w, clarinet_hp).forward(mel, z, quantize=False)
    for k in ('mean_tot', 'scale_tot', 'log_scale_tot', 'x'):
        assert np.abs(got[k] - ref[k]).max() < 2e-5, k


def test_fastgen_port_step_matches_oracle(teacher_hp):
    w = O.init_teacher_weights(teacher_hp, seed=12345, bias_std=0.02)
    rng = np.random.default_rng(0)
    T = 12
    enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
    wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
    ref = O.fastgen_run(w, teacher_hp, enc, np.float64, teacher_force=wav)['out']
    port = torch_port.FastgenPort(w, teacher_hp, 1)
    import torch
    x = torch.zeros(1, 1)
    for i in range(T):
        out = port.step(x, torch.from_numpy(enc[:, i]))
        assert np.abs(out.numpy() - ref[:, i]).max() < 2e-5, i
        x = torch.from_numpy(wav[:, i:i + 1])
