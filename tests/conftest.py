import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG_DIR = os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons')
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_hparams(name):
    from oracle import wavenet_oracle as O
    return O.load_hparams(os.path.join(CONFIG_DIR, name))


@pytest.fixture(scope='session')
def student_hp():
    return load_hparams('parallel_wavenet.json')


@pytest.fixture(scope='session')
def clarinet_hp():
    return load_hparams('parallel_wavenet_gauss.json')


@pytest.fixture(scope='session')
def teacher_hp():
    return load_hparams('wavenet_mol.json')


def synth_inputs(hp, B, F, seed=54321, gauss=False):
    """mel ~ U[0,1), z = logistic from u ~ U[1e-5, 1-1e-5] (or N(0,1))."""
    from oracle import wavenet_oracle as O
    rng = np.random.default_rng(seed)
    mel = rng.uniform(0, 1, (B, F, 80)).astype(np.float32)
    T = O.iaf_length(F, hp)
    if gauss:
        z = rng.standard_normal((B, T)).astype(np.float32)
    else:
        u = rng.uniform(1e-5, 1 - 1e-5, (B, T))
        z = O.logistic_from_uniform(u).astype(np.float32)
    return mel, z
