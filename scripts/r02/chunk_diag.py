import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from argparse import Namespace
import json
from nsynth_wavenet_b200 import FastgenEngine
from nsynth_wavenet_b200.weights_init import init_teacher_weights
os.environ['NSW_FASTGEN_ENGINE'] = 'latency'
hp = Namespace(**json.load(open('nsynth_wavenet_b200/config_jsons/wavenet_mol.json')))
eng = FastgenEngine(hp, init_teacher_weights(hp, seed=12345, bias_std=0.02), device=0, engine='ffma')
rng = np.random.default_rng(77)
T = 300
enc = rng.uniform(-1, 1, (1, T, 256)).astype(np.float32)
wav = rng.uniform(-0.5, 0.5, (1, T)).astype(np.float32)
os.environ.pop('NSW_FASTGEN_CHUNK', None)
_, t0 = eng.run_host(enc, teacher_force=wav, want_out=True)
a0, o0 = eng.run_host(enc, seed=13, want_out=True)
for chunk in ('97', '150', '299'):
    os.environ['NSW_FASTGEN_CHUNK'] = chunk
    _, t1 = eng.run_host(enc, teacher_force=wav, want_out=True)
    a1, o1 = eng.run_host(enc, seed=13, want_out=True)
    d = np.abs(t1 - t0).max(axis=(0, 2))
    bad = np.nonzero(d > 0)[0]
    print('chunk', chunk, 'teacher-forced: first differing step', bad[:5], 'max', d.max(), 'free-running audio first diff', np.nonzero(a0[0] != a1[0])[0][:5])
    if len(bad):
        t = bad[0]
        print('   step', t, 'out diff', (t1 - t0)[0, t, :6], 'later', d[t:t + 12])
