"""Throughput of the two fastgen engines: latency engine (B = 1) and batched engine at B = 1, 2, 4, 8, for
wavenet_mol.json (gate 512) and wavenet_ce.json (gate 1024, 256-way head).  Prints one JSON line per case."""
import json
import os
import sys
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nsynth_wavenet_b200 import FastgenEngine  # noqa: E402
from nsynth_wavenet_b200.weights_init import init_teacher_weights  # noqa: E402

T = int(os.environ.get('T', 6000))
cases = sys.argv[1:] or ['mol:latency:1', 'mol:gn:1', 'mol:gn:2', 'mol:gn:4', 'mol:gn:8', 'ce:gn:1', 'ce:gn:8']
engines = {}
for case in cases:
    cfg, which, B = case.split(':')
    B = int(B)
    if cfg not in engines:
        with open(os.path.join(ROOT, 'nsynth_wavenet_b200', 'config_jsons', 'wavenet_%s.json' % cfg)) as f:
            hp = Namespace(**json.load(f))
        engines[cfg] = FastgenEngine(hp, init_teacher_weights(hp, seed=12345), device=0)
    eng = engines[cfg]
    os.environ['NSW_FASTGEN_ENGINE'] = which
    g = torch.Generator(device='cpu').manual_seed(1)
    enc = (torch.rand((B, T, 256), generator=g) * 2 - 1).cuda()
    eng.run_device(enc[:, :512].contiguous(), seed=1)
    torch.cuda.synchronize()
    times = []
    for rep in range(int(os.environ.get('REPS', 3))):
        eng.run_device(enc, seed=2 + rep)
        torch.cuda.synchronize()
        times.append(eng.last_timing())
    ms = sorted(times)[len(times) // 2]
    print(json.dumps({'case': case, 'B': B, 'T': T, 'ms': ms, 'all_ms': [round(x, 1) for x in times], 'us_per_step': 1e3 * ms / T,
                      'samples_per_s': B * T / (ms * 1e-3), 'rtf_aggregate': B * T / (ms * 1e-3) / 16000}), flush=True)
