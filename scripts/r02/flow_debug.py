"""Per-layer timeline of iaf_flow_tc_kernel at the benchmark shape (NSW_LAYER_DEBUG counters)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import load_hparams  # noqa: E402
from nsynth_wavenet_b200 import IAFEngine  # noqa: E402
from nsynth_wavenet_b200.weights_init import init_student_weights  # noqa: E402

hp = load_hparams('student')
eng = IAFEngine(hp, init_student_weights(hp, seed=12345), device=0)
B, F = int(os.environ.get('B', 8)), int(os.environ.get('F', 39))
mel = torch.rand((B, F, 80), device='cuda')
for i in range(3):
    eng.forward_device(mel, None, seed=i)
torch.cuda.synchronize()
for cta in [int(c) for c in sys.argv[1:]] or [0]:
    os.environ['NSW_LAYER_DEBUG'] = '1'
    os.environ['NSW_LAYER_DEBUG_CTA'] = str(cta)
    sys.stderr.write('==== cta %d ====\n' % cta)
    eng.forward_device(mel, None, seed=7)
    torch.cuda.synchronize()
del os.environ['NSW_LAYER_DEBUG']
