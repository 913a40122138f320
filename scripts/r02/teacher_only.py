"""Teacher full-sequence forward alone (configs[4] shape: 7 x 7680), for ncu captures and A/B timing of the conv-GEMM."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench import load_hparams
from nsynth_wavenet_b200 import TeacherEngine
from nsynth_wavenet_b200.weights_init import init_teacher_weights

reps = int(os.environ.get('REPS', '3'))
thp = load_hparams('wavenet_mol.json')
te = TeacherEngine(thp, init_teacher_weights(thp, seed=12345), device=0)
g = torch.Generator(device='cpu').manual_seed(1)
mel = torch.rand((7, 39, 80), generator=g).cuda()
wav = (torch.rand((7, 7680), generator=g) * 2 - 1).cuda()
ms = []
for i in range(reps):
    out = te.forward_device(wav, mel)
    torch.cuda.synchronize()
    ms.append(te.last_timing())
print(json.dumps({'teacher_forward_ms': ms, 'out_abs_max': float(out.abs().max())}))
