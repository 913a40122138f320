#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/fg_ab.py <<'PY'
import sys, os, json, numpy as np, torch
sys.path.insert(0, '.')
from argparse import Namespace
from nsynth_wavenet_b200 import FastgenEngine
from oracle import wavenet_oracle as O
hp = Namespace(**json.load(open('nsynth_wavenet_b200/config_jsons/wavenet_mol.json')))
w = O.init_teacher_weights(hp, seed=12345)
eng = FastgenEngine(hp, w, device=0)
enc = (torch.rand((1, 8000, 256)) * 2 - 1).cuda()
eng.run_device(enc[:, :1000], seed=1); torch.cuda.synchronize()
for flags in (0, 1, 2, 3, 4, 5, 6, 7, 0):
    os.environ['NSW_FASTGEN_FLAGS'] = str(flags)
    eng.run_device(enc, seed=2); torch.cuda.synchronize()
    print('flags', flags, 'seq' if flags & 1 else 'conc', 'volatile' if flags & 2 else 'relaxed.gpu', '1rep' if flags & 4 else '8rep',
          'us/step %.2f' % (eng.last_timing() * 1e3 / 8000), flush=True)
PY
timeout 600 python /tmp/fg_ab.py > gpurun_out/fastgen_ab.log 2>&1; echo "fastgen ab rc=$?"; cat gpurun_out/fastgen_ab.log | tail -12
cat > /tmp/one_fwd.py <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from argparse import Namespace
from nsynth_wavenet_b200 import IAFEngine
from oracle import wavenet_oracle as O
hp = Namespace(**json.load(open('nsynth_wavenet_b200/config_jsons/parallel_wavenet.json')))
w = O.init_student_weights(hp, seed=12345)
eng = IAFEngine(hp, w, device=0, engine=sys.argv[1])
mel = torch.rand((8, 39, 80), device='cuda')
for i in range(2):
    eng.forward_device(mel, None, seed=i)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_layer_tc_kernel -s 70 -c 1 -o gpurun_out/prof_layer_tc_v3 python /tmp/one_fwd.py tc2 > gpurun_out/ncu_layer_tc_v3.log 2>&1; echo "ncu rc=$?"
