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
for flags in (0, 5, 0, 5):
    os.environ['NSW_FASTGEN_FLAGS'] = str(flags)
    eng.run_device(enc, seed=2); torch.cuda.synchronize()
    print('flags', flags, 'nostream' if flags & 8 else 'stream', 'us/step %.2f' % (eng.last_timing() * 1e3 / 8000), flush=True)
os.environ['NSW_FASTGEN_DEBUG'] = '1'
os.environ["NSW_FASTGEN_FLAGS"] = "0"
eng.run_device(enc, seed=2); torch.cuda.synchronize()
PY
timeout 600 python /tmp/fg_ab.py > gpurun_out/fastgen_ab8.log 2>&1; echo "fastgen ab rc=$?"; tail -10 gpurun_out/fastgen_ab8.log
timeout 600 python -m pytest tests/test_fastgen_gpu.py -x -q --timeout 600 2>&1 | tail -2
