#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_fwd.py <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from argparse import Namespace
from nsynth_wavenet_b200 import IAFEngine
from oracle import wavenet_oracle as O
hp = Namespace(**json.load(open("nsynth_wavenet_b200/config_jsons/parallel_wavenet.json")))
w = O.init_student_weights(hp, seed=12345)
eng = IAFEngine(hp, w, device=0, engine=sys.argv[1])
mel = torch.rand((8, 39, 80), device="cuda")
for i in range(2):
    eng.forward_device(mel, None, seed=i)
torch.cuda.synchronize()
PY
NSW_LAYER_DEBUG=1 python /tmp/one_fwd.py tc2 2>&1 | tail -40 > gpurun_out/layer_dbg.log; cat gpurun_out/layer_dbg.log
