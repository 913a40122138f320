#!/bin/bash
# tc2 (tcgen05 residual layers) bring-up + fastgen cycle breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py -x -q -s --timeout 300 > gpurun_out/test_tc2.log 2>&1; echo "tc/tc2 tests rc=$?" | tee -a gpurun_out/summary3.txt
tail -30 gpurun_out/test_tc2.log
timeout 600 python bench.py --engine tc2 --steps 10 --warmup 3 --no-fastgen > gpurun_out/bench3_tc2.json 2> gpurun_out/bench3_tc2.err; echo "bench tc2 rc=$?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench3_tc2.json; tail -3 gpurun_out/bench3_tc2.err
NSW_FASTGEN_DEBUG=1 timeout 600 python - > gpurun_out/fastgen_dbg.log 2>&1 <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from argparse import Namespace
from nsynth_wavenet_b200 import FastgenEngine
from oracle import wavenet_oracle as O
hp = Namespace(**json.load(open('nsynth_wavenet_b200/config_jsons/wavenet_mol.json')))
w = O.init_teacher_weights(hp, seed=12345)
eng = FastgenEngine(hp, w, device=0)
enc = (torch.rand((1, 16000, 256)) * 2 - 1).cuda()
eng.run_device(enc[:, :2000], seed=1); torch.cuda.synchronize()
eng.run_device(enc, seed=2); torch.cuda.synchronize()
print('ms', eng.last_timing(), 'us/step', eng.last_timing() * 1e3 / 16000)
PY
echo "fastgen dbg rc=$?" | tee -a gpurun_out/summary3.txt; cat gpurun_out/fastgen_dbg.log | tail -12
