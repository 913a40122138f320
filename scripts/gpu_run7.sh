#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py -x -q --timeout 600 > gpurun_out/test7.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/summary7.txt
tail -5 gpurun_out/test7.log
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
for flags in (0, 1, 4, 5, 0):
    os.environ['NSW_FASTGEN_FLAGS'] = str(flags)
    eng.run_device(enc, seed=2); torch.cuda.synchronize()
    print('flags', flags, 'seq' if flags & 1 else 'conc', 'volatile' if flags & 2 else 'relaxed.gpu', '1rep' if flags & 4 else '8rep',
          'us/step %.2f' % (eng.last_timing() * 1e3 / 8000), flush=True)
os.environ['NSW_FASTGEN_DEBUG'] = '1'
for flags in (0, 5):
    os.environ['NSW_FASTGEN_FLAGS'] = str(flags)
    eng.run_device(enc, seed=2); torch.cuda.synchronize()
PY
timeout 600 python /tmp/fg_ab.py > gpurun_out/fastgen_ab7.log 2>&1; echo "fastgen ab rc=$?"; tail -16 gpurun_out/fastgen_ab7.log
timeout 600 python bench.py --engine tc2 --steps 10 --warmup 3 --no-fastgen > gpurun_out/bench7_tc2.json 2> gpurun_out/bench7_tc2.err; echo "bench tc2 rc=$?" | tee -a gpurun_out/summary7.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench7_tc2.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('stage',d['stage_ms']); print('roofline',d['roofline']['launch_ms'],d['roofline']['frac'])
PY
