#!/bin/bash
# ncu full captures (layer tc kernel, conv-GEMM tc kernel, ffma2 layer kernel) + fastgen breakdown
mkdir -p gpurun_out
cat > /tmp/one_fwd.py <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from argparse import Namespace
from nsynth_wavenet_b200 import IAFEngine
from oracle import wavenet_oracle as O
eng_name = sys.argv[1]
hp = Namespace(**json.load(open('nsynth_wavenet_b200/config_jsons/parallel_wavenet.json')))
w = O.init_student_weights(hp, seed=12345)
eng = IAFEngine(hp, w, device=0, engine=eng_name)
mel = torch.rand((8, 39, 80), device='cuda')
for i in range(2):
    eng.forward_device(mel, None, seed=i)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_layer_tc_kernel -s 70 -c 2 -o gpurun_out/prof_layer_tc python /tmp/one_fwd.py tc2 > gpurun_out/ncu_layer_tc.log 2>&1; echo "ncu layer_tc rc=$?" | tee -a gpurun_out/summary4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc python /tmp/one_fwd.py tc2 > gpurun_out/ncu_gemm_tc.log 2>&1; echo "ncu gemm_tc rc=$?" | tee -a gpurun_out/summary4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_layer_kernel -s 70 -c 1 -o gpurun_out/prof_layer_ffma python /tmp/one_fwd.py tc > gpurun_out/ncu_layer_ffma.log 2>&1; echo "ncu layer_ffma rc=$?" | tee -a gpurun_out/summary4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_head_kernel -s 4 -c 1 -o gpurun_out/prof_head python /tmp/one_fwd.py tc > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?" | tee -a gpurun_out/summary4.txt
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
tail -6 gpurun_out/fastgen_dbg.log; ls -la gpurun_out/*.ncu-rep
