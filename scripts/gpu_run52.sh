#!/bin/bash
# mel front-end on the GPU (new), bundle-directory synthesis test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mel_gpu.py tests/test_iaf_gpu.py -m gpu -x -q --timeout 600 -s > gpurun_out/test52.log 2>&1; echo "gpu tests rc=$?"
grep -i "mel \|passed\|failed\|error" gpurun_out/test52.log | cut -c1-250 | tail -14
python - <<'PY'
import time, numpy as np, torch
from nsynth_wavenet_b200.auxilaries.mel_extractor import MelExtractor
from oracle import mel_oracle
ex = MelExtractor(0)
wav = torch.from_numpy(np.random.default_rng(0).uniform(-.5,.5,(8,160000)).astype(np.float32)).cuda()
for _ in range(3): ex.device(wav)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ex.device(wav)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
t=time.perf_counter(); mel_oracle.batch_melspectrogram(wav[:1].cpu().numpy()); cpu=(time.perf_counter()-t)*8e3
print('mel 8x10s: GPU %.3f ms (%.0fx real time), NumPy oracle %.0f ms for the same batch' % (ms, 80e3/ms, cpu))
PY
