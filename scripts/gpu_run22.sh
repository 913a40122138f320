#!/bin/bash
# timing experiments on the flow kernel (results are numerically wrong by construction): which traffic costs time?
mkdir -p gpurun_out
for e in 0 32 64 128 256 512 96 480 495 1007; do
  NSW_FLOW_EXP=$e timeout 300 python bench.py --steps 10 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/exp22_$e.json 2> gpurun_out/exp22_$e.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exp22_$e.json'))
    print('exp $e: ms_per_step %.4f layers %.4f cond %.4f' % (d['ms_per_step'], d['stage_ms']['layers'], d['stage_ms']['cond']))
except Exception as ex: print('exp $e failed', ex)
PY
done
