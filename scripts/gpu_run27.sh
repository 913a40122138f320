#!/bin/bash
# CUDA-graph replay A/B.  NOTE: at the time of this run the switch was NSW_NO_GRAPH (graph on by default); the committed
# code has it opt-in as NSW_USE_GRAPH=1 because the gain was 0.5 %
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -x -q --timeout 600 > gpurun_out/test27.log 2>&1; echo "iaf tests rc=$?"
tail -3 gpurun_out/test27.log
for v in "" 1; do
export NSW_NO_GRAPH=$v
if [ -z "$v" ]; then unset NSW_NO_GRAPH; fi
timeout 300 python bench.py --steps 20 --warmup 4 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/bench27_$v.json 2> gpurun_out/bench27_$v.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench27_$v.json'))
    print('nograph="$v" value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e: print('no bench', e)
PY
tail -2 gpurun_out/bench27_$v.err
done
