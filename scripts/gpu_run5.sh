#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py -x -q --timeout 600 > gpurun_out/test5.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/summary5.txt
tail -5 gpurun_out/test5.log
timeout 600 python bench.py --engine tc2 --steps 10 --warmup 3 > gpurun_out/bench5_tc2.json 2> gpurun_out/bench5_tc2.err; echo "bench tc2 rc=$?" | tee -a gpurun_out/summary5.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5_tc2.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('stage',d['stage_ms']); print('roofline',d['roofline']['launch_ms'],d['roofline']['frac'])
print('fastgen',d.get('fastgen'))
PY
tail -3 gpurun_out/bench5_tc2.err
