#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -x -q -s --timeout 600 > gpurun_out/test14.log 2>&1; echo "tests rc=$?"
grep -E "errors|max-abs|passed|failed|Error|error" gpurun_out/test14.log | tail -16
timeout 600 python bench.py --steps 10 --warmup 3 --no-fastgen --no-distill > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench14.json'))
print('engine',d['config']['engine'],'value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'])
print('stage',d['stage_ms']); print('layer launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'])
PY
tail -3 gpurun_out/bench14.err
