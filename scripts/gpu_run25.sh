#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py -x -q -s --timeout 300 -k "tc3" > gpurun_out/test25.log 2>&1; echo "tc3 tests rc=$?"
grep -E "max-abs|tc3 vs|passed|failed|Error|error|watchdog" gpurun_out/test25.log | tail -14
for v in "" 1; do
NSW_COND_NOCLUSTER=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/bench25_$v.json 2> gpurun_out/bench25_$v.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench25_$v.json'))
    print('nocluster="$v" value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'], 'stage',d['stage_ms'])
except Exception as e: print('no bench', e)
PY
tail -3 gpurun_out/bench25_$v.err
done
