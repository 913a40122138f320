#!/bin/bash
# first run of engine tc3 (smem-resident flow kernel): parity tests, then a short bench + timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py -x -q -s --timeout 300 -k "tc3" > gpurun_out/test17.log 2>&1; echo "tc3 tests rc=$?"
grep -E "max-abs|errors|passed|failed|Error|error|watchdog|assert" gpurun_out/test17.log | tail -30
timeout 300 python bench.py --steps 10 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench17.json'))
    print('engine',d['config']['engine'],'value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'])
    print('stage',d['stage_ms']); print('layer launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'])
except Exception as e: print('no bench', e)
PY
tail -5 gpurun_out/bench17.err
NSW_LAYER_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > /dev/null 2> gpurun_out/dbg17.err; grep -A5 "flow_tc dbg" gpurun_out/dbg17.err | tail -24
