#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -x -q -s -m gpu --timeout 900 > gpurun_out/test13.log 2>&1; echo "gpu tests rc=$?"
grep -E "max-abs|errors|err|passed|failed|mol score|distillation" gpurun_out/test13.log | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-fastgen > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench13.json'))
print('engine',d['config']['engine'],'value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'])
print('stage',d['stage_ms'])
PY
