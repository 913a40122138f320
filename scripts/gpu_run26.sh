#!/bin/bash
mkdir -p gpurun_out
for v in "" 1; do
export NSW_FLOW_NOCOOP=$v
if [ -z "$v" ]; then unset NSW_FLOW_NOCOOP; fi
timeout 300 python bench.py --steps 20 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/bench26_$v.json 2> gpurun_out/bench26_$v.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench26_$v.json'))
    print('nocoop="$v" value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'], 'stage',d['stage_ms'])
except Exception as e: print('no bench', e)
PY
done
