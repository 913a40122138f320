#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/test11.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/summary11.txt
tail -5 gpurun_out/test11.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-fastgen > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?" | tee -a gpurun_out/summary11.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench11.json'))
print('engine',d['config']['engine'],'value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('stage',d['stage_ms']); print('roofline',d['roofline']['launch_ms'],d['roofline']['frac'], 'gemm', d['roofline_cond_gemm']['achieved'])
PY
tail -3 gpurun_out/bench11.err
