#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/test59.log 2>&1; echo "gpu tests rc=$?"
tail -2 gpurun_out/test59.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench59.json 2> gpurun_out/bench59.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench59.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'fastgen',d['fastgen'].get('rtf'),d['fastgen'].get('us_per_step'),'distill',d['distill'].get('ms'),'cpu',d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
PY
timeout 300 compute-sanitizer --tool memcheck python scripts/fastgen_exp.py --steps 192 --flags default > gpurun_out/memcheck59.log 2>&1; echo "memcheck rc=$?"; grep -i "error summary" gpurun_out/memcheck59.log
