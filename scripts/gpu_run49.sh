#!/bin/bash
# full GPU suite + smoke + default bench + launch list on the fastgen-1.38x tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/test49.log 2>&1; echo "gpu tests rc=$?"
tail -2 gpurun_out/test49.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench49.json 2> gpurun_out/bench49.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench49.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'fastgen',d['fastgen'].get('rtf'),d['fastgen'].get('us_per_step'),'distill',d['distill'].get('ms'),'cpu',d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench49_ref.json 2> gpurun_out/bench49_ref.err; echo "ref rc=$?"; cat gpurun_out/bench49_ref.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches49.csv python bench.py --steps 2 --warmup 1 --no-distill --fastgen-steps 512 > gpurun_out/ncu49.log 2>&1; echo "ncu rc=$?"
