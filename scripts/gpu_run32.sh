#!/bin/bash
# final sanity of the committed tree: all GPU tests, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/test32.log 2>&1; echo "gpu tests rc=$?"
tail -2 gpurun_out/test32.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench32.json 2> gpurun_out/bench32.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench32.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'fastgen',d['fastgen'].get('rtf'),'distill',d['distill'].get('ms'))
PY
