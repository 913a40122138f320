#!/bin/bash
# full GPU suite + default bench on the fastgen-1.15x tree, then one ncu capture of fastgen_kernel (2048 samples)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/test39.log 2>&1; echo "gpu tests rc=$?"
tail -2 gpurun_out/test39.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench39.json 2> gpurun_out/bench39.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench39.json'))
print('value',d['value'],'rtf',d['rtf'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'fastgen',d['fastgen'],'distill',d['distill'].get('ms'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastgen_kernel -c 1 -f -o gpurun_out/fastgen_ncu_run39 python scripts/fastgen_exp.py --steps 2048 --flags default > gpurun_out/ncu39.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu39.log
ls -la gpurun_out/*.ncu-rep
