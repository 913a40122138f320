#!/bin/bash
# r02 run34: whole GPU suite + smoke + bench after the conv-GEMM / teacher changes
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_test34.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r02_test34.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench34.json 2> gpurun_out/r02_bench34.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench34.json'))
print('value',d['value'],'ms',d['ms_per_step'],'sustained',d['sustained']['ms_per_step_median'],'e2e',d['e2e']['value'],'py',d['e2e_python'].get('value'))
print('stages',d['stage_ms'])
print('roofline',{k:d['roofline'][k] for k in ('bound','frac','frac_model_hbm','frac_dram')})
print('fastgen',d['fastgen'].get('rtf'),d['fastgen'].get('batched',{}).get('value'),d['fastgen'].get('ce_double_gate_batched',{}).get('value'),d['fastgen'].get('e2e',{}).get('value'))
print('distill',d['distill'].get('ms'),d['distill'].get('teacher_forward_ms'),'clarinet',d['clarinet'].get('value'),'cpu',d['cpu_baseline']['value'])
PY
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches34_teacher.csv python scripts/r02/teacher_only.py > /dev/null 2>&1; echo "list rc=$?"
