#!/bin/bash
# r02 run63 (= run48 with the pair flow kernel as default): final state: whole GPU suite, smoke, default bench, reference arm, launch list,
# ncu --set full of the flow / cond kernels, of the teacher's pair GEMMs and of the batched fastgen kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_test63.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/r02_test63.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench63.json 2> gpurun_out/r02_bench63.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench63.json'))
print('value',d['value'],'ms',d['ms_per_step'],'sustained',d['sustained']['ms_per_step_median'],'e2e',d['e2e']['value'],'py',d['e2e_python'].get('value'))
print('stages',d['stage_ms'])
print('roofline',{k:d['roofline'][k] for k in ('bound','frac','frac_model_hbm','frac_dram')})
print('fastgen',d['fastgen'].get('rtf'),d['fastgen'].get('batched',{}).get('value'),d['fastgen'].get('ce_double_gate_batched',{}).get('value'),d['fastgen'].get('e2e',{}).get('value'))
print('distill',d['distill'].get('ms'),d['distill'].get('teacher_forward_ms'),'clarinet',d['clarinet'].get('value'),'cpu',d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench63_ref.json 2>/dev/null; echo "ref rc=$?"
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-sustained --no-python-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches63.csv python bench.py --steps 2 --warmup 1 $LEAN > gpurun_out/r02_ncu63_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum \
  -k regex:"iaf_flow|cond_proj|conv_gemm" -s 12 -c 6 -o gpurun_out/r02_prof63 python bench.py --steps 1 --warmup 1 $LEAN > gpurun_out/r02_ncu63_full.log 2>&1; echo "ncu full rc=$?"
REPS=2 timeout 600 ncu --set full --clock-control none -k regex:"conv_gemm_tc2" -s 68 -c 4 -o gpurun_out/r02_prof63_teacher python scripts/r02/teacher_only.py > gpurun_out/r02_ncu63_teacher.log 2>&1; echo "ncu teacher rc=$?"
ls -la gpurun_out/*63*.ncu-rep
