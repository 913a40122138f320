#!/bin/bash
# r02 run07: new bench.py (all blocks), launch list, ncu --set full (+ shared-memory / pipe counters) of the flow and
# cond kernels at the benchmark shape
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench07.json 2> gpurun_out/r02_bench07.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/r02_bench07.json; tail -5 gpurun_out/r02_bench07.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench07_ref.json 2>/dev/null; echo "ref rc=$?"
cut -c1-400 gpurun_out/r02_bench07_ref.json
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-sustained --no-python-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches07.csv python bench.py --steps 2 --warmup 1 $LEAN > gpurun_out/r02_ncu07_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum \
  -k regex:"iaf_flow_tc|cond_proj" -s 8 -c 5 -o gpurun_out/r02_prof07 python bench.py --steps 1 --warmup 1 $LEAN > gpurun_out/r02_ncu07_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/r02_ncu07_full.log
ls -la gpurun_out/r02_prof07.ncu-rep
