#!/bin/bash
# first GPU bring-up: parity (fp32 engine, then tcgen05 engine), short bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests/test_iaf_gpu.py -x -q --timeout 900 > gpurun_out/test_ffma.log 2>&1; echo "ffma tests rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py -q --timeout 300 -s > gpurun_out/test_tc.log 2>&1; echo "tc tests rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --engine ffma --steps 5 --warmup 3 > gpurun_out/bench_ffma.json 2> gpurun_out/bench_ffma.err; echo "bench ffma rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --engine tc --steps 10 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench tc rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --engine tc --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/test_ffma.log; tail -15 gpurun_out/test_tc.log; cat gpurun_out/bench_ffma.json gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
