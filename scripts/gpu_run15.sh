#!/bin/bash
# re-entry baseline of HEAD: all GPU parity tests, full bench, ncu launch list, ncu --set full of the layer kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi15.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 900 > gpurun_out/test15.log 2>&1; echo "gpu tests rc=$?"
grep -E "max-abs|errors|passed|failed|Error" gpurun_out/test15.log | tail -25
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"
cat gpurun_out/bench15.json; tail -3 gpurun_out/bench15.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches15.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu15_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_layer_tc -s 4 -c 2 -o gpurun_out/prof15_layer python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu15_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
