#!/bin/bash
# state of the tc3 engine: all GPU parity tests, full bench, ncu launch list, ncu --set full of the flow kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 900 > gpurun_out/test21.log 2>&1; echo "gpu tests rc=$?"
grep -E "max-abs|errors|passed|failed|Error" gpurun_out/test21.log | tail -30
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"
cat gpurun_out/bench21.json; tail -3 gpurun_out/bench21.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches21.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu21_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iaf_flow_tc -s 4 -c 2 -o gpurun_out/prof21_flow python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu21_full.log 2>&1; echo "ncu full rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -2
