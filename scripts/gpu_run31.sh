#!/bin/bash
# end-of-round state of engine tc3 (fused start conv + head, CTA-pair cond projection): GPU tests, smoke, bench,
# launch list, ncu --set full of the two dominant kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 900 > gpurun_out/test31.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|Error" gpurun_out/test31.log | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench31.json 2> gpurun_out/bench31.err; echo "bench rc=$?"
cat gpurun_out/bench31.json; tail -3 gpurun_out/bench31.err
timeout 600 python bench.py --config clarinet --steps 10 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/bench31_clarinet.json 2> gpurun_out/bench31_clarinet.err; echo "clarinet bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches31.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu31_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"iaf_flow_tc|cond_proj" -s 2 -c 4 -o gpurun_out/prof31 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu31_full.log 2>&1; echo "ncu full rc=$?"
