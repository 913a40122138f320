#!/bin/bash
# fastgen exchange experiments: baseline, pipelined polling, timing-only switches, with cycle counters
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,64,8,128,384,392,456 --debug > gpurun_out/fg34.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg34.log | tail -50
