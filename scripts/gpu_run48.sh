#!/bin/bash
# fastgen: 16 / 32 replicas (lean builds); memcheck of the default kernel on a short run
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,18944:16,35328:16,default > gpurun_out/fg48.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg48.log | cut -c1-200 | tail -6
timeout 600 compute-sanitizer --tool memcheck python scripts/fastgen_exp.py --steps 192 --flags default > gpurun_out/memcheck48.log 2>&1
echo "memcheck rc=$?"
grep -i "error summary\|flags" gpurun_out/memcheck48.log | tail -3
