#!/bin/bash
# fastgen: lean compile-time-flag build vs the run-time-switch build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test46_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test46_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags generic,default,generic,default > gpurun_out/fg46.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg46.log | cut -c1-200 | tail -8
