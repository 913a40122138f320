#!/bin/bash
# fastgen with 2 replicas as the default: parity tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test58_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test58_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,35328:16,2564:16,default > gpurun_out/fg58.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg58.log | cut -c1-200 | tail -6
