#!/bin/bash
# fastgen with the new defaults (red.max publish, bulk history prefetch, L2 eviction hints): parity tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test38_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -3 gpurun_out/test38_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags 0,default > gpurun_out/fg38.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg38.log | tail
