#!/bin/bash
# fastgen: critical-row weights preloaded into registers during the barrier wait
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test40_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test40_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,default,2568 --debug > gpurun_out/fg40.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg40.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-300 | tail
