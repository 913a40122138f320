#!/bin/bash
# fastgen: 12 warps per CTA, past taps in their own warp group beside the critical section
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test42_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test42_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags default,2564:16,2560:12,2568 --debug > gpurun_out/fg42.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg42.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-330 | tail
