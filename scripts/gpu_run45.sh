#!/bin/bash
# fastgen: past taps gated on the publish (default) vs ungated (8192); no proxy fence before the history prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test45_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test45_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags default,10752:16,2568 --debug > gpurun_out/fg45.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg45.log | grep -v "cta   1\|cta 127\|cta  64" | cut -c1-460 | tail -14
