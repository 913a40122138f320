#!/bin/bash
# fastgen: register-resident cycle counters (no global RMW stalls), early vs late preload of the critical rows
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,2560:16,6656:16,2568,6664 --debug > gpurun_out/fg41.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg41.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-300 | tail
