#!/bin/bash
# fastgen: incremental ring counters, finer slack counters; replica / L2-hint sweep around the best point
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,2560:16,2564:16,2564:14,2564:18,2560:10,2568 --debug > gpurun_out/fg37.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg37.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-300 | tail -40
