#!/bin/bash
# fastgen: 8 (default) vs 4 vs 2 replicas of the exchange slots on the final kernel
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,18944:16,35328:16,2564:16,default > gpurun_out/fg57.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg57.log | cut -c1-200 | tail -6
