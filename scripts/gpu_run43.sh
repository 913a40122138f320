#!/bin/bash
# fastgen: is the phase bound by crit + hop?  artificial delay before the critical section / before polling
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 2560:16,2560:16:0:500,2560:16:0:1000,2560:16:1500,2560:16:2000,2564:16:1800 --debug > gpurun_out/fg43.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg43.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-330 | tail -14
