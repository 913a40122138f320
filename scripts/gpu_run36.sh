#!/bin/bash
# fastgen: red.max publish + history prefetch (2560) with one replica (+4), delayed polling, L2 eviction hints on the weight stream
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,2560,2564,2560:0:500,2560:0:900,2560:12,2560:16,2560:20,2560:24,2560:32,2568 --debug > gpurun_out/fg36.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg36.log | grep -v "cta   1\|cta  64\|cta 127" | cut -c1-260 | tail -40
