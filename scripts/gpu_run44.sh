#!/bin/bash
# fastgen: how long is one poll sweep, how many sweeps per phase
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 2560:16,2560:16:1500,2564:16,2568 --debug > gpurun_out/fg44.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg44.log | grep -v "cta   1\|cta 127" | cut -c1-430 | tail -14
