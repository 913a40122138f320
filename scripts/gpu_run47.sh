#!/bin/bash
# fastgen: ablations with compile-time-flag builds (no dead code in any of them)
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,0:0,512:0,2048:0,2560:0,2560:8,2560:12,2560:16,2560:20,2564:16,10752:16,6656:16,generic > gpurun_out/fg47.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg47.log | cut -c1-200 | tail -14
