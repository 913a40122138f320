#!/bin/bash
# fastgen: red.max publish (512), bulk-copy exchange poll (1024), bulk history prefetch (2048)
mkdir -p gpurun_out
timeout 600 python scripts/fastgen_exp.py --steps 16000 --flags 0,512,1024,2048,1536,2560,3072,3584 --debug > gpurun_out/fg35.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg35.log | grep -v "cta   1\|cta  64" | cut -c1-260 | tail -40
