#!/bin/bash
# ncu --set full capture of the final fastgen kernel (2048 samples)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastgen_kernel -c 1 -f -o gpurun_out/fastgen_ncu_run56 python scripts/fastgen_exp.py --steps 2048 --flags default > gpurun_out/ncu56.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu56.log
