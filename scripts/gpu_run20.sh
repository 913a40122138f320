#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cond_proj -s 4 -c 2 -o gpurun_out/prof20_cond python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu20_full.log 2>&1; echo "ncu full rc=$?"
