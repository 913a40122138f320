#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/probes/umma_rate
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iaf_flow_tc -s 4 -c 1 -o gpurun_out/prof18_flow python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/ncu18_full.log 2>&1; echo "ncu full rc=$?"
