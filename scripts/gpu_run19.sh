#!/bin/bash
mkdir -p gpurun_out
NSW_LAYER_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > /dev/null 2> gpurun_out/dbg19.err; grep -A6 "flow_tc dbg" gpurun_out/dbg19.err | tail -28
