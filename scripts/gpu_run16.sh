#!/bin/bash
# timeline of CTA 0 of the persistent layer kernel at the bench shape
mkdir -p gpurun_out
NSW_LAYER_DEBUG=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > gpurun_out/dbg16.json 2> gpurun_out/dbg16.err; echo rc=$?
grep -A7 "layer_tc dbg" gpurun_out/dbg16.err | tail -64
