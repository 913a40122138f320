#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/r02/fastgen_batched_bench.py > gpurun_out/r02_fastgen_batched.log 2>&1; echo rc=$?
cat gpurun_out/r02_fastgen_batched.log | cut -c1-300
