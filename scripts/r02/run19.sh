#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_fastgen_gpu.py tests/test_iaf_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r02_test19.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_test19.log
timeout 300 python scripts/r02/fastgen_batched_bench.py mol:latency:1 2>&1 | cut -c1-250
