#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gn_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r02_test18_gn.log 2>&1; echo "gn tests rc=$?"
tail -2 gpurun_out/r02_test18_gn.log
timeout 900 python scripts/r02/fastgen_batched_bench.py mol:gn:8 mol:gn:1 mol:gn:4 ce:gn:8 ce:gn:1 2>&1 | cut -c1-260
