#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gn_gpu.py -m gpu -x -q -s --timeout 600 > gpurun_out/r02_test_gn2.log 2>&1; echo "gn rc=$?"
tail -3 gpurun_out/r02_test_gn2.log
NSW_FASTGEN_DEBUG=1 T=2000 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:1 mol:gn:8 ce:gn:8 > gpurun_out/r02_fastgen_gn_dbg2.log 2>&1; echo rc=$?
grep -v "^$" gpurun_out/r02_fastgen_gn_dbg2.log | grep -E "cta   0|case" | cut -c1-260 | tail -12
timeout 600 python scripts/r02/fastgen_batched_bench.py > gpurun_out/r02_fastgen_batched2.log 2>&1; echo rc=$?
cut -c1-200 gpurun_out/r02_fastgen_batched2.log
