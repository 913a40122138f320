#!/bin/bash
mkdir -p gpurun_out
NSW_FASTGEN_DEBUG=1 T=2000 timeout 900 python scripts/r02/fastgen_batched_bench.py mol:gn:1 mol:gn:8 ce:gn:1 ce:gn:8 > gpurun_out/r02_fastgen_gn_dbg.log 2>&1; echo rc=$?
grep -v "^$" gpurun_out/r02_fastgen_gn_dbg.log | cut -c1-260
