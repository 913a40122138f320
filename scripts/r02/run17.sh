#!/bin/bash
mkdir -p gpurun_out
NSW_FASTGEN_DEBUG=1 REPS=1 T=2000 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:8 ce:gn:8 mol:gn:1 > gpurun_out/r02_fastgen_gn_dbg17.log 2>&1; echo rc=$?
grep -E "cta   0|case" gpurun_out/r02_fastgen_gn_dbg17.log | cut -c1-250
