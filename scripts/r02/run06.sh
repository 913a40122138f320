#!/bin/bash
mkdir -p gpurun_out
for f in 0 1; do
echo "== NSW_GN_FLAGS=$f"
NSW_GN_FLAGS=$f NSW_FASTGEN_DEBUG=1 T=2000 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:1 mol:gn:8 ce:gn:8 > gpurun_out/r02_fastgen_gn_dbg3_$f.log 2>&1; echo rc=$?
grep -v "^$" gpurun_out/r02_fastgen_gn_dbg3_$f.log | grep -E "cta   0|cta  64|case" | cut -c1-260 | tail -12
NSW_GN_FLAGS=$f T=4000 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:1 mol:gn:8 mol:gn:8 ce:gn:8 2>&1 | cut -c1-200
done
