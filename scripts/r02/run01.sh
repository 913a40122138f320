#!/bin/bash
# r02 run01: per-layer timeline of the flow kernel at 8x7680 (K=4 first CTA, K=4 mid-clip CTA, K=3 CTA)
mkdir -p gpurun_out
timeout 600 python scripts/r02/flow_debug.py 0 2 10 > gpurun_out/r02_flow_debug.log 2>&1
echo rc=$?
grep -c "flow_tc dbg" gpurun_out/r02_flow_debug.log
