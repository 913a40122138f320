#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py -x -q -s --timeout 600 > gpurun_out/test23.log 2>&1; echo "tc tests rc=$?"
grep -E "tc3 vs|passed|failed|Error|error|watchdog" gpurun_out/test23.log | tail -20
