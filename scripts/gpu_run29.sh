#!/bin/bash
# compute-sanitizer memcheck over the tc3 engine on the small parity cases
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_iaf_tc_gpu.py -x -q -s --timeout 800 -k "tc3 and (golden_config1 or clarinet)" > gpurun_out/sanitizer29.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer29.log | tail -8
