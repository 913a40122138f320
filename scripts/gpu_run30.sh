#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py -x -q --timeout 600 -k "odd_shapes" > gpurun_out/test30.log 2>&1; echo "odd-shape tests rc=$?"
tail -5 gpurun_out/test30.log
