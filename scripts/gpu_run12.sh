#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py -x -q -s --timeout 900 > gpurun_out/test_teacher.log 2>&1; echo "teacher tests rc=$?"
tail -25 gpurun_out/test_teacher.log
timeout 900 python -m pytest tests/test_iaf_gpu.py tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py -x -q --timeout 600 > gpurun_out/test12.log 2>&1; echo "other gpu tests rc=$?"; tail -3 gpurun_out/test12.log
