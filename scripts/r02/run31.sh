#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py tests/test_distill_gpu.py tests/test_trained_regime_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/r02_test31.log 2>&1; echo "tests rc=$?"
grep -n "max-abs err\|trained-regime err\|passed\|failed" gpurun_out/r02_test31.log | cut -c1-250
NSW_TEACHER_COND_SEPARATE=1 timeout 600 python -m pytest tests/test_teacher_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -n "max-abs err\|passed\|failed" | cut -c1-250
