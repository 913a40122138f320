#!/bin/bash
# r02 run32: teacher: split accumulators (hi*hi | small products) on/off x cond fused/separate: error vs fp64 and time
mkdir -p gpurun_out
for sa in 0 1; do for cs in 0 1; do
echo "== NSW_TEACHER_SPLIT_ACC=$sa NSW_TEACHER_COND_SEPARATE=$cs"
NSW_TEACHER_SPLIT_ACC=$sa NSW_TEACHER_COND_SEPARATE=$cs timeout 600 python -m pytest tests/test_teacher_gpu.py -k "teacher" -m gpu -q -s --timeout 600 2>&1 | grep "max-abs err\|trained-regime err\|passed\|failed" | cut -c1-200
NSW_TEACHER_SPLIT_ACC=$sa NSW_TEACHER_COND_SEPARATE=$cs REPS=5 python scripts/r02/teacher_only.py
done; done
