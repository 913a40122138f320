#!/bin/bash
# r02 run33: teacher: residual/skip accumulate through identity k-blocks (no fp32 master), fast gate; error vs fp64 and time
mkdir -p gpurun_out
run() {
echo "== $*"
env "$@" timeout 600 python -m pytest tests/test_teacher_gpu.py -k "teacher" -m gpu -q -s --timeout 600 2>&1 | grep "max-abs err\|passed\|failed" | cut -c1-200
env "$@" REPS=5 python scripts/r02/teacher_only.py
}
run NSW_TEACHER_SPLIT_ACC=0
run NSW_TEACHER_SPLIT_ACC=1
run NSW_TEACHER_COND_SEPARATE=1
run NSW_TEACHER_COND_SEPARATE=1 NSW_TEACHER_FP32_MASTER=1
run NSW_TEACHER_FP32_MASTER=1
