#!/bin/bash
# r02 run35: teacher with the early-release epilogue of the split-accumulator mode
run() {
echo "== $*"
env "$@" timeout 600 python -m pytest tests/test_teacher_gpu.py -k "teacher" -m gpu -q -s --timeout 600 2>&1 | grep "max-abs err\|passed\|failed" | cut -c1-200
env "$@" REPS=5 python scripts/r02/teacher_only.py
}
run NSW_TEACHER_SPLIT_ACC=1
run NSW_TEACHER_SPLIT_ACC=0
