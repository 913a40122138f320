#!/bin/bash
# r02 run49: teacher residual/skip epilogue with 64-column chunks (whole-line stores of the split pair)
timeout 900 python -m pytest tests/test_teacher_gpu.py tests/test_distill_gpu.py tests/test_trained_regime_gpu.py -k "teacher or distill" -m gpu -q -s --timeout 300 2>&1 | grep "max-abs err\|trained-regime err\|passed\|failed" | cut -c1-200 | tail -8
REPS=5 python scripts/r02/teacher_only.py
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches49_teacher.csv python scripts/r02/teacher_only.py > /dev/null 2>&1; echo "list rc=$?"
