#!/bin/bash
# r02 run41: accumulate source through N = 64 identity MMAs (4 KB identity piece per k-block instead of 32 KB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_teacher_gpu.py tests/test_distill_gpu.py -m gpu -q -s --timeout 300 2>&1 | grep "max-abs err\|passed\|failed" | cut -c1-200 | tail -24
REPS=5 python scripts/r02/teacher_only.py
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches41_teacher.csv python scripts/r02/teacher_only.py > /dev/null 2>&1; echo "list rc=$?"
