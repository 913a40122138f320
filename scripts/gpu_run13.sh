#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests/test_teacher_gpu.py -x -q -s --timeout 900 > gpurun_out/test13.log 2>&1; echo "gpu tests rc=$?"
grep -E "max-abs|errors|err|passed|failed|mol score|distillation" gpurun_out/test13.log | tail -30
