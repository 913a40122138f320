#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_teacher_gpu.py tests/test_distill_gpu.py tests/test_cli_dropin.py tests/test_mel_gpu.py tests/test_iaf_tc_gpu.py -m gpu -q -s --timeout 900 > gpurun_out/r02_test09.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|tf stft|power loss|calculate_loss" gpurun_out/r02_test09.log | cut -c1-400 | tail -20
