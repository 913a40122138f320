#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -s --timeout 300 2>&1 | grep -v "^$" | tail -25
