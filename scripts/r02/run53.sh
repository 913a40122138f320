#!/bin/bash
# r02 run53: compute-sanitizer memcheck over the pair conv-GEMM (hook tests), the upsampling stack and one teacher forward
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x --timeout 900 > gpurun_out/r02_sanitizer53_gemm.log 2>&1; echo "memcheck conv_gemm rc=$?"
tail -4 gpurun_out/r02_sanitizer53_gemm.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_iaf_gpu.py -m gpu -q -x --timeout 900 -k "deconv or resize or trans_conv or config1 or golden" > gpurun_out/r02_sanitizer53_iaf.log 2>&1; echo "memcheck iaf rc=$?"
tail -4 gpurun_out/r02_sanitizer53_iaf.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_teacher_gpu.py -m gpu -q -x --timeout 900 -k "small" > gpurun_out/r02_sanitizer53_teacher.log 2>&1; echo "memcheck teacher rc=$?"
tail -4 gpurun_out/r02_sanitizer53_teacher.log
