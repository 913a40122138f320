#!/bin/bash
# r02 run55: compute-sanitizer initcheck over the pair conv-GEMM hook tests and the upsampling stack
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 99 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x --timeout 900 > gpurun_out/r02_initcheck55_gemm.log 2>&1; echo "initcheck conv_gemm rc=$?"
grep -c "Uninitialized" gpurun_out/r02_initcheck55_gemm.log; tail -3 gpurun_out/r02_initcheck55_gemm.log
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 99 python -m pytest tests/test_iaf_gpu.py -m gpu -q -x --timeout 900 -k "deconv or resize or trans_conv" > gpurun_out/r02_initcheck55_iaf.log 2>&1; echo "initcheck iaf rc=$?"
grep -c "Uninitialized" gpurun_out/r02_initcheck55_iaf.log; tail -3 gpurun_out/r02_initcheck55_iaf.log
grep -m3 -A12 "Uninitialized" gpurun_out/r02_initcheck55_iaf.log | head -40
