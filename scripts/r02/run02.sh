#!/bin/bash
# r02 run02: new parity tests (batched fastgen engine, samplers on supplied noise, trained regime, entry points)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gn_gpu.py -m gpu -x -q -s --timeout 600 > gpurun_out/r02_test_gn.log 2>&1; echo "gn rc=$?"
tail -3 gpurun_out/r02_test_gn.log
timeout 900 python -m pytest tests/test_trained_regime_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/r02_test_trained.log 2>&1; echo "trained rc=$?"
tail -3 gpurun_out/r02_test_trained.log
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q -s --timeout 600 > gpurun_out/r02_test_fastgen.log 2>&1; echo "fastgen rc=$?"
tail -3 gpurun_out/r02_test_fastgen.log
