#!/bin/bash
# fastgen persistent kernel bring-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -x -q -s --timeout 600 > gpurun_out/test_fastgen.log 2>&1; echo "fastgen tests rc=$?" | tee -a gpurun_out/summary2.txt
tail -25 gpurun_out/test_fastgen.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
