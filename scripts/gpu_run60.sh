#!/bin/bash
# fastgen: l part of the exchange read from the history ring (no second publish): parity tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test60_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -1 gpurun_out/test60_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,default > gpurun_out/fg60.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg60.log | cut -c1-200 | tail -6
