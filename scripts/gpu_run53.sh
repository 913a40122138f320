#!/bin/bash
# fastgen: on-arrival contraction in the poll group (phases 2..L publish without the CTA barrier / reload / long dot)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastgen_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/test53_fastgen.log 2>&1; echo "fastgen tests rc=$?"
tail -3 gpurun_out/test53_fastgen.log
timeout 600 python scripts/fastgen_exp.py --steps 32000 --flags default,generic,2560:0,default > gpurun_out/fg53.log 2>&1
echo rc=$?
grep -v "^$" gpurun_out/fg53.log | cut -c1-200 | tail -6
timeout 300 python scripts/fastgen_exp.py --steps 8000 --flags generic --debug 2>&1 | grep "cta   0\|flags" | cut -c1-460
