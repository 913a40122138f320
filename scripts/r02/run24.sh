#!/bin/bash
# r02 run24: cond projection with 256-column cta_group::2 work items vs the 128-column shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2; do
for tn in 256 128; do
NSW_COND_TN=$tn timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_COND_TN=$tn ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
