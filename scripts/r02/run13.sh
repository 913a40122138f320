#!/bin/bash
# r02 run13: what bounds cond_proj_tc_kernel?  stage time with the global stores removed / with 1/12 of the MMAs
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for e in 0 1 2; do
NSW_COND_EXP=$e timeout 600 python bench.py --steps 20 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_COND_EXP=$e ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done
