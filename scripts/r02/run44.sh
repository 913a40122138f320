#!/bin/bash
# r02 run44: cond projection: one GEMM for all flows (NSW_COND_ALL_FLOWS=1) vs one per flow (default), final kernels
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
for pf in 1 ""; do
NSW_COND_ALL_FLOWS=$pf timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_COND_ALL_FLOWS=$pf ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
