#!/bin/bash
# r02 run25: one cond projection for all flows (shared upsampling stack) vs one per flow
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py tests/test_trained_regime_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2; do
for pf in "" 1; do
NSW_COND_PER_FLOW=$pf timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_COND_PER_FLOW=$pf ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
