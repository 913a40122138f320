#!/bin/bash
# r02 run57: CTA-pair flow kernel (NSW_FLOW_PAIR=1): parity, then A/B against the single-CTA kernel
mkdir -p gpurun_out
NSW_FLOW_PAIR=1 timeout 600 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -m gpu -q -x --timeout 200 > gpurun_out/r02_test57.log 2>&1; echo "pair parity rc=$?"
grep -n "watchdog\|Error\|error\|passed\|failed" gpurun_out/r02_test57.log | head -12
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
for v in 1 0; do
NSW_FLOW_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_FLOW_PAIR=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
