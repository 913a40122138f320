#!/bin/bash
# r02 run61: pair flow kernel with CTA-scope remote arrives: parity, timeline, A/B
NSW_FLOW_PAIR=1 timeout 600 python -m pytest tests/test_iaf_tc_gpu.py -m gpu -q -x --timeout 200 2>&1 | tail -2
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
NSW_FLOW_PAIR=1 NSW_FLOW_PAIR_DEBUG=1 NSW_FLOW_PAIR_DEBUG_PAIR=0 timeout 200 python bench.py --steps 1 --warmup 1 $LEAN 2>&1 >/dev/null | grep -A5 "flow_pair dbg" | tail -6 | cut -c1-900
for rep in 1 2 3; do
for v in 1 0; do
NSW_FLOW_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_FLOW_PAIR=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
