#!/bin/bash
# r02 run26: where the cond projection's MMA issuer waits (NSW_COND_DEBUG), 256- and 128-column items
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for tn in 256 128; do
echo "== NSW_COND_TN=$tn"
NSW_COND_TN=$tn NSW_COND_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 $LEAN 2>&1 >/dev/null | grep "cond_proj dbg" | tail -4
done
