#!/bin/bash
# r02 run60: CTA-pair flow kernel: issuer wait accounting and the timeline of tasks 8..15 (layers 2-3 of a K = 4 pair)
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
NSW_FLOW_PAIR=1 NSW_FLOW_PAIR_DEBUG=1 NSW_FLOW_PAIR_DEBUG_PAIR=0 timeout 200 python bench.py --steps 1 --warmup 1 $LEAN 2>&1 >/dev/null | grep -A5 "flow_pair dbg" | tail -12 | cut -c1-900
