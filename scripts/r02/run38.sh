#!/bin/bash
# r02 run38: cond projection epilogue that releases the accumulator stage before its stores, vs the previous build
# (PREV_LIB = libnsw_b200.so of the commit before; not kept in the tree)
mkdir -p gpurun_out
PREV_LIB=${PREV_LIB:-scripts/r02/_lib/libnsw_prev_cond_epilogue.so}
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
for lib in "" "$PREV_LIB"; do
NSW_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
NSW_COND_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 $LEAN 2>&1 >/dev/null | grep "cond_proj dbg" | tail -2
