#!/bin/bash
# r02 run39: flow kernel: dilated-conv weights released / awaited tap by tap around the layer boundary, vs the previous
# build (PREV_LIB = libnsw_b200.so of the commit before; not kept in the tree)
mkdir -p gpurun_out
PREV_LIB=${PREV_LIB:-scripts/r02/_lib/libnsw_prev.so}
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py tests/test_trained_regime_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
for lib in "" "$PREV_LIB"; do
NSW_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
NSW_LAYER_DEBUG=1 timeout 300 python scripts/r02/flow_debug.py 2>&1 | grep -A9 "cta 0" | head -24
