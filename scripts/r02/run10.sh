#!/bin/bash
# r02 run10: split weight barriers in the flow kernel: parity, timeline, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py tests/test_trained_regime_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r02_test10.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/r02_test10.log
timeout 300 python scripts/r02/flow_debug.py 2 > gpurun_out/r02_flow_debug10.log 2>&1; echo rc=$?
grep -E "layer start|end " gpurun_out/r02_flow_debug10.log | cut -c1-330 | head -4
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e"
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'sustained', d['sustained']['ms_per_step_median'], d['stage_ms'])"; done
