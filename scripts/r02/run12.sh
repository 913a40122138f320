#!/bin/bash
# r02 run12: cheaper range tracking: A/B + guard tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trained_regime_gpu.py tests/test_iaf_tc_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r02_test12.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/r02_test12.log
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for i in 1 2 3; do
for lib in "" scripts/r02/_lib/libnsw_noguard.so; do
NSW_LIB=$lib timeout 600 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
