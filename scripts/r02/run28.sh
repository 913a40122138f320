#!/bin/bash
# r02 run28: pair conv-GEMM with 8 epilogue warps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py tests/test_distill_gpu.py tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r02_test28.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/r02_test28.log
LEAN="--no-cpu-baseline --no-fastgen --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2; do
for v in "" 1; do
NSW_GEMM_1CTA=$v timeout 600 python bench.py --steps 20 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_GEMM_1CTA=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, 'distill', d['distill'].get('ms'), 'teacher', d['distill'].get('teacher_forward_ms'))"
done; done
