#!/bin/bash
# r02 run62: CTA-scope remote arrives in the pair GEMM kernels too: parity suites, then timings (student stages, teacher)
timeout 1200 python -m pytest tests/test_conv_gemm_gpu.py tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py tests/test_teacher_gpu.py tests/test_fastgen_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
for v in 1 0; do
NSW_FLOW_PAIR=$v timeout 200 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_FLOW_PAIR=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done
REPS=5 python scripts/r02/teacher_only.py
