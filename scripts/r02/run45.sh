#!/bin/bash
# r02 run45: upsampling stack: layer 2 over all clips as one flattened clip (zero-padded intermediate)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_iaf_gpu.py tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py tests/test_teacher_gpu.py tests/test_conv_gemm_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2 3; do
timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches45.csv python bench.py --steps 1 --warmup 1 $LEAN > /dev/null 2>&1
grep -i "conv_gemm" gpurun_out/r02_launches45.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -6
