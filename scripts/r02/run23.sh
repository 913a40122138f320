#!/bin/bash
# r02 run23: resize-conv tests with the calibrated bar, fastgen suite after the L2-priority reset, bench fastgen block,
# ncu of the teacher forward (distill block only), then the N=2 bench through torchrun
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_iaf_gpu.py tests/test_fastgen_gpu.py tests/test_fastgen_gn_gpu.py -m gpu -q --timeout 900 > gpurun_out/r02_test23.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/r02_test23.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-distill --no-clarinet --no-sustained --no-python-e2e > gpurun_out/r02_bench23.json 2> gpurun_out/r02_bench23.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench23.json'))
print('value',d['value'],'ms',d['ms_per_step'])
print('fastgen',json.dumps(d['fastgen'])[:1500])
PY
LEAN="--no-cpu-baseline --no-fastgen --no-clarinet --no-sustained --no-python-e2e"
timeout 900 ncu --set full --clock-control none -k regex:"teacher|conv_gemm|tc_gemm|mol|kl" -c 12 -o gpurun_out/r02_prof23_teacher python bench.py --steps 1 --warmup 1 $LEAN > gpurun_out/r02_ncu23_teacher.log 2>&1; echo "ncu teacher rc=$?"
tail -3 gpurun_out/r02_ncu23_teacher.log
