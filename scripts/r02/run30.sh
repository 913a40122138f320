#!/bin/bash
# r02 run30: teacher forward: conditioning fused into the dilated-conv GEMM + epilogue rows prefetched, vs the earlier schemes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py tests/test_distill_gpu.py tests/test_trained_regime_gpu.py tests/test_iaf_tc_gpu.py tests/test_fastgen_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r02_test30.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_test30.log
for rep in 1 2; do
echo "default:        $(REPS=5 python scripts/r02/teacher_only.py)"
echo "cond separate:  $(NSW_TEACHER_COND_SEPARATE=1 REPS=5 python scripts/r02/teacher_only.py)"
echo "1-CTA kernel:   $(NSW_GEMM_1CTA=1 REPS=5 python scripts/r02/teacher_only.py)"
done
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches30_teacher.csv python scripts/r02/teacher_only.py > /dev/null 2>&1; echo "list rc=$?"
