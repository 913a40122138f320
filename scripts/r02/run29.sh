#!/bin/bash
# r02 run29: ncu --set full of the teacher's layer GEMMs, pair kernel and 128 x 128 kernel
mkdir -p gpurun_out
REPS=4 python scripts/r02/teacher_only.py
NSW_GEMM_1CTA=1 REPS=4 python scripts/r02/teacher_only.py
REPS=2 timeout 600 ncu --set full --clock-control none -k regex:"conv_gemm_tc" -s 68 -c 4 -o gpurun_out/r02_prof29_pair python scripts/r02/teacher_only.py > gpurun_out/r02_ncu29_pair.log 2>&1; echo "ncu rc=$?"
NSW_GEMM_1CTA=1 REPS=2 timeout 600 ncu --set full --clock-control none -k regex:"conv_gemm_tc" -s 68 -c 4 -o gpurun_out/r02_prof29_1cta python scripts/r02/teacher_only.py > gpurun_out/r02_ncu29_1cta.log 2>&1; echo "ncu rc=$?"
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches29_teacher.csv python scripts/r02/teacher_only.py > /dev/null 2>&1; echo "list rc=$?"
