#!/bin/bash
# new: Gaussian KL (kl_loss_gauss) kernel + ClariNet distillation pipeline test, fastgen switch-variant test, smoke with fastgen
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py tests/test_fastgen_gpu.py -m gpu -x -q --timeout 900 -s > gpurun_out/test50.log 2>&1; echo "gpu tests rc=$?"
grep -i "gauss kl\|clarinet dist\|passed\|failed\|error" gpurun_out/test50.log | cut -c1-300 | tail -12
python __graft_entry__.py --smoke 2>&1 | tail -2
