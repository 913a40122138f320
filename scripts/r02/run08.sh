#!/bin/bash
# r02 run08: whole GPU suite + smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_test08.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r02_test08.log
python __graft_entry__.py --smoke 2>&1 | tail -2
