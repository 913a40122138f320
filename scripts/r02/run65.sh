#!/bin/bash
# r02 run65 (= run54 on the last commit): the driver's round-end commands on the final tree, verbatim
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_test65.log 2>&1; echo "pytest -m gpu rc=$?"
tail -3 gpurun_out/r02_test65.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke() returned')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench65.json 2> gpurun_out/r02_bench65.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench65.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
