#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --no-fastgen --no-distill > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -3 gpurun_out/bench_n8.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n8.json')); print(d['n_gpus'], d['value'], d['rtf'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
