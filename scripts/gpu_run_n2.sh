#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-fastgen > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n2.json')); print(d['n_gpus'], d['value'], d['rtf'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
