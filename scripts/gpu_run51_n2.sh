#!/bin/bash
# the driver's own N=2 command (no extra flags): rank 0 also runs the fastgen / distill secondaries
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench51_n2.json 2> gpurun_out/bench51_n2.err; echo "rc=$?"
tail -2 gpurun_out/bench51_n2.err
python -c "
import json
d=json.load(open('gpurun_out/bench51_n2.json')); print(d['n_gpus'], d['value'], d['rtf'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d.get('fastgen',{}).get('rtf'), d.get('distill',{}).get('ms'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench51_n2_ref.json 2> gpurun_out/bench51_n2_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench51_n2_ref.json
