#!/bin/bash
# r02 run50: 8-GPU bench through torchrun, as the driver launches it (final tree)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench50_n8.json 2> gpurun_out/r02_bench50_n8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/r02_bench50_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench50_n8.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('distill',d.get('distill',{}).get('ms'),'clarinet',d.get('clarinet',{}).get('value'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02_bench50_n8_ref.json 2>/dev/null; echo "ref n8 rc=$?"; tail -c 300 gpurun_out/r02_bench50_n8_ref.json
