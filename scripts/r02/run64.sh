#!/bin/bash
# r02 run64: 2-GPU bench through torchrun, as the driver launches it (student + clarinet + distill keys, max over ranks)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench64_n2.json 2> gpurun_out/r02_bench64_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r02_bench64_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench64_n2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('distill',d.get('distill',{}).get('ms'),'clarinet',d.get('clarinet',{}).get('value'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench64_n2_ref.json 2>/dev/null; echo "ref n2 rc=$?"; tail -c 300 gpurun_out/r02_bench64_n2_ref.json
