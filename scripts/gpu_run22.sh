#!/bin/bash
# timing experiments on the flow kernel (results numerically wrong by construction): what slows the MMA1 burst?
# NOTE: needs the NSW_FLOW_EXP instrumentation patch (strip switches in iaf_flow_tc_kernel) that was applied for this
# run only and is not part of the committed kernel; the numbers it produced are in profiles/r01/README.md
mkdir -p gpurun_out
for e in 0 31 1 2 4 8 16 6 30; do
  NSW_FLOW_EXP=$e timeout 300 python bench.py --steps 10 --warmup 3 --no-fastgen --no-distill --no-cpu-baseline > gpurun_out/exp22_$e.json 2> gpurun_out/exp22_$e.err
  NSW_FLOW_EXP=$e NSW_LAYER_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fastgen --no-distill > /dev/null 2> gpurun_out/dbg22_$e.err
  python - <<PY
import json,re
try:
    d=json.load(open('gpurun_out/exp22_$e.json'))
    txt=open('gpurun_out/dbg22_$e.err').read()
    m=re.findall(r'layers 0\.\.30 .*?\n\s*MMA1 \(operands ready, issued\)\s*:(.*)', txt)
    bursts=[]
    if m:
        for tok in m[-1].split():
            a,b=tok.split('-'); bursts.append(int(b)-int(a))
    print('exp $e: ms_per_step %.4f layers %.4f | MMA1 burst cycles %s' % (d['ms_per_step'], d['stage_ms']['layers'], bursts))
except Exception as ex: print('exp $e failed', ex)
PY
done
