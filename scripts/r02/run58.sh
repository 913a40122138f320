#!/bin/bash
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
NSW_FLOW_PAIR=1 NSW_FLOW_PAIR_VERBOSE=1 timeout 200 python bench.py --steps 5 --warmup 2 $LEAN 2>&1 | grep "flow_pair\|Error\|error" | head -5
NSW_FLOW_PAIR=1 NSW_FLOW_PAIR_NOCOOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches58.csv python bench.py --steps 1 --warmup 1 $LEAN > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches58.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
names=rows[hdr]; ki=names.index('Kernel Name'); vi=names.index('Metric Value'); gi=names.index('Grid Size')
for r in rows[hdr+1:]:
    if r[0].isdigit() and ('flow' in r[ki] or 'cond' in r[ki]): print(r[ki][:50], r[vi], r[gi])
PY
grep -i "error" gpurun_out/r02_launches58.csv | head -3
