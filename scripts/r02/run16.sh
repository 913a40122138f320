#!/bin/bash
# r02 run16: cond projection: staggered n-tile order x {cta_group::2, 1-CTA pair kernel}; batched fastgen, 3 repeats per case
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_iaf_tc_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for rep in 1 2; do
for lib in "" "$NOSTAGGER_LIB"; do   # NOSTAGGER_LIB = a build with -DNSW_COND_STAGGER=0 (not kept in the tree)
for v in "" 1; do
NSW_LIB=$lib NSW_COND_1CTA=$v timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib NSW_COND_1CTA=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done; done; done
timeout 900 python scripts/r02/fastgen_batched_bench.py mol:gn:8 mol:gn:1 mol:gn:4 ce:gn:8 ce:gn:1 2>&1 | cut -c1-260
NSW_FASTGEN_L2LAST=0 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:8 ce:gn:8 2>&1 | cut -c1-260
