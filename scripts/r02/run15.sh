#!/bin/bash
# r02 run15: 2-CTA MMA probe; cta_group::2 cond projection (parity + stage time, A/B with NSW_COND_1CTA=1);
# batched fastgen with K-split dots + L2 hints (parity + throughput)
mkdir -p gpurun_out
timeout 120 scripts/probes/umma_2cta_probe > gpurun_out/r02_umma_2cta_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02_umma_2cta_probe.log
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py tests/test_iaf_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/r02_test15_iaf.log 2>&1; echo "iaf tests rc=$?"
tail -3 gpurun_out/r02_test15_iaf.log
LEAN="--no-cpu-baseline --no-fastgen --no-distill --no-clarinet --no-python-e2e --no-sustained"
for v in "" 1; do
NSW_COND_1CTA=$v timeout 300 python bench.py --steps 40 --warmup 5 $LEAN 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NSW_COND_1CTA=$v ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()})"
done
timeout 900 python -m pytest tests/test_fastgen_gn_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/r02_test15_gn.log 2>&1; echo "gn tests rc=$?"
tail -3 gpurun_out/r02_test15_gn.log
timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:1 mol:gn:4 mol:gn:8 mol:gn:8 ce:gn:1 ce:gn:8 ce:gn:8 2>&1 | cut -c1-220
NSW_FASTGEN_L2LAST=0 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:8 ce:gn:8 2>&1 | cut -c1-220
