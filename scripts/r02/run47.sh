#!/bin/bash
# r02 run47: batched fastgen: flag-array grid barrier (NSW_GN_FLAGS=2) vs the atomic counter
for f in 0 2 0 2; do
echo "== NSW_GN_FLAGS=$f"
NSW_GN_FLAGS=$f REPS=3 T=4000 timeout 600 python scripts/r02/fastgen_batched_bench.py mol:gn:8 mol:gn:1 ce:gn:8 2>&1 | grep "us_per_step" | cut -c1-200
done
NSW_GN_FLAGS=2 timeout 900 python -m pytest tests/test_fastgen_gn_gpu.py -m gpu -q --timeout 600 2>&1 | tail -2
