#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/probes/umma_2cta_probe > gpurun_out/r02_umma_2cta_probe.log 2>&1; echo rc=$?
cat gpurun_out/r02_umma_2cta_probe.log
