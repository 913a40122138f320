#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) and synccheck of the final fastgen kernel on a short utterance
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python scripts/fastgen_exp.py --steps 40 --flags default > gpurun_out/racecheck55.log 2>&1; echo "racecheck rc=$?"
grep -i "racecheck summary\|hazard\|flags" gpurun_out/racecheck55.log | sort | uniq -c | sort -rn | head -12
timeout 300 compute-sanitizer --tool synccheck python scripts/fastgen_exp.py --steps 40 --flags default > gpurun_out/synccheck55.log 2>&1; echo "synccheck rc=$?"
grep -i "error summary\|flags" gpurun_out/synccheck55.log | tail -3
