#!/bin/bash
# compute-sanitizer over one small run of the split-operand kernel (memcheck, synccheck, racecheck)
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool"; timeout 600 compute-sanitizer --tool $tool python scripts/x3_small.py 5 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r02_x3_sanitize_$tool.log
done
