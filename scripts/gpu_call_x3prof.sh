#!/bin/bash
# ncu --set full of the split-operand kernel (first decoder launch of scripts/x3_bench.py) + compute-sanitizer over a small run
mkdir -p gpurun_out
echo "== x3 full (decoder launch)"; timeout 600 ncu --set full --clock-control none --import-source on -k x3_kernel -s 13 -c 1 -o gpurun_out/r02_x3_dec -f \
    python scripts/x3_bench.py 50000 > gpurun_out/r02_x3_dec_full.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/r02_x3_dec_full.log | head -5
echo "== x3 full (encoder launch)"; timeout 600 ncu --set full --clock-control none --import-source on -k x3_kernel -s 3 -c 1 -o gpurun_out/r02_x3_enc -f \
    python scripts/x3_bench.py 50000 > gpurun_out/r02_x3_enc_full.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/r02_x3_enc_full.log | head -5
bash scripts/gpu_call_sanitize.sh
