#!/bin/bash
# 96-channel weight-gradient slabs: training parity tests, training bench, ncu of wgrad_kernel
mkdir -p gpurun_out
echo "== training tests"; timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "train or grad or backward" 2>&1 | tail -5 | tee gpurun_out/r02_wgrad96_tests.log
echo "== train bench"; timeout 300 python scripts/train_bench.py 2>/dev/null | tail -1 | tee gpurun_out/r02_wgrad96_train.json
echo "== wgrad full"; TRAIN_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 -o gpurun_out/r02_wgrad96 -f \
    python scripts/train_bench.py > gpurun_out/r02_wgrad96.log 2>&1; grep -i "error\|==PROF== Report" gpurun_out/r02_wgrad96.log | head -5
