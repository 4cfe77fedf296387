#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_tc.py -m gpu -q -x -k "split" 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/r3g_tests.log
