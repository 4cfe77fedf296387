#!/bin/bash
# two-launch backward with the first part's weight gradients beside the last wave: tests + A/B of the training step
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "train or split or grad or loop" 2>&1 | tail -6 | tee gpurun_out/r3f_tests.log
echo "== train bench, overlap on"; timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | tee gpurun_out/r3f_train_overlap.json
echo "== train bench, overlap off"; TURBOAE_B200_WGRAD_OVERLAP=0 timeout 300 python scripts/train_bench.py 2>&1 | tail -1 | tee gpurun_out/r3f_train_single.json
