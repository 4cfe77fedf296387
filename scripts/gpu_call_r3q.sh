#!/bin/bash
# wgrad pipeline: 4 stages of a quarter group (default build) vs 2 stages of half a group (lib/libtae_wgrad2.so): tests + A/B
mkdir -p gpurun_out
echo "== wgrad + training tests (4 stages)"; timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "train or wgrad or split or grad or loop" 2>&1 | tail -4 | tee gpurun_out/r3q_tests.log
echo "== 4 stages"; timeout 300 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee gpurun_out/r3q_stages4.json
echo "== 2 stages"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libtae_wgrad2.so timeout 300 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee gpurun_out/r3q_stages2.json
echo "== 4 stages again"; timeout 300 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee -a gpurun_out/r3q_stages4.json
