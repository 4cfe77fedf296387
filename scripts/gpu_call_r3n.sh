#!/bin/bash
# weight-gradient jobs split by taps (N = 112) instead of channel slabs: tests + A/B of the graph-replayed training step
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_parity.py -m gpu -q -x -k "train or wgrad or split or grad or loop" 2>&1 | tail -4 | tee gpurun_out/r3n_tests.log
echo "== tap split on"; timeout 300 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee gpurun_out/r3n_tapsplit_on.json
echo "== tap split off"; TURBOAE_B200_WGRAD_TAPSPLIT=0 timeout 300 python scripts/train_overlap_ab.py 2>&1 | tail -1 | tee gpurun_out/r3n_tapsplit_off.json
