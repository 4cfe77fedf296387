#!/bin/bash
# 2 GPUs: the NCCL test (sharded encode + data-parallel training step) and bench.py --gpus 2 with the leaner training step
mkdir -p gpurun_out
echo "== 2-rank test"; timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r3d_multirank.log
bash scripts/gpu_call_n2b.sh r3d_n2
