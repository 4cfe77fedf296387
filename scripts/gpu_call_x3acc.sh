#!/bin/bash
# accuracy of the split-operand path with bf16 and with fp16 operand terms (two builds of the library), then the x3 tests
mkdir -p gpurun_out
echo "== bf16 split"; timeout 700 python scripts/x3_accuracy.py 2>&1 | tail -60 | tee gpurun_out/r02_x3_accuracy_bf16.json
echo "== fp16 split"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_f16.so timeout 700 python scripts/x3_accuracy.py 2>&1 | tail -60 | tee gpurun_out/r02_x3_accuracy_fp16.json
echo "== fp16 split: x3 tests"; TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_f16.so timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_x3_f16_tests.log
